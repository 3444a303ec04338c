"""Randomised differential check of the Rho (multi-knock-down) path: device fit with tape -> oracle replay."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import mnem_b200 as mb, oracle as orc
from mnem_b200 import simdata
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(5)
bad = 0; forced_total = 0
for it in range(n):
    S = int(rng.integers(3, 9)); Eg = int(rng.integers(2, 7)); k = int(rng.integers(2, 4)); starts = int(rng.integers(1, 4))
    Cn = int(rng.integers(200, 2500)); complete = bool(rng.integers(0, 2))
    m = simdata.make_multi(S=S, Egenes=Eg, C=Cn, mw=tuple([1.0 / k] * k), multi=(0.25, 0.1), uninform=int(rng.integers(0, 3)), seed=int(rng.integers(1 << 30)))
    init = simdata.init_phi(S, k, starts, seed=int(rng.integers(1 << 30)))
    with mb.Context(m["D"], m["labels"]) as cx:
        r = mb.em_run(cx, init, complete=complete, tape=True, max_iter_guard=40)
    ok = True
    for s in range(starts):
        o = orc.replay_start(m["D"], m["pert_enc"], S, init[s], r["tape"][s], complete=complete, max_iter_guard=40)
        forced_total += sum(v["forced"] for v in o["report"].values())
        same = (not o["diverged"]) and np.array_equal(r["adj"][s], o["adj"]) and np.array_equal(r["theta"][s], o["theta"]) and r["n_iter"][s] == o["n_iter"]
        fin = np.isfinite(o["ll"]) and np.isfinite(r["ll"][s])
        close = (abs(r["ll"][s] - o["ll"]) <= 1e-9 * max(abs(o["ll"]), 1e-300)) if fin else (str(r["ll"][s]) == str(o["ll"]))
        if not (same and close):
            ok = False
            print("  MISMATCH start", s, o["first_hard"], r["n_iter"][s], o["n_iter"], r["ll"][s], o["ll"])
    print(it, dict(S=S, E=m["E"], C=Cn, k=k, starts=starts, complete=complete), "ok" if ok else "BAD")
    bad += 0 if ok else 1
print("rho configs", n, "bad", bad, "decisions overridden", forced_total)
sys.exit(1 if bad else 0)
