"""Randomised differential check (run on a B200): random small configs -> device fit with tape -> oracle replay.
Prints one line per config and a summary; exits 1 on the first genuine disagreement."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import mnem_b200 as mb, oracle as orc
from mnem_b200 import simdata
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
forced_total = 0
t0 = time.time()
for it in range(n):
    S = int(rng.integers(2, int(sys.argv[3]) if len(sys.argv) > 3 else 13)); Eg = int(rng.integers(1, 9)); k = int(rng.integers(2, 5)); starts = int(rng.integers(1, 6))
    Cn = int(rng.integers(S * 3, int(sys.argv[4]) if len(sys.argv) > 4 else 4000)); un = int(rng.integers(0, 4)); complete = bool(rng.integers(0, 2))
    batch = int(rng.integers(0, 4)); nullcomp = bool(rng.integers(0, 5) == 0); use_mw = bool(rng.integers(0, 4) == 0)
    sim = simdata.make_config(dict(S=S, Egenes=Eg, C=Cn, mw=[1.0 / k] * k, k=k, starts=starts, uninform=un), seed=int(rng.integers(1 << 30)))
    D = simdata.host_data(sim) * float(rng.choice([0.3, 1.0, 1.0, 3.0]))
    kk = k + (1 if nullcomp else 0)
    init = simdata.init_phi(S, kk, starts, seed=int(rng.integers(1 << 30)))
    if nullcomp:
        init[:, -1] = 0
    mw = rng.dirichlet(np.ones(kk)) if use_mw else None
    try:
        with mb.Context(D, sim["pert"]) as cx:
            r = mb.em_run(cx, init, complete=complete, batch=batch, tape=True, max_iter_guard=40, nullcomp=nullcomp, mw=mw)
    except mb._lib.MnemcuError as ex:
        print(it, "device error", ex); bad += 1; continue
    ok = True
    for s in range(starts):
        o = orc.replay_start(D, sim["pert"], S, init[s], r["tape"][s], complete=complete, max_iter_guard=40, nullcomp=nullcomp, mw=mw)
        forced = sum(v["forced"] for v in o["report"].values())
        forced_total += forced
        same = (not o["diverged"]) and np.array_equal(r["adj"][s], o["adj"]) and np.array_equal(r["theta"][s], o["theta"]) and r["n_iter"][s] == o["n_iter"]
        fin = np.isfinite(o["ll"]) and np.isfinite(r["ll"][s])
        close = (abs(r["ll"][s] - o["ll"]) <= 1e-9 * max(abs(o["ll"]), 1e-300)) if fin else (str(r["ll"][s]) == str(o["ll"]))
        if not (same and close):
            ok = False
            print("  MISMATCH start", s, o["first_hard"], r["n_iter"][s], o["n_iter"], r["ll"][s], o["ll"])
    print(it, dict(S=S, E=sim["E"], C=sim["C"], k=kk, starts=starts, complete=complete, batch=batch, nullcomp=nullcomp, mw=use_mw), "ok" if ok else "BAD")
    bad += 0 if ok else 1
print("configs", n, "bad", bad, "decisions overridden", forced_total, "seconds", round(time.time() - t0, 1))
sys.exit(1 if bad else 0)
