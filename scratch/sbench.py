"""Greedy-search timing: n searches of S x E in lock step through mnemcu_nem_greedy (includes H2D/D2H of R)."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
import mnem_b200 as mb
S = int(sys.argv[1]) if len(sys.argv) > 1 else 30
E = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rep = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rng = np.random.default_rng(0)
def planted():
    a = np.triu((rng.random((S, S)) < 0.3).astype(np.uint8), 1)
    t = mb.transitive_closure(a)
    p = rng.permutation(S); t = t[np.ix_(p, p)]
    th = rng.integers(1, S + 1, E)
    return np.where(t[:, th - 1].T > 0, 1.0, -1.0) * 4 + rng.normal(size=(E, S)) * 3
R = np.stack([planted() for _ in range(n)])
for r in range(rep):
    l0 = mb.launch_count(); t0 = time.time()
    out = mb.nem_greedy(R, None)
    dt = time.time() - t0
    print(json.dumps(dict(S=S, E=E, n=n, seconds=dt, n_iter=int(out["n_iter"].max()), iters_sum=int(out["n_iter"].sum()),
                          scored=int(out["n_scored"].sum()), cand_per_s=float(out["n_scored"].sum() / dt),
                          launches=mb.launch_count() - l0)))
