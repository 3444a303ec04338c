import sys, hashlib, time
sys.path.insert(0, '.')
import numpy as np
import mnem_b200 as mb
from mnem_b200 import simdata
cfgno = int(sys.argv[1]) if len(sys.argv) > 1 else 4
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = dict(simdata.CONFIGS[cfgno])
sim = simdata.make_config(cfg, seed=1000 + cfgno)
starts = 3
phi = simdata.init_phi(sim["S"], k, starts, seed=3000 + 10 * cfgno + k)
def dig(r, s):
    return hashlib.sha1(np.ascontiguousarray(r["adj"][s]).tobytes() + np.ascontiguousarray(r["theta"][s]).tobytes()).hexdigest()[:10]
with mb.Context.synthetic(sim) as cx:
    for rep in range(3):
        t0 = time.time()
        r = mb.em_run(cx, phi, complete=sim["complete"], want_probs=False)
        print("all3 rep", rep, round(time.time() - t0, 3), [dig(r, s) for s in range(starts)], r["n_iter"].tolist(), [float(x) for x in r["ll"]])
    for s in range(starts):
        it = iter([s])
        r1 = mb.em_run(cx, phi, complete=sim["complete"], want_probs=False, next_start=lambda: next(it, -1))
        print("only start", s, dig(r1, s), r1["n_iter"].tolist(), float(r1["ll"][s]))
    r2 = mb.em_run(cx, phi[1:2], complete=sim["complete"], want_probs=False)
    print("start 1 alone (B=1):", dig(r2, 0), r2["n_iter"].tolist(), float(r2["ll"][0]))
