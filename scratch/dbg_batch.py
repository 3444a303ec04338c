import sys, numpy as np
sys.path.insert(0, '.')
import mnem_b200 as mb
from mnem_b200 import simdata
sim = simdata.make_config(dict(S=20, Egenes=50, C=100000, mw=[0.2] * 5, k=5, starts=6, uninform=0, complete=True), seed=1003)
cx = mb.Context.synthetic(sim)
init = simdata.init_phi(sim["S"], 5, 6, seed=77)
a = mb.em_run(cx, init, complete=True, batch=4)
c = mb.em_run(cx, init, complete=True, batch=1)
np.set_printoptions(precision=17, linewidth=200)
for s in range(6):
    same = np.array_equal(a["adj"][s], c["adj"][s]) and np.array_equal(a["theta"][s], c["theta"][s])
    print("start", s, "same graphs", same, "n_iter", a["n_iter"][s], c["n_iter"][s], "ll", a["ll"][s], c["ll"][s])
    if not same or a["n_iter"][s] != c["n_iter"][s]:
        n = min(len(a["lls"][s]), len(c["lls"][s]))
        print("  lls a", a["lls"][s][:n+1]); print("  lls c", c["lls"][s][:n+1])
        print("  phievo", a["phievo"][s], c["phievo"][s]); print("  thetaevo", a["thetaevo"][s], c["thetaevo"][s])
