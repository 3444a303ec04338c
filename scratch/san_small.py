import sys
sys.path.insert(0, '.')
import numpy as np
import mnem_b200 as mb
from mnem_b200 import simdata
rng = np.random.default_rng(3)
R = rng.normal(size=(6, 130, 9)) * 3
out = mb.nem_greedy(R, None)
print("greedy ok", out["n_iter"].tolist())
sim = simdata.make_config(dict(S=6, Egenes=8, C=700, mw=[0.3, 0.7], k=2, starts=5, uninform=2), seed=77)
D = simdata.host_data(sim)
init = simdata.init_phi(6, 2, 5, seed=9)
with mb.Context(D, sim["pert"]) as cx:
    r = mb.em_run(cx, init, batch=2, tape=True, probs="winner")
print("em ok", r["n_iter"].tolist(), r["best"])
