import sys, numpy as np
sys.path.insert(0, '.')
import mnem_b200 as mb, oracle as orc
from mnem_b200 import simdata
sim = simdata.make_config(dict(S=6, Egenes=5, C=1500, mw=[0.3, 0.3, 0.4], k=3, starts=5, uninform=2), seed=12)
D = simdata.host_data(sim) + np.random.default_rng(3).normal(size=(sim["E"], sim["C"])) * 0.5
init = simdata.init_phi(sim["S"], 3, 5, seed=2)
ref = orc.mnem_em(D, sim["pert"], sim["S"], init)
res = mb.mnem(D, sim["pert"], phi=init, batch=1)
np.set_printoptions(precision=12, linewidth=200)
for s in range(5):
    lim = res["limits"][s]
    print("start", s, "n_iter", res["n_iter"][s], ref["n_iter"][s], "ll", lim["ll"], ref["ll"][s])
    print("  lls gpu", lim["lls"]); print("  lls ref", ref["lls"][s])
    print("  mwevo gpu", lim["mwevo"]); print("  mwevo ref", ref["mwevo"][s])
    print("  phievo", lim["phievo"], ref["phievo"][s], "thetaevo", lim["thetaevo"], ref["thetaevo"][s])
    print("  mw gpu", lim["mw"]); print("  mw ref", ref["mw"][s])
    print("  probs maxdiff", np.abs(lim["probs"]-ref["probs"][s]).max())
