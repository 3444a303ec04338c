import sys
sys.path.insert(0, '.')
import numpy as np
import mnem_b200 as mb, oracle as orc
from mnem_b200 import simdata
from mnem_b200.api import parse_tape
sim = simdata.make_config(dict(S=6, Egenes=8, C=3000, mw=[0.3, 0.7], k=2, starts=4, uninform=2), seed=77)
D = simdata.host_data(sim)
init = simdata.init_phi(6, 2, 4, seed=9)
S, E, k = 6, sim["E"], 2
mw0 = np.array([0.85, 0.15])
with mb.Context(D, sim["pert"]) as cx:
    res = [dict(adj=init[0][i]) for i in range(k)]
    g = mb.getProbs(np.zeros((k, sim["C"])), k, cx, res, mw=mw0)
    r = mb.em_run(cx, init[:1], mw=mw0, tape=True)
    ru = mb.em_run(cx, init[:1], tape=True)
o = orc.get_probs(D, sim["pert"], S, init[0], None, np.zeros((k, sim["C"])), mw0)
print("getProbs theta-free mw0: ll gpu", g["ll"], "orc", o["ll"], "mw gpu", g["mw"], "orc", o["mw"], "theta equal", np.array_equal(g["theta_used"], o["theta_used"]))
rec = orc.record_start(D, sim["pert"], S, init[0], mw=mw0)
t_orc = parse_tape(rec["words"], k, E, S)[0]
for t in range(10):
    a, b, u = r["tape"][0]["init_theta"][t], t_orc["init_theta"][t], ru["tape"][0]["init_theta"][t]
    print("inner", t, "gpu(mw0) vs orc(mw0) differing", int((a != b).sum()), "| gpu(mw0) vs gpu(uniform)", int((a != u).sum()), "| orc(mw0) vs gpu(uniform)", int((b != u).sum()))
print("final theta of getProbs (gpu) vs tape inner 9:", int((g["theta_used"] != r["tape"][0]["init_theta"][9]).sum()))
