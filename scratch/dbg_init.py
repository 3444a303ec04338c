import sys, json
sys.path.insert(0, '.')
import numpy as np
import mnem_b200 as mb, oracle as orc
from mnem_b200 import simdata
from mnem_b200.api import parse_tape
sim = simdata.make_config(dict(S=6, Egenes=8, C=3000, mw=[0.3, 0.7], k=2, starts=4, uninform=2), seed=77)
D = simdata.host_data(sim)
init = simdata.init_phi(6, 2, 4, seed=9)
S, E, k = 6, sim["E"], 2
for mw0 in (None, np.array([0.85, 0.15])):
    with mb.Context(D, sim["pert"]) as cx:
        r = mb.em_run(cx, init, mw=mw0, tape=True)
        r1 = mb.em_run(cx, init[:1], mw=mw0, tape=True)
    print("mw0", mw0, "batch4-vs-batch1 start0 init theta equal:", all(np.array_equal(r["tape"][0]["init_theta"][t], r1["tape"][0]["init_theta"][t]) for t in range(10)))
    for s in range(4):
        rec = orc.record_start(D, sim["pert"], S, init[s], acc="ld", start=s) if mw0 is None else None
        o = orc.replay_start(D, sim["pert"], S, init[s], r["tape"][s], mw=mw0)
        print(" start", s, "diverged", o["diverged"], o["first_hard"], "n_iter gpu", r["n_iter"][s], "orc", o["n_iter"])
        if rec is not None:
            t_orc = parse_tape(rec["words"], k, E, S)[s]
            for t in range(10):
                a, b = r["tape"][s]["init_theta"][t], t_orc["init_theta"][t]
                if not np.array_equal(a, b):
                    print("   inner", t, "theta differs at", np.argwhere(a != b).tolist()[:5], "gpu", a[a != b][:5], "orc", b[a != b][:5])
            print("   init_best gpu", [r["tape"][s]["init_best"][t] for t in range(10)], "orc", [t_orc["init_best"][t] for t in range(10)])
# theta-free getProbs standalone
with mb.Context(D, sim["pert"]) as cx:
    res = [dict(adj=init[0][i]) for i in range(k)]
    g = mb.getProbs(np.zeros((k, sim["C"])), k, cx, res)
o = orc.get_probs(D, sim["pert"], S, init[0], None, np.zeros((k, sim["C"])), None)
print("getProbs theta-free: ll gpu", g["ll"], "orc", o["ll"], "theta equal", np.array_equal(g["theta_used"], o["theta_used"]), "probs maxabs", np.abs(g["probs"] - o["probs"]).max())
