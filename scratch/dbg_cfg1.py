import sys
sys.path.insert(0, '.')
import numpy as np
import mnem_b200 as mb, oracle as orc
from mnem_b200 import simdata
sim = simdata.make_config(1)
D = simdata.host_data(sim)
init = simdata.init_phi(sim["S"], sim["k"], sim["starts"], seed=1)
ref = orc.mnem_em(D, sim["pert"], sim["S"], init)
with mb.Context(D, sim["pert"]) as cx:
    r = mb.em_run(cx, init, tape=True)
for s in range(sim["starts"]):
    o = orc.replay_start(D, sim["pert"], sim["S"], init[s], r["tape"][s])
    rel = np.max(np.abs(r["lls"][s] - ref["lls"][s][:len(r["lls"][s])]) / np.abs(ref["lls"][s][:len(r["lls"][s])])) if len(r["lls"][s]) == len(ref["lls"][s]) else -1
    forced = {q: (v["forced"], v["max_forced"]) for q, v in o["report"].items() if v["forced"]}
    print(s, "diverged", o["diverged"], o["first_hard"], "lls rel vs unforced oracle", rel, "vs replay", np.max(np.abs(r["lls"][s] - o["lls"]) / np.abs(o["lls"])), "forced", forced,
          "min_margin", ref["min_margin"][s], "accept margin", ref["min_accept_margin"][s])
