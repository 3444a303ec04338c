import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
import mnem_b200 as mb
from mnem_b200 import simdata
cfg = dict(simdata.CONFIGS[5]); cfg["starts"] = int(sys.argv[1]) if len(sys.argv) > 1 else 12
sim = simdata.make_config(cfg, seed=1005)
init = simdata.init_phi(sim["S"], sim["k"], sim["starts"], seed=2005)
E, Cn = sim["E"], sim["C"]
host = torch.empty((Cn, E), dtype=torch.float64, pin_memory=True)
hnp = host.numpy()
with mb.Context.synthetic(sim) as c0:
    mb._lib.check(mb.load_library().mnemcu_ctx_read_data(c0._h, 0, Cn, mb._lib.dp(hnp)))
Dh = hnp.T
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    cx = mb.Context(Dh, sim["pert"])
    torch.cuda.synchronize(); t1 = time.time()
    r = mb.em_run(cx, init, complete=True, probs="winner")
    torch.cuda.synchronize(); t2 = time.time()
    cx.close()
    torch.cuda.synchronize(); t3 = time.time()
    print(json.dumps(dict(rep=rep, ctx_create=t1 - t0, em_run_wall=t2 - t1, em_run_device=r["stats"]["ms_total"] * 1e-3, close=t3 - t2)))
