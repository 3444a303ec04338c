#!/usr/bin/env python
"""Model selection on the perturb-seq-shaped config (BASELINE cfg4: 24 S-genes x 1000 E-genes x 200k cells, k = 1..5 by
getIC, 3 starts per k) with the (k, start) fits spread over the GPUs of one box (mnem_b200.dist.mnemk_sharded).

  python scripts/mnemk_bench.py                                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         scripts/mnemk_bench.py                                       # 8 GPUs
Prints one JSON line (rank 0): chosen k, the information criteria, per-k winners' ll and a digest of their graphs, and
the device-synchronised wall time of the sweep (max over ranks).  The chosen k, ics and digests must not depend on N."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--ks", default="1,2,3,4,5")
    ap.add_argument("--cells", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    import mnem_b200 as mb
    from mnem_b200 import dist as mdist
    from mnem_b200 import simdata
    rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(simdata.CONFIGS[a.config])
    if a.cells:
        cfg["C"] = a.cells
    sim = simdata.make_config(cfg, seed=1000 + a.config)
    ks = [int(x) for x in a.ks.split(",")]
    starts = sim["starts"]
    phi = {k: simdata.init_phi(sim["S"], k, starts, seed=3000 + 10 * a.config + k) for k in ks if k > 1}
    times = []
    with mb.Context.synthetic(sim, device=local) as cx:
        for rep in range(a.reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.time()
            r = mdist.mnemk_sharded(cx, ks, starts, phi=phi, complete=sim["complete"], device=dev)
            torch.cuda.synchronize()
            t = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
    if rank == 0:
        dig = {str(k): hashlib.sha1(np.ascontiguousarray(r["fits"][k]["adj"]).tobytes() +
                                    np.ascontiguousarray(r["fits"][k]["theta"]).tobytes()).hexdigest()[:16] for k in ks}
        print(json.dumps(dict(workload="cfg%d mnemk: S=%d E=%d C=%d ks=%s starts=%d" % (a.config, sim["S"], sim["E"], sim["C"], a.ks, starts),
                              n_gpus=world, k=r["k"], ics={str(k): v for k, v in r["ics"].items()},
                              lls={str(k): v for k, v in r["lls"].items()}, digests=dig,
                              owners={str(k): r["fits"][k]["owner"] for k in ks}, seconds=times, seconds_best=min(times))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
