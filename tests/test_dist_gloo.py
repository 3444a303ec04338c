"""The N > 1 path on CPU: world_size-2 gloo processes shard the EM starts, each fits its shard (the CPU oracle stands
in for the per-rank GPU fit -- this test is about the host-side sharding and the winner reduction, not about kernels)
and the (ll, start id) reduction must pick the same winner as a single process (R/mnems.r:2222-2230)."""
import os
import socket
import subprocess
import sys

import numpy as np

from mnem_b200 import dist as mdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import oracle as orc
from mnem_b200 import simdata, dist as mdist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
sim = simdata.make_config(dict(S=4, Egenes=3, C=300, mw=[0.5, 0.5], k=2, starts=5, uninform=0), seed=3)
D = simdata.host_data(sim)
init = simdata.init_phi(4, 2, 5, seed=9)
my = mdist.shard_starts(5, rank, world)
r = orc.mnem_em(D, sim["pert"], 4, init[my])
winner, lls = mdist.gather_winner(my, r["ll"], 5)
print(json.dumps(dict(rank=rank, my=my, winner=int(winner), lls=[float(x) for x in lls])))
dist.destroy_process_group()
"""


# the shared start queue: ranks pull tickets from the process group's store until it is drained
WORKER_QUEUE = r"""
import os, sys, json, time
import numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import oracle as orc
from mnem_b200 import simdata, dist as mdist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
sim = simdata.make_config(dict(S=4, Egenes=3, C=300, mw=[0.5, 0.5], k=2, starts=7, uninform=0), seed=3)
D = simdata.host_data(sim)
init = simdata.init_phi(4, 2, 7, seed=9)
lls = {}
for rep in range(2):                       # two queues in a row: keys must not collide
    q = mdist.StartQueue(7)
    lls = {}
    while True:
        s = q()
        if s < 0:
            break
        lls[s] = float(orc.mnem_em(D, sim["pert"], 4, init[s:s + 1])["ll"][0])
        time.sleep(0.01 * (rank + 1))     # uneven ranks: the faster one takes more starts
    assert q() == -1
    assert sorted(lls) == sorted(q.taken)
    dist.barrier()
ids = sorted(lls)
winner, all_lls = mdist.gather_winner(ids, [lls[i] for i in ids], 7)
print(json.dumps(dict(rank=rank, my=ids, winner=int(winner), lls=[float(x) for x in all_lls])))
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_and_pick_winner_helpers():
    assert mdist.shard_starts(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((mdist.shard_starts(7, r, 3) for r in range(3)), [])) == list(range(7))
    assert mdist.pick_winner([1.0, 3.0, 3.0, 2.0]) == 2          # last maximal start
    assert mdist.pick_winner([-np.inf, -np.inf]) == 1


def test_two_rank_gloo_winner_matches_single_process():
    import json

    import oracle as orc
    from mnem_b200 import simdata
    port = _free_port()
    src = WORKER % dict(root=ROOT, port=port)
    procs = [subprocess.Popen([sys.executable, "-c", src, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=300)
        assert p.returncode == 0, e[-2000:]
        outs.append(json.loads(o.strip().splitlines()[-1]))
    sim = simdata.make_config(dict(S=4, Egenes=3, C=300, mw=[0.5, 0.5], k=2, starts=5, uninform=0), seed=3)
    D = simdata.host_data(sim)
    init = simdata.init_phi(4, 2, 5, seed=9)
    ref = orc.mnem_em(D, sim["pert"], 4, init)
    assert outs[0]["winner"] == outs[1]["winner"] == ref["best"]
    np.testing.assert_array_equal(outs[0]["lls"], ref["ll"])
    assert sorted(outs[0]["my"] + outs[1]["my"]) == list(range(5))


def test_two_rank_shared_start_queue_covers_every_start_once():
    import json

    import oracle as orc
    from mnem_b200 import simdata
    port = _free_port()
    src = WORKER_QUEUE % dict(root=ROOT, port=port)
    procs = [subprocess.Popen([sys.executable, "-c", src, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=300)
        assert p.returncode == 0, e[-2000:]
        outs.append(json.loads(o.strip().splitlines()[-1]))
    sim = simdata.make_config(dict(S=4, Egenes=3, C=300, mw=[0.5, 0.5], k=2, starts=7, uninform=0), seed=3)
    D = simdata.host_data(sim)
    init = simdata.init_phi(4, 2, 7, seed=9)
    ref = orc.mnem_em(D, sim["pert"], 4, init)
    assert sorted(outs[0]["my"] + outs[1]["my"]) == list(range(7))          # each start exactly once
    assert outs[0]["winner"] == outs[1]["winner"] == ref["best"]
    np.testing.assert_array_equal(outs[0]["lls"], ref["ll"])


def test_start_queue_without_process_group_is_a_local_counter():
    q = mdist.StartQueue(3)
    assert [q(), q(), q(), q(), q()] == [0, 1, 2, -1, -1]
    assert q.taken == [0, 1, 2]


# mnemk() with the (k, start) fits spread over the ranks: every k has its own queue, ranks walk the k's in rotated order
WORKER_MNEMK = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import oracle as orc
from mnem_b200 import simdata, dist as mdist
world = int(sys.argv[2])
if world > 1:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=world)
S, starts = 4, 3
sim = simdata.make_config(dict(S=S, Egenes=3, C=400, mw=[0.5, 0.5], k=2, starts=starts, uninform=0), seed=5)
D = simdata.host_data(sim)
class Cx:                                   # what mnemk_sharded reads from a Context
    S, C, E = S, sim["C"], sim["E"]
def fit_fn(k, init, q):                     # the CPU oracle stands in for the per-rank device fit
    ll = np.full(starts, -np.inf); adj = np.zeros((starts, k, S, S), dtype=np.uint8)
    theta = np.zeros((starts, k, sim["E"]), dtype=np.int32); mw = np.zeros((starts, k)); best = -1
    while True:
        s = q()
        if s < 0:
            break
        r = orc.mnem_em(D, sim["pert"], S, init[s:s + 1])
        ll[s], adj[s], theta[s], mw[s] = r["ll"][0], r["adj"][0], r["theta"][0], r["mw"][0]
        if best < 0 or ll[s] > ll[best] or (ll[s] == ll[best] and s > best):
            best = s
    return dict(ll=ll, adj=adj, theta=theta, mw=mw, best=best, probs=None)
def k1_fn():
    R = orc.do_mean(D, sim["pert"], S)
    f = orc.nem_greedy(R, None)
    g = orc.get_probs(D, sim["pert"], S, f["adj"][None], f["theta"][None], np.zeros((1, sim["C"])), np.ones(1))
    return dict(ll=g["ll"], comp=[dict(phi=f["adj"], theta=f["theta"])], probs=g["probs"])
def ic_fn(x, n):                            # getIC(degree = 4), R/mnems.r:1477-1529
    fpar = 0
    for c in x["comp"]:
        t = orc.mytc(c["phi"]).copy(); np.fill_diagonal(t, 0); fpar += int((t != 0).sum())
    k = len(x["comp"])
    fpar += k * len(x["comp"][0]["theta"]) + k - 1
    return np.log(n) * fpar - 2 * x["ll"] * np.log(2)
if world > 1:
    dist.barrier()                          # start together, like ranks that have just built their contexts
r = mdist.mnemk_sharded(Cx, [1, 2, 3], starts, seed=11, fit_fn=fit_fn, k1_fn=k1_fn, ic_fn=ic_fn)
print(json.dumps(dict(k=r["k"], ics={str(q): v for q, v in r["ics"].items()}, lls={str(q): v for q, v in r["lls"].items()},
                      owners={str(q): r["fits"][q]["owner"] for q in r["fits"]}, n_fitted=sum(len(v) for v in r["taken"].values()),
                      adj=[r["fits"][q]["adj"].astype(int).tolist() for q in (1, 2, 3)])))
if world > 1:
    dist.destroy_process_group()
"""


def test_two_rank_mnemk_picks_the_same_k_and_fits_as_one_process():
    import json
    port = _free_port()
    src = WORKER_MNEMK % dict(root=ROOT, port=port)

    def run(world):
        procs = [subprocess.Popen([sys.executable, "-c", src, str(r), str(world)], stdout=subprocess.PIPE,
                                  stderr=subprocess.PIPE, text=True) for r in range(world)]
        outs = []
        for p in procs:
            o, e = p.communicate(timeout=600)
            assert p.returncode == 0, e[-3000:]
            outs.append(json.loads(o.strip().splitlines()[-1]))
        return outs
    one = run(1)[0]
    two = run(2)
    for o in two:
        assert o["k"] == one["k"] and o["ics"] == one["ics"] and o["lls"] == one["lls"] and o["adj"] == one["adj"]
    assert two[0]["owners"] == two[1]["owners"]
    assert two[0]["n_fitted"] + two[1]["n_fitted"] == 1 + 3 + 3 and min(o["n_fitted"] for o in two) >= 1   # spread


# bootstrap(): the resamples dealt to the ranks, closure counts all-reduced
WORKER_BOOT = r"""
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import oracle as orc
from mnem_b200 import dist as mdist
world = int(sys.argv[2])
if world > 1:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=world)
rng = np.random.default_rng(5)
k, E, S = 2, 40, 5
R = rng.normal(size=(k, E, S)) * 2
def fit_fn(Rb):                              # the CPU oracle stands in for the batch of device searches
    return np.stack([orc.mytc(orc.nem_greedy(r, None)["adj"]) for r in Rb])
sup, idx = mdist.bootstrap_sharded(None, None, size=23, p=0.8, seed=4, R=R, fit_fn=fit_fn)
print(json.dumps(dict(sup=sup.tolist(), idx_sum=int(idx.sum()))))
if world > 1:
    dist.destroy_process_group()
"""


def test_two_rank_bootstrap_equals_one_process():
    import json
    port = _free_port()
    src = WORKER_BOOT % dict(root=ROOT, port=port)

    def run(world):
        procs = [subprocess.Popen([sys.executable, "-c", src, str(r), str(world)], stdout=subprocess.PIPE,
                                  stderr=subprocess.PIPE, text=True) for r in range(world)]
        outs = []
        for p in procs:
            o, e = p.communicate(timeout=600)
            assert p.returncode == 0, e[-3000:]
            outs.append(json.loads(o.strip().splitlines()[-1]))
        return outs
    one = run(1)[0]
    two = run(2)
    assert two[0] == two[1] == one
    sup = np.array(one["sup"])
    assert sup.shape == (2, 5, 5) and (np.diagonal(sup, axis1=1, axis2=2) == 1.0).all() and sup.min() >= 0 and sup.max() <= 1
