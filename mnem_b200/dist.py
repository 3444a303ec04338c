"""Multi-GPU plumbing: EM starts are independent until the final argmax (R/mnems.r:2166-2170, 2222-2230), so they
are sharded across ranks (one process per GPU) with no data-path collective.  The only exchange is the tiny
(ll, start id) reduction that picks the winner, done with torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
import numpy as np

_queue_serial = [0]


class StartQueue:
    """One queue of EM starts drained by all ranks: a ticket counter in the torch.distributed key-value store (the
    rendezvous store every process group already has; `add` is atomic on the store's server).  A rank asks for its
    next start whenever one of its slots frees up, so ranks whose starts converge quickly take more of them and all
    GPUs run dry together -- the role snowfall's worker pool plays in the reference (R/mnems.r:2166-2170).  This is
    control traffic (one small TCP round trip per start), not a data-path collective.

    Every rank must construct its queues in the same order (the key is a per-process serial number).  Without an
    initialised process group it degenerates to a local counter."""

    def __init__(self, n_starts, store=None):
        import torch.distributed as dist
        self.n = int(n_starts)
        self.taken = []
        _queue_serial[0] += 1
        self.key = "mnem_b200/start_queue/%d" % _queue_serial[0]
        self.store = store
        if store is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from torch.distributed import distributed_c10d
            self.store = distributed_c10d._get_default_store()
        self._local = 0

    def __call__(self):
        if self.store is not None:
            t = int(self.store.add(self.key, 1)) - 1
        else:
            t = self._local
            self._local += 1
        if t >= self.n:
            return -1
        self.taken.append(t)
        return t


def shard_starts(n_starts, rank, world):
    """Start s runs on rank s mod world (round-robin keeps the per-rank queues the same length +-1)."""
    return list(range(rank, n_starts, world))


def pick_winner(lls):
    """R/mnems.r:2222-2230: `max(best$ll) <= max(limits[[i]]$ll)` scanning upwards => the LAST maximal start wins."""
    lls = np.asarray(lls, dtype=np.float64)
    best = 0
    for i in range(1, len(lls)):
        if lls[best] <= lls[i]:
            best = i
    return best


def gather_winner(local_ids, local_lls, n_starts, device=None):
    """All ranks contribute (start id, ll) of their shard; every rank returns (winner start id, all lls).

    One all_reduce(SUM) of an n_starts vector whose untouched entries are 0 plus a mask: 16 bytes per start."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(2, n_starts, dtype=torch.float64, device=device)
    if len(local_ids):
        idx = torch.as_tensor(list(local_ids), dtype=torch.long, device=device)
        buf[0, idx] = torch.as_tensor(np.asarray(local_lls, dtype=np.float64), device=device)
        buf[1, idx] = 1.0
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    got = buf.cpu().numpy()
    if not (got[1] == 1.0).all():
        raise RuntimeError("starts were not covered exactly once across ranks")
    return pick_winner(got[0]), got[0]


def mnemk_sharded(cx, ks, starts, seed=0, phi=None, complete=False, logtype=2, batch=0, device=None, fit_fn=None, k1_fn=None,
                  ic_fn=None):
    """mnemk() (R/mnems.r:1560-1574: fit every k, keep the smallest information criterion) with the (k, start) fits spread
    over all ranks.  Every k has its own shared queue of starts (StartQueue); rank r walks the k's in an order rotated by
    r and drains whatever is left of each queue, so with few starts per k (the perturb-seq config has 3) different ranks
    work on different k's at the same time and no GPU waits for a whole k to finish.  Exchanges: one 16-byte-per-start
    reduction per k (gather_winner) and one all_gather_object of the per-k winners' graphs -- no data-path collective.

    cx: this rank's Context.  phi: optional {k: [starts, k, S, S]} initial graphs (default: sample_init(seed), the same on
    every rank).  fit_fn(k, init, next_start) / k1_fn() / ic_fn(fit, n_cells) replace the device calls in CPU tests.
    Returns dict(k, ics, lls, fits = {k: winner summary}, best = fits[k], taken = {k: starts this rank fitted}); `probs`
    of a winner is present only on the rank that fitted it (its `owner`)."""
    import torch.distributed as dist

    from . import api
    on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if on else 0
    world = dist.get_world_size() if on else 1
    ks = [int(k) for k in ks]
    queues = {k: StartQueue(starts if k > 1 else 1) for k in ks}            # constructed in the same order on every rank
    if fit_fn is None:
        def fit_fn(k, init, q):
            return api.em_run(cx, init, complete=complete, logtype=logtype, batch=batch, next_start=q, probs="winner")
    if k1_fn is None:
        def k1_fn():
            return api._mnem_k1(cx, np.zeros((cx.S, cx.S), dtype=np.uint8), complete, logtype)
    local = {}
    for i in range(len(ks)):
        k = ks[(i + rank) % len(ks)]
        q = queues[k]
        if k == 1:
            if q() == 0:
                f = k1_fn()
                local[k] = dict(my=[0], ll=np.array([f["ll"]]), adj=np.asarray(f["comp"][0]["phi"])[None, None],
                                theta=np.asarray(f["comp"][0]["theta"])[None, None], mw=np.ones((1, 1)), probs=f["probs"][None],
                                best=0)
            else:
                local[k] = dict(my=[], ll=np.zeros(0))
            continue
        init = np.asarray(phi[k]) if phi is not None else api.sample_init(cx.S, k, starts, seed)
        r = fit_fn(k, init, q)
        my = sorted(q.taken)
        local[k] = dict(my=my, ll=np.asarray(r["ll"])[my], adj=r["adj"], theta=r["theta"], mw=r["mw"], probs=r.get("probs"),
                        best=r["best"])
    fits = {}
    for k in ks:
        n = starts if k > 1 else 1
        L = local[k]
        win, lls = gather_winner(L["my"], L["ll"], n, device=device)
        mine = win in L["my"]
        summ = None
        if mine:
            p = L.get("probs")
            summ = dict(k=k, owner=rank, start=int(win), ll=float(lls[win]), adj=np.asarray(L["adj"][win]),
                        theta=np.asarray(L["theta"][win]), mw=np.asarray(L["mw"][win]), lls=[float(x) for x in lls])
            if p is not None and L["best"] == win:
                summ["probs"] = p[0] if len(p) == 1 else p[win]
        if on:
            box = [None] * world
            dist.all_gather_object(box, None if summ is None else {q_: v for q_, v in summ.items() if q_ != "probs"})
            got = next(b for b in box if b is not None)
            fits[k] = summ if mine else got
        else:
            fits[k] = summ
    ics, lls = {}, {}
    for k in ks:
        f = fits[k]
        comp = [dict(phi=f["adj"][i], theta=f["theta"][i]) for i in range(k)]
        ics[k] = float(ic_fn(dict(comp=comp, ll=f["ll"]), cx.C) if ic_fn else api.getIC(dict(comp=comp, ll=f["ll"]), cx.C, logtype=logtype))
        lls[k] = f["ll"]
    kb = min(ics, key=lambda q_: ics[q_])
    return dict(k=kb, ics=ics, lls=lls, fits=fits, best=fits[kb], taken={k: list(local[k]["my"]) for k in ks})


def bootstrap_sharded(x, cx, size=1000, p=1.0, logtype=2, complete=False, seed=0, device=None, R=None, fit_fn=None):
    """bootstrap() (R/mnems.r:2348-2369) with the `size` resamples dealt round-robin to the ranks: every rank draws the SAME
    seeded row samples, fits its share as one batch of device searches per component and the closure counts are summed
    by one all_reduce of k * S * S doubles.  Returns (support [k, S, S], idx) on every rank -- equal to the single-process
    result (the counts are small integers: their sum is exact in any order).
    R / fit_fn: collapsed data [k, E, S] and a replacement for the device searches (CPU tests)."""
    import torch
    import torch.distributed as dist

    from . import api
    on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    part = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    if R is None:
        sup, idx = api.bootstrap(x, cx, size=size, p=p, logtype=logtype, complete=complete, seed=seed, part=part)
    else:
        R = np.asarray(R)
        sup, idx = api._bootstrap_core(R, R.shape[0], R.shape[2], size, p, seed, 256, part, fit_fn)
    if on:
        t = torch.as_tensor(sup * size, dtype=torch.float64, device=device)      # integer counts
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        sup = t.cpu().numpy() / size
    return sup, idx
