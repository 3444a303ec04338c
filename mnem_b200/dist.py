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
