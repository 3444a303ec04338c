"""Multi-GPU plumbing: EM starts are independent until the final argmax (R/mnems.r:2166-2170, 2222-2230), so they
are sharded across ranks (one process per GPU) with no data-path collective.  The only exchange is the tiny
(ll, start id) reduction that picks the winner, done with torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
import numpy as np


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
