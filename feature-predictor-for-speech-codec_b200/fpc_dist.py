"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the GPU box, gloo in the
CPU tests).  Only what the two paths need:

* closed-loop encode shards by utterance -- no collective on the data path; `shard_range`
  gives each rank its contiguous utterance range, `merge_histograms` sums the five cb_tot
  tables afterwards (off the timed path).
* k-means shards the residual vectors; `allreduce_kmeans` sums the per-centroid float64
  sums / counts (<= 147 KB) across ranks once per Lloyd iteration, which is the only exchange
  step of cb_func.update.

Every function degrades to the single-process case when torch.distributed is not initialised.
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed(group=None):
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def rank(group=None):
    dist = _dist()
    return dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0


def world_size(group=None):
    dist = _dist()
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def shard_range(n_units, rank_, world):
    """Contiguous, balanced range of units (utterances / residual vectors) for one rank:
    the first n % world ranks get one extra.  Returns (first, count)."""
    if world < 1 or not (0 <= rank_ < world):
        raise ValueError("bad rank %r / world %r" % (rank_, world))
    base, extra = divmod(int(n_units), world)
    first = rank_ * base + min(rank_, extra)
    return first, base + (1 if rank_ < extra else 0)


def allreduce_kmeans(sums, counts, n_local, group=None, want_total=True):
    """In-place SUM of the per-centroid accumulators over all ranks; returns the global number
    of vectors (cb_func.py:94 divides the counts by it).  `counts` may be None.
    want_total=False skips the device->host read of that number (returns None): the Lloyd loop
    does not need it on the host, fpc_kmeans_finalize derives it from the counts."""
    if not is_distributed(group):
        return int(n_local)
    import torch
    dist = _dist()
    flat = torch.empty(sums.numel() + (counts.numel() if counts is not None else 0) + 1, dtype=torch.float64,
                       device=sums.device)
    flat[:sums.numel()] = sums.reshape(-1)
    if counts is not None:
        flat[sums.numel():-1] = counts.reshape(-1)
    flat[-1] = float(n_local)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)   # one message: sums | counts | N
    sums.copy_(flat[:sums.numel()].reshape(sums.shape))
    if counts is not None:
        counts.copy_(flat[sums.numel():-1].reshape(counts.shape))
    if not want_total:
        return None
    return int(round(float(flat[-1].item())))


def chain_in_rank_order(step, carry, group=None):
    """Runs `step()` on rank 0, 1, ... in turn, handing the device tensor `carry` from each rank to all others before
    the next one starts: a reduction that must visit the shards in order (the float32 row-order sum behind
    np.mean(data, 0), cb_func.py:34).  Returns the global number of ranks visited."""
    if not is_distributed(group):
        step()
        return 1
    dist = _dist()
    me, n = dist.get_rank(group), dist.get_world_size(group)
    for r in range(n):
        if r == me:
            step()
        t = carry if dist.get_backend(group) == "nccl" else carry.cpu()
        dist.broadcast(t, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        if t is not carry:
            carry.copy_(t)
    return n


def broadcast_array(arr, group=None, src=0):
    """NumPy array from `src` to every rank (the LBG jitter, cb_func.py:41)."""
    if not is_distributed(group):
        return arr
    import torch
    dist = _dist()
    backend = dist.get_backend(group)
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if backend == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()


def merge_histograms(cb_tot, group=None):
    """Sums the five cb_tot tables of Wavernn.encoder over ranks (never-hit tables are the
    int 0 of wavernn.py:189, so sizes are exchanged first)."""
    if not is_distributed(group):
        return cb_tot
    import torch
    dist = _dist()
    backend = dist.get_backend(group)
    dev = "cuda" if backend == "nccl" else "cpu"
    sizes = torch.tensor([0 if np.isscalar(h) else len(h) for h in cb_tot], dtype=torch.int64, device=dev)
    dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=group)
    out = []
    for h, n in zip(cb_tot, sizes.tolist()):
        if n == 0:
            out.append(0)
            continue
        t = torch.zeros(n, dtype=torch.float64, device=dev)
        if not np.isscalar(h):
            t += torch.from_numpy(np.asarray(h, dtype=np.float64)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        out.append(t.cpu().numpy())
    return out
