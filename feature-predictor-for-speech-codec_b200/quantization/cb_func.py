"""Drop-in for /root/reference/src/quantization/cb_func.py (k-means / LBG codebook learning).

    vq_train(data, codebook, nb_entries)   :28-54    grow-by-one LBG + Lloyd
    find_nearest(data, codebook)           :56-68    float64 direct-form argmin (first minimum)
    update(data, codebook, nb_entries_tmp) :71-100   one Lloyd iteration (= BASELINE's "k-means iter")
    quantize(codebook, data)               :103-112  nearest-centroid gather

`data` may be a NumPy array (copied to the device once per call) or a CUDA float32 tensor
(N,17) that stays resident -- vq_train / train_cb.py call `update` thousands of times on the
same residuals, so callers that care pass the tensor.  Codebooks are float64 NumPy arrays like
the reference's.

Several GPUs: when torch.distributed is initialised (one process per GPU), `data` is THIS
rank's shard of the residual vectors; per-centroid float64 sums and counts are all-reduced
(NCCL over NVLink) between the assign kernel and the divide, so every rank ends an iteration
with the same codebook.  The jitter of vq_train comes from NumPy's global RNG exactly as in the
reference (:41); with several ranks rank 0's draw is broadcast.

`ordered=True` (update / update_device / vq_train; default from the environment variable FPC_KMEANS_ORDERED=1): the
per-centroid sums are taken in DATA ORDER like the reference's accumulation loop (:82-86) by
fpc_kmeans_accumulate_ordered, so that on one GPU `update` returns the reference's codebook BIT FOR BIT and the same
bits on every run.  The default sums with float64 atomics inside the assign kernel: faster (one pass instead of five),
equal to ~1e-16 relative per sum, not bit-reproducible.  With several ranks the ordered sums of the shards are still
combined by the all-reduce, i.e. reproducible for a given rank count but not the one-GPU bits.
"""
import os

import numpy as np

import fpc_dist
import fpc_native as N


def _torch():
    import torch
    return torch


def _data_on_device(data):
    """(nb_vectors, 17) training vectors on the device.  float64 input stays float64 -- the data of a later training
    stage is `quantize(cb, r) - r` in float64 (train_cb.py:200) and NumPy then computes distances, sums and the seed
    mean in float64; everything else becomes float32, what the encoder's residuals are."""
    torch = _torch()
    N.require_cuda()
    if isinstance(data, torch.Tensor):
        t = data.detach()
        if not t.is_cuda:
            t = t.cuda()
        if t.dtype != torch.float64:
            t = t.to(torch.float32)
    else:
        a = np.asarray(data)
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64 if a.dtype == np.float64 else np.float32)).cuda()
    if t.dim() != 2 or t.shape[1] != 17:
        raise ValueError("data must be (nb_vectors, 17), got %r" % (tuple(t.shape),))
    return t.contiguous()


def _cb_on_device(codebook, dev):
    torch = _torch()
    cb = np.ascontiguousarray(np.asarray(codebook), dtype=np.float64)
    if cb.ndim != 2 or cb.shape[1] != 17:
        raise ValueError("codebook must be (nb_entries, 17), got %r" % (cb.shape,))
    return torch.from_numpy(cb).to(dev)


def _assign_fn(data_dev):
    """The C entry point for the dtype of the vectors (float64: later training stages)."""
    return N.lib().fpc_kmeans_assign_accumulate_f64 if data_dev.dtype == _torch().float64 else N.lib().fpc_kmeans_assign_accumulate


def assign_accumulate(data_dev, cb_dev, want_idx=False, want_sums=True):
    """One pass of fpc_kmeans_assign_accumulate over a device shard.
    Returns (sums (K,17) f64, counts (K,) f64, idx (N,) int32 or None), all on the device."""
    torch = _torch()
    dev = data_dev.device
    K = cb_dev.shape[0]
    n = data_dev.shape[0]
    sums = torch.zeros((K, 17), dtype=torch.float64, device=dev) if want_sums else None
    counts = torch.zeros((K,), dtype=torch.float64, device=dev) if want_sums else None
    idx = torch.empty((n,), dtype=torch.int32, device=dev) if want_idx else None
    with torch.cuda.device(dev):
        # scratch: replicated accumulation tables of small codebooks + the operand image of the tensor-core screen
        need = N.lib().fpc_kmeans_workspace_bytes(n, K)
        ws = torch.empty(need, dtype=torch.uint8, device=dev) if need else None
        N.check(_assign_fn(data_dev)(
            data_dev.data_ptr(), n, cb_dev.data_ptr(), K,
            sums.data_ptr() if want_sums else None, counts.data_ptr() if want_sums else None,
            idx.data_ptr() if want_idx else None, ws.data_ptr() if ws is not None else None, need,
            N.current_stream(dev)), "fpc_kmeans_assign_accumulate")
    return sums, counts, idx


_acc_cache = {}     # (device, K) -> [accumulator, workspace]: zeroed once, re-zeroed by every finalize


def _accumulator(dev, K, n):
    torch = _torch()
    hit = _acc_cache.get((str(dev), K))
    need = N.lib().fpc_kmeans_workspace_bytes(n, K)
    if hit is None or hit[1].numel() < need:
        if len(_acc_cache) > 64:
            _acc_cache.clear()
        hit = [torch.zeros(K * 18, dtype=torch.float64, device=dev), torch.empty(max(need, 1), dtype=torch.uint8, device=dev)]
        _acc_cache[(str(dev), K)] = hit
    return hit


def _ordered_default():
    return os.environ.get("FPC_KMEANS_ORDERED", "0") not in ("", "0")


_ord_cache = {}     # device -> [row indices (int32), workspace (uint8)] of the ordered accumulation


def _ordered_scratch(dev, n, K):
    torch = _torch()
    need = N.lib().fpc_kmeans_ordered_workspace_bytes(n, K)
    if need == 0 and n > 0:
        raise ValueError("ordered accumulation needs K <= 2048 and fewer than 2^31 vectors (K = %d, N = %d)" % (K, n))
    hit = _ord_cache.get(str(dev))
    if hit is None or hit[0].numel() < n or hit[1].numel() < need:
        hit = [torch.empty((max(n, 1),), dtype=torch.int32, device=dev), torch.empty(max(need, 1), dtype=torch.uint8, device=dev)]
        _ord_cache[str(dev)] = hit
    return hit


def update_device(data_dev, cb_dev, group=None, ordered=None):
    """One Lloyd iteration entirely on the device (+ the all-reduce when distributed), no host synchronisation.
    Returns (new codebook (K,17) f64 device tensor, stats (5,) f64 device tensor = min count, max count, #empty,
    sum (count/N)^2, N; the global vector count N as a 0-d device tensor view of stats[4]).

    Per iteration: the assign kernel(s) accumulate into ONE buffer [sums (K,17) | counts (K)], the ranks all-reduce that
    buffer in place (one message, no pack / unpack copies), and fpc_kmeans_finalize_acc divides and leaves the buffer
    zeroed for the next iteration -- no memset and no allocation inside the loop."""
    torch = _torch()
    dev = data_dev.device
    K = cb_dev.shape[0]
    n = data_dev.shape[0]
    cb_dev = cb_dev.contiguous()
    acc, ws = _accumulator(dev, K, n)
    out = torch.empty((K, 17), dtype=torch.float64, device=dev)
    stats = torch.empty((5,), dtype=torch.float64, device=dev)
    if ordered is None:
        ordered = _ordered_default()
    try:
        with torch.cuda.device(dev):
            if ordered:
                # indices only, then the sums in data order (cb_func.py:82-86 to the bit)
                idx, ows = _ordered_scratch(dev, n, K)
                N.check(_assign_fn(data_dev)(
                    data_dev.data_ptr(), n, cb_dev.data_ptr(), K, None, None, idx.data_ptr(),
                    ws.data_ptr(), ws.numel(), N.current_stream(dev)), "fpc_kmeans_assign_accumulate")
                N.check(N.lib().fpc_kmeans_accumulate_ordered(
                    data_dev.data_ptr(), int(data_dev.dtype == torch.float64), n, idx.data_ptr(), K, acc.data_ptr(),
                    acc.data_ptr() + K * 17 * 8, ows.data_ptr(), ows.numel(), N.current_stream(dev)),
                    "fpc_kmeans_accumulate_ordered")
            else:
                N.check(_assign_fn(data_dev)(
                    data_dev.data_ptr(), n, cb_dev.data_ptr(), K, acc.data_ptr(), acc.data_ptr() + K * 17 * 8, None,
                    ws.data_ptr(), ws.numel(), N.current_stream(dev)), "fpc_kmeans_assign_accumulate")
            if fpc_dist.is_distributed(group):
                fpc_dist._dist().all_reduce(acc, op=fpc_dist._dist().ReduceOp.SUM, group=group)
            # n_total = 0: nb_vectors is the sum of the (all-reduced) counts, taken on the device
            N.check(N.lib().fpc_kmeans_finalize_acc(acc.data_ptr(), K, 0.0, out.data_ptr(), stats.data_ptr(),
                                                    N.current_stream(dev)), "fpc_kmeans_finalize_acc")
    except Exception:
        acc.zero_()         # never leave a half-filled accumulator behind
        raise
    return out, stats, stats[4]


def find_nearest(data, codebook):
    """cb_func.py:56-68 -> (nb_vectors,) int64 (a CUDA tensor if `data` was one)."""
    torch = _torch()
    d = _data_on_device(data)
    cb = _cb_on_device(codebook, d.device)
    _, _, idx = assign_accumulate(d, cb, want_idx=True, want_sums=False)
    if isinstance(data, torch.Tensor):
        return idx.long()
    return idx.cpu().numpy().astype(np.int64)


def update(data, codebook, nb_entries_tmp, group=None, verbose=True, ordered=None):
    """cb_func.py:71-100.  Prints the same statistics line as the reference (:96-97)."""
    d = _data_on_device(data)
    cb = _cb_on_device(np.asarray(codebook)[:nb_entries_tmp], d.device)
    out, stats, _ = update_device(d, cb, group, ordered)
    s = stats.cpu().numpy()
    if verbose and fpc_dist.rank(group) == 0:
        print('{} - min: {}, max: {}, small: {}, error: {}'.format(
            nb_entries_tmp, np.array([s[0]]), np.array([s[1]]), np.array([int(s[2])]), s[3]))
    return out.cpu().numpy()


def quantize(codebook, data):
    """cb_func.py:103-112 -> (nb_vectors, 17) float64."""
    torch = _torch()
    d = _data_on_device(data)
    cb = _cb_on_device(codebook, d.device)
    _, _, idx = assign_accumulate(d, cb, want_idx=True, want_sums=False)
    q = torch.empty((d.shape[0], 17), dtype=torch.float64, device=d.device)
    with torch.cuda.device(d.device):
        N.check(N.lib().fpc_kmeans_gather(cb.data_ptr(), cb.shape[0], idx.data_ptr(), d.shape[0], q.data_ptr(),
                                          N.current_stream(d.device)), "fpc_kmeans_gather")
    if isinstance(data, torch.Tensor):
        return q
    return q.cpu().numpy()


def vq_train(data, codebook, nb_entries, group=None, verbose=False, rng=None, ordered=None):
    """cb_func.py:28-54.  The residuals go to the device once and the codebook stays there for the whole schedule:
    each of the 4(K-1)+10 Lloyd iterations is one assign kernel, one (optional) all-reduce and one finalize kernel,
    and the grow-by-one step (copy entry 0, add the jitter) is two small device operations, so the host never waits
    for the GPU inside the loop.  The jitter is the reference's: `.001 * np.random.rand(e, ndims) / 2` drawn for
    e = 1, 2, ... from NumPy's global RNG (:41) -- drawn here in one call, which yields the same numbers because
    consecutive `rand` calls continue one stream."""
    torch = _torch()
    d = _data_on_device(data)
    dev = d.device
    ndims = d.shape[1]
    codebook = np.array(codebook, dtype=np.float64, copy=True)
    draw = (rng.rand if rng is not None else np.random.rand)
    # codebook[0] = np.mean(data, 0)  (:33).  The first stage's training set is float32 (train_cb.py:182-187), a later
    # one float64 (:200), and NumPy adds the rows up one after the other and divides IN THAT dtype: fpc_kmeans_colsum_f32 /
    # _f64 perform exactly those additions (serially, one CTA; ranks continue one another's sums in rank order), the
    # division is NumPy's own.
    n_box = torch.tensor([float(d.shape[0])], dtype=torch.float64, device=dev)
    n_total = fpc_dist.allreduce_kmeans(n_box, None, d.shape[0], group)
    carry = torch.zeros(17, dtype=d.dtype, device=dev)
    colsum = N.lib().fpc_kmeans_colsum_f64 if d.dtype == torch.float64 else N.lib().fpc_kmeans_colsum_f32

    def _my_rows():
        with torch.cuda.device(dev):
            N.check(colsum(d.data_ptr(), d.shape[0], carry.data_ptr(), N.current_stream(dev)), "fpc_kmeans_colsum")
    fpc_dist.chain_in_rank_order(_my_rows, carry, group)
    mean0 = np.true_divide(carry.cpu().numpy(), int(n_total))          # in the data's dtype, as in np.mean
    cb_full = torch.from_numpy(np.ascontiguousarray(codebook[:nb_entries])).to(dev)
    cb_full[0] = torch.from_numpy(mean0.astype(np.float64)).to(dev)
    n_draws = ndims * (nb_entries - 1) * nb_entries // 2
    jitter = None
    if n_draws > 0:
        flat = fpc_dist.broadcast_array(.001 * (draw(n_draws) / 2), group)      # rank 0's draw when distributed
        jitter = torch.from_numpy(np.ascontiguousarray(flat)).to(dev)
    e, off = 1, 0
    while e < nb_entries:
        cb_full[e] = cb_full[0]
        cb_full[:e] += jitter[off:off + e * ndims].view(e, ndims)
        off += e * ndims
        e += 1
        for _ in range(4):
            cb, stats, _ = update_device(d, cb_full[:e], group, ordered)
            cb_full[:e] = cb
        if verbose and fpc_dist.rank(group) == 0:
            s = stats.cpu().numpy()
            print('{} - min: {}, max: {}, small: {}, error: {}'.format(e, s[0], s[1], int(s[2]), s[3]))
    cb = cb_full[:nb_entries]
    for _ in range(10):
        cb, stats, _ = update_device(d, cb, group, ordered)
    return cb.cpu().numpy()
