"""Codebook-training data path on the device.

Mirrors the body of the reference's train_cb.py main loop (/root/reference/src/train_cb.py:141-211)
without the host round trips:

    res = model.encode_device(cfg, feat, None, l1, l2, qtz=False)      # :167-168, the documented flow
    sets = training_sets(res.r, res.r_under)                            # :177-187
    codebooks = train_stages(sets["vq_above"], n_entries=[1024, 1024])  # :191-211
    scl = scalar_codebook(sets["scl_above"], 256)                       # :219-226 (commented out there)

`training_sets` keeps the rows in order and on the device (C ABI fpc_compact_rows), so the k-means
kernels consume the residuals where the encoder left them; `train_stages` forms the next stage's
vectors as `quantize(cb, r) - r` (the reference's sign, :200) with fpc_kmeans_stage_residual.
The next stage's data is float64 as in the reference (float64 codebook - float32 residual, :200): the k-means entry
points take float64 vectors (fpc_kmeans_assign_accumulate_f64, fpc_kmeans_colsum_f64) and use them for the exact
distances, the sums and the seed mean.
The scalar learner is a plain 1-D Lloyd iteration; parity with the (unpinned, commented-out)
sklearn KMeans of the reference is not claimed.
"""
import numpy as np
import torch

import fpc_native as N
from quantization import cb_func


def compact_rows(src, col0, ncols):
    """Rows of the 2-D float32 CUDA tensor `src` whose columns [col0, col0+ncols) are not all zero
    (`sum(abs(row)) != 0`, train_cb.py:187), those columns only, original order.  Returns a CUDA tensor (m, ncols)."""
    N.require_cuda()
    if not (isinstance(src, torch.Tensor) and src.is_cuda):
        raise N.FpcError("compact_rows needs a CUDA tensor: no CPU fallback")
    t = src.detach().to(torch.float32)
    t = t.reshape(-1, t.shape[-1]).contiguous()
    n, stride = t.shape
    dev = t.device
    dst = torch.empty((n, ncols), dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        need = N.lib().fpc_compact_workspace_bytes(n)
        ws = torch.empty(max(need, 4), dtype=torch.uint8, device=dev)
        N.check(N.lib().fpc_compact_rows(t.data_ptr(), n, stride, col0, ncols, dst.data_ptr(), count.data_ptr(),
                                         ws.data_ptr(), ws.numel(), N.current_stream(dev)), "fpc_compact_rows")
    m = int(count.item())
    return dst[:m]


def training_sets(r, r_under, code_dims=17):
    """train_cb.py:177-187.  r / r_under: (B, L, 18) outputs of encoder(qtz=False) (above- / below-threshold masked
    residuals).  Returns device tensors: scl_above (m,), scl_below (m,), vq_above (m, code_dims), vq_below (m, code_dims)."""
    first = 18 - code_dims
    return {
        "scl_above": compact_rows(r, 0, 1)[:, 0],
        "scl_below": compact_rows(r_under, 0, 1)[:, 0],
        "vq_above": compact_rows(r, first, code_dims),
        "vq_below": compact_rows(r_under, first, code_dims),
    }


def stage_residual(codebook, data, keep_float64=True):
    """r = quantize(codebook, r) - r on the device (train_cb.py:199-200).  data: CUDA (N,17) float32 or float64.
    The result is float64 as in the reference; keep_float64=False rounds it to float32 (half the memory, and the next
    stage then runs on the tensor-core assignment -- a stated deviation from :200)."""
    cb = cb_func._cb_on_device(codebook, data.device)
    _, _, idx = cb_func.assign_accumulate(data, cb, want_idx=True, want_sums=False)
    with torch.cuda.device(data.device):
        if keep_float64:
            nxt = torch.empty(data.shape, dtype=torch.float64, device=data.device)
            N.check(N.lib().fpc_kmeans_stage_residual_f64(cb.data_ptr(), cb.shape[0], idx.data_ptr(), data.data_ptr(),
                                                          1 if data.dtype == torch.float64 else 0, data.shape[0], nxt.data_ptr(),
                                                          N.current_stream(data.device)), "fpc_kmeans_stage_residual_f64")
        else:
            d32 = data.to(torch.float32)
            nxt = torch.empty_like(d32)
            N.check(N.lib().fpc_kmeans_stage_residual(cb.data_ptr(), cb.shape[0], idx.data_ptr(), d32.data_ptr(), d32.shape[0],
                                                      nxt.data_ptr(), N.current_stream(data.device)), "fpc_kmeans_stage_residual")
        torch.cuda.current_stream(data.device).synchronize()
    return nxt


def train_stages(data, n_entries, codebooks=None, first_batch=True, group=None, rng=None, keep_float64=True, ordered=None):
    """train_cb.py:189-211 for one batch of residual vectors: for every stage, `vq_train` (first batch) or ten
    `update` calls (later batches), then the next stage trains on quantize(cb, r) - r.
    ordered=True: sums in data order (cb_func.update_device), the reference's codebooks bit for bit on one GPU.
    Returns the list of (K_i, 17) float64 codebooks."""
    d = cb_func._data_on_device(data)
    out = []
    for i, K in enumerate(n_entries):
        cb0 = np.zeros((K, 17)) if codebooks is None else np.asarray(codebooks[i], dtype=np.float64)
        if first_batch:
            cb = cb_func.vq_train(d, cb0, K, group=group, rng=rng, ordered=ordered)
        else:
            cb = cb0
            for _ in range(10):
                cb = cb_func.update(d, cb, K, group=group, verbose=False, ordered=ordered)
        out.append(cb)
        if i + 1 < len(n_entries):
            d = stage_residual(cb, d, keep_float64)
    return out


def scalar_codebook(values, n_levels, iters=50):
    """1-D Lloyd codebook for the c0 residuals (the role of the commented-out sklearn KMeans, train_cb.py:219-226).
    values: 1-D CUDA tensor.  Initialised at equal-count quantiles of the sorted data; every iteration moves the
    decision boundaries to the midpoints and the levels to the cell means (sorted data + prefix sums).  Returns an
    (n_levels, 1) float64 NumPy array, ascending, in the layout scl_quantize reads."""
    if not (isinstance(values, torch.Tensor) and values.is_cuda):
        raise N.FpcError("scalar_codebook needs a CUDA tensor: no CPU fallback")
    x = values.detach().reshape(-1).to(torch.float64)
    if x.numel() == 0:
        raise ValueError("no training values")
    x, _ = torch.sort(x)
    n = x.numel()
    csum = torch.cat([torch.zeros(1, dtype=torch.float64, device=x.device), torch.cumsum(x, 0)])
    q = (torch.arange(n_levels, device=x.device, dtype=torch.float64) + 0.5) / n_levels
    levels = x[(q * (n - 1)).round().long()].clone()
    for _ in range(iters):
        bounds = (levels[1:] + levels[:-1]) / 2
        cut = torch.searchsorted(x, bounds)
        lo = torch.cat([torch.zeros(1, dtype=torch.long, device=x.device), cut])
        hi = torch.cat([cut, torch.full((1,), n, dtype=torch.long, device=x.device)])
        cnt = (hi - lo).to(torch.float64)
        mean = (csum[hi] - csum[lo]) / torch.clamp(cnt, min=1.0)
        new = torch.where(cnt > 0, mean, levels)          # an empty cell keeps its level
        if torch.equal(new, levels):
            break
        levels = new
    return torch.sort(levels)[0].cpu().numpy().reshape(-1, 1)
