"""Drop-in for /root/reference/src/quantization/vq_func.py (stand-alone quantiser calls).

Same names, argument meaning and return values as the reference:

    vq_quantize(r (n,17) ndarray, cb_path)      -> (qr (n,17) in the codebook dtype, [hist per stage])   :134-164
    scl_quantize(data (n,1) ndarray, cb_path)   -> (q (n,1), hist (n_code,))                            :167-185
    quantize_mstage(x (17,), n_entries, CB)     -> (csum (17,), idx (stages,))                          :82-131

(vq_quantize_mbest, :10-24, is the inner loop of quantize_mstage; its survivor lists never leave
the chip here, so it has no host-callable twin.)

The search runs in the sm_100a kernels behind the C ABI (`fpc_vq_quantize_packed`,
`fpc_scl_quantize`): exact direct-form distances with numpy's roundings, 5 survivors, lowest
index on ties.  Codebook files are read and packed once per (path, mtime, size) instead of on
every call.  Inputs may be NumPy arrays (results come back as NumPy, like the reference) or
CUDA tensors (results stay on the device).  No CPU fallback.
"""
import hashlib

import numpy as np

import fpc_codebooks
import fpc_native as N

SURVIVORS = 5    # vq_func.py:3
NB_BANDS = 18    # vq_func.py:4

_array_images = {}


def _torch():
    import torch
    return torch


def _to_device(x, cols):
    torch = _torch()
    N.require_cuda()
    if isinstance(x, torch.Tensor):
        was_tensor = True
        t = x.detach()
        if not t.is_cuda:
            t = t.cuda()
    else:
        was_tensor = False
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32)).cuda()
    t = t.to(torch.float32).reshape(-1, cols).contiguous()
    return t, was_tensor


def _vq_run(img, x_dev):
    torch = _torch()
    a = img.arrays["vq"]
    stages, K = a.shape[0], a.shape[1]
    is32 = a.dtype == np.float32
    n = x_dev.shape[0]
    q = torch.empty((n, 17), dtype=torch.float32 if is32 else torch.float64, device=x_dev.device)
    idx = torch.empty((n, stages), dtype=torch.int32, device=x_dev.device)
    with torch.cuda.device(x_dev.device):
        N.check(N.lib().fpc_vq_quantize_packed(x_dev.data_ptr(), n, img.ptr(), 0, N.FPC_F32 if is32 else N.FPC_F64,
                                               stages, q.data_ptr(), idx.data_ptr(), N.current_stream(x_dev.device)),
                "fpc_vq_quantize_packed")
    return q, idx, stages, K


def vq_quantize(r, cb_path):
    """vq_func.py:134-164."""
    x, was_tensor = _to_device(r, 17)
    img = fpc_codebooks.single(cb_path, "vq", x.device)
    q, idx, stages, K = _vq_run(img, x)
    if was_tensor:
        hist = [_torch().bincount(idx[:, s].long(), minlength=K).to(_torch().float64) for s in range(stages)]
        return q, hist
    idx_h = idx.cpu().numpy()
    cb_tot = [np.bincount(idx_h[:, s], minlength=K).astype(np.float64) for s in range(stages)]
    return q.cpu().numpy(), cb_tot


def vq_quantize_indices(r, cb_path):
    """Like vq_quantize but returns (qr, idx (n,stages)) -- the per-vector indices the reference
    only exposes through its histograms."""
    x, was_tensor = _to_device(r, 17)
    img = fpc_codebooks.single(cb_path, "vq", x.device)
    q, idx, _, _ = _vq_run(img, x)
    return (q, idx) if was_tensor else (q.cpu().numpy(), idx.cpu().numpy())


def scl_quantize(data, cb_path):
    """vq_func.py:167-185."""
    torch = _torch()
    x, was_tensor = _to_device(data, 1)
    img = fpc_codebooks.single(cb_path, "scl", x.device)
    codes = img.arrays["scl"].reshape(-1)
    is32 = codes.dtype == np.float32
    n = x.shape[0]
    q = torch.empty((n,), dtype=torch.float32 if is32 else torch.float64, device=x.device)
    idx = torch.empty((n,), dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().fpc_scl_quantize(x.data_ptr(), n, img.scl_device_ptr(), N.FPC_F32 if is32 else N.FPC_F64,
                                         len(codes), q.data_ptr(), idx.data_ptr(), N.current_stream(x.device)),
                "fpc_scl_quantize")
    if was_tensor:
        return q[:, None], torch.bincount(idx.long(), minlength=len(codes)).to(torch.float64)
    cb_tot = np.bincount(idx.cpu().numpy(), minlength=len(codes)).astype(np.float64)
    return q.cpu().numpy()[:, None], cb_tot


def _image_for_array(CB):
    CB = np.asarray(CB)
    if CB.dtype not in (np.float32, np.float64):
        CB = CB.astype(np.float64)
    CB = np.ascontiguousarray(CB)
    key = (CB.shape, CB.dtype.str, hashlib.blake2b(CB.tobytes(), digest_size=16).digest())
    img = _array_images.get(key)
    if img is None:
        if len(_array_images) > 16:
            _array_images.clear()
        img = fpc_codebooks.PackedCodebooks(vq=CB)
        _array_images[key] = img
    return img


def quantize_mstage(x, n_entries, CEPS_CODEBOOK):
    """vq_func.py:82-131 for one vector: (csum, index[:, 0])."""
    CB = np.asarray(CEPS_CODEBOOK)
    if CB.ndim == 2:
        CB = CB[None]
    n_entries = [int(v) for v in np.atleast_1d(n_entries)]
    if any(k != CB.shape[1] for k in n_entries[:CB.shape[0]]):
        CB = CB[:, :n_entries[0], :]
    img = _image_for_array(CB[:len(n_entries)])
    xd, _ = _to_device(np.asarray(x)[None], 17)
    q, idx, _, _ = _vq_run(img, xd)
    return q.cpu().numpy()[0], idx.cpu().numpy()[0].astype(np.int64)
