"""ctypes front-end of the CPU oracle (oracle/fpc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does (tests/test_no_oracle_in_product.py
greps for it).

Every function cites the reference lines its C counterpart restates; see the header of
fpc_oracle.c for the arithmetic contract and DESIGN.md section 3 for how the oracle is
pinned against the real reference (tests/golden/*.npz, produced by oracle/gen_golden.py).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_WEIGHT_KEYS = (
    "rnn1.weight_ih_l0", "rnn1.weight_hh_l0", "rnn1.bias_ih_l0", "rnn1.bias_hh_l0",
    "rnn2.weight_ih_l0", "rnn2.weight_hh_l0", "rnn2.bias_ih_l0", "rnn2.bias_hh_l0",
    "dual_fc.0.weight", "dual_fc.0.bias",
)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libfpc_oracle.so")
        if not os.path.exists(path):
            import importlib.util
            spec = importlib.util.spec_from_file_location("_fpc_oracle_build", os.path.join(_HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            path = mod.build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_expf.restype = ctypes.c_float
        _LIB.orc_expf.argtypes = [ctypes.c_float]
        _LIB.orc_sigmoidf.restype = ctypes.c_float
        _LIB.orc_sigmoidf.argtypes = [ctypes.c_float]
        _LIB.orc_tanhf.restype = ctypes.c_float
        _LIB.orc_tanhf.argtypes = [ctypes.c_float]
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _train_data(a):
    """Training vectors keep float64 when they are float64 (a later stage's `quantize(cb, r) - r`, train_cb.py:200);
    everything else is float32, what the encoder's residuals are.  Returns (array, suffix of the C entry points)."""
    a = np.asarray(a)
    if a.dtype == np.float64:
        return np.ascontiguousarray(a), "_d"
    return _f32(a), ""


def weights_from_state_dict(sd):
    """state_dict (torch tensors or ndarrays, keys as wavernn.py:37-38,48-52) -> dict of f32 arrays."""
    out = {}
    for k in _WEIGHT_KEYS:
        v = sd[k]
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        out[k] = _f32(v)
    return out


def _weight_args(w):
    in_f = w["rnn1.weight_ih_l0"].shape[1]
    h1 = w["rnn1.weight_hh_l0"].shape[1]
    h2 = w["rnn2.weight_hh_l0"].shape[1]
    fc = w["dual_fc.0.weight"].shape[0]
    assert w["rnn1.weight_ih_l0"].shape[0] == 3 * h1 and w["rnn2.weight_ih_l0"].shape == (3 * h2, h1)
    return [ctypes.c_int(in_f), ctypes.c_int(h1), ctypes.c_int(h2), ctypes.c_int(fc)] + \
        [_p(w[k]) for k in _WEIGHT_KEYS], (in_f, h1, h2, fc)


class Codebooks:
    """The four codebook files the encoder reads (wavernn.py:219-237), as arrays.

    vq:     (stages, K, 17)  cfg['cb_path']          float32 or float64
    bl_vq:  (stages, K, 17)  cfg['bl_cb_path']       or None for ''
    scl:    (n, 1)           cfg['scl_cb_path']
    bl_scl: (n, 1)           cfg['bl_scl_cb_path']   or None for ''
    """

    def __init__(self, vq, scl, bl_vq=None, bl_scl=None):
        self.vq = self._prep_vq(vq)
        self.bl_vq = self._prep_vq(bl_vq)
        self.scl = self._prep_scl(scl)
        self.bl_scl = self._prep_scl(bl_scl)

    @staticmethod
    def _prep_vq(a):
        if a is None:
            return None
        a = np.asarray(a)
        if a.ndim != 3:
            # vq_func.py:143-146 computes n_entries before expand_dims, so 2-D files crash
            raise IndexError("VQ codebook must be 3-D (stages, K, ndim)")
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        return np.ascontiguousarray(a)

    @staticmethod
    def _prep_scl(a):
        if a is None:
            return None
        a = np.asarray(a)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        return np.ascontiguousarray(a.reshape(-1))

    def hist_sizes(self):
        return [len(self.scl), 0 if self.bl_scl is None else len(self.bl_scl),
                self.vq.shape[1], self.vq.shape[1] if self.vq.shape[0] > 1 else 0,
                0 if self.bl_vq is None else self.bl_vq.shape[1]]


def _vq_args(a):
    if a is None:
        return [ctypes.c_int(0), ctypes.c_int(0), None, None], None
    K = (ctypes.c_int * a.shape[0])(*([a.shape[1]] * a.shape[0]))
    return [ctypes.c_int(0 if a.dtype == np.float32 else 1), ctypes.c_int(a.shape[0]), K, _p(a)], K


def _scl_args(a):
    if a is None:
        return [ctypes.c_int(0), ctypes.c_int(0), None]
    return [ctypes.c_int(0 if a.dtype == np.float32 else 1), ctypes.c_int(len(a)), _p(a)]


def encode(weights, cbs, feat, l1, l2, mask=None, qtz=True, nthreads=0):
    """Wavernn.encoder (wavernn.py:165-256).  Returns a dict with the reference's outputs plus
    the per-frame indices idx (B,L,4) = [scalar idx, vq stage-1 idx (or below-cb idx),
    vq stage-2 idx, flags(bit0 ind1, bit1 ind2)], -1 where nothing was coded."""
    feat = _f32(feat)
    B, L, C = feat.shape
    wargs, (in_f, h1, h2, fc) = _weight_args(weights)
    assert C == in_f
    out = {
        "c_in": np.zeros((B, L, C), np.float32), "r": np.zeros((B, L, fc), np.float32),
        "r_qtz": np.zeros((B, L, fc), np.float32), "r_under": np.zeros((B, L, fc), np.float32),
        "ind1": np.zeros((B, L, 1), np.float32), "ind2": np.zeros((B, L, 1), np.float32),
        "idx": np.zeros((B, L, 4), np.int32),
    }
    m = None if mask is None else _f32(mask)
    a_vq, k1 = _vq_args(cbs.vq if cbs is not None else None)
    a_bl, k2 = _vq_args(cbs.bl_vq if cbs is not None else None)
    a_s = _scl_args(cbs.scl if cbs is not None else None)
    a_bs = _scl_args(cbs.bl_scl if cbs is not None else None)
    rc = lib().orc_encode(*wargs, *a_vq, *a_bl, *a_s, *a_bs, _p(feat), ctypes.c_int(B), ctypes.c_int(L),
                          ctypes.c_float(l1), ctypes.c_float(l2), _p(m), ctypes.c_int(1 if qtz else 0),
                          _p(out["c_in"]), _p(out["r"]), _p(out["r_qtz"]), _p(out["r_under"]),
                          _p(out["ind1"]), _p(out["ind2"]), _p(out["idx"]), ctypes.c_int(nthreads))
    if rc:
        raise RuntimeError("orc_encode failed with status %d" % rc)
    return out


def decode(weights, r_qtz, pitch):
    """Receiver-side replay c[t] = f(c[t-1]) + r_qtz[t]."""
    r_qtz = _f32(r_qtz)
    pitch = _f32(pitch)
    B, L, fc = r_qtz.shape
    wargs, (in_f, h1, h2, fc2) = _weight_args(weights)
    out = np.zeros((B, L, in_f), np.float32)
    rc = lib().orc_decode(*wargs, _p(r_qtz), _p(pitch), ctypes.c_int(B), ctypes.c_int(L), _p(out))
    if rc:
        raise RuntimeError("orc_decode failed with status %d" % rc)
    return out


def forward(weights, x, h1=None, h2=None):
    """Wavernn.forward (wavernn.py:63-102) over (B,T,in).  Returns y (B,T,fc), h1, h2."""
    x = _f32(x)
    B, T, C = x.shape
    wargs, (in_f, h1n, h2n, fc) = _weight_args(weights)
    h1 = np.zeros((B, h1n), np.float32) if h1 is None else _f32(h1).reshape(B, h1n).copy()
    h2 = np.zeros((B, h2n), np.float32) if h2 is None else _f32(h2).reshape(B, h2n).copy()
    y = np.zeros((B, T, fc), np.float32)
    lib().orc_forward(*wargs, _p(x), ctypes.c_int(B), ctypes.c_int(T), _p(y), _p(h1), _p(h2))
    return y, h1, h2


def histograms(idx, cbs):
    """cb_tot of wavernn.py:189,221-240 from the index record: 5 float64 count tables
    [scl_above, scl_below, vq_stage1, vq_stage2, vq_below]; never-hit tables are the int 0
    the reference starts from."""
    idx = np.asarray(idx).reshape(-1, 4)
    sizes = cbs.hist_sizes()
    above1 = (idx[:, 3] & 1) != 0
    above2 = (idx[:, 3] & 2) != 0
    # below threshold the LAST stage of the book is counted: cb_tot[4] += cb_t[-1] (wavernn.py:240)
    below_last = np.where(idx[~above2, 2] >= 0, idx[~above2, 2], idx[~above2, 1])
    sel = [(idx[above1, 0], sizes[0]), (idx[~above1, 0], sizes[1]), (idx[above2, 1], sizes[2]),
           (idx[above2, 2], sizes[3]), (below_last, sizes[4])]
    out = []
    for v, n in sel:
        v = v[v >= 0]
        out.append(np.bincount(v, minlength=n).astype(np.float64) if (n > 0 and len(v) > 0) else 0)
    return out


def vq_quantize(cb, x):
    """vq_quantize / quantize_mstage (vq_func.py:134-164, 82-131) on an array codebook.
    Returns (q float64 (n,ndim) -- the values numpy returns before any cast, idx (n,stages))."""
    cb = Codebooks._prep_vq(cb)
    x = _f32(x)
    n, nd = x.shape
    a, K = _vq_args(cb)
    q = np.zeros((n, nd), np.float64)
    idx = np.zeros((n, cb.shape[0]), np.int32)
    rc = lib().orc_vq_quantize(a[0], a[3], a[1], a[2], ctypes.c_int(nd), _p(x), ctypes.c_int(n), _p(q), _p(idx))
    if rc:
        raise RuntimeError("orc_vq_quantize failed with status %d" % rc)
    return q, idx


def scl_quantize(codes, x):
    """scl_quantize (vq_func.py:167-185)."""
    codes = Codebooks._prep_scl(codes)
    x = _f32(x).reshape(-1)
    q = np.zeros(len(x), np.float64)
    idx = np.zeros(len(x), np.int32)
    a = _scl_args(codes)
    rc = lib().orc_scl_quantize(a[0], a[2], a[1], _p(x), ctypes.c_int(len(x)), _p(q), _p(idx))
    if rc:
        raise RuntimeError("orc_scl_quantize failed with status %d" % rc)
    return q, idx


def find_nearest(data, codebook):
    """cb_func.find_nearest (cb_func.py:56-68)."""
    data, suf = _train_data(data)
    cb = np.ascontiguousarray(codebook, dtype=np.float64)
    idx = np.zeros(len(data), np.int32)
    rc = getattr(lib(), "orc_find_nearest" + suf)(_p(data), ctypes.c_long(len(data)), _p(cb), ctypes.c_int(len(cb)),
                                ctypes.c_int(data.shape[1]), _p(idx))
    if rc:
        raise RuntimeError("orc_find_nearest failed with status %d" % rc)
    return idx


def kmeans_update(data, codebook, with_details=False):
    """cb_func.update (cb_func.py:71-100): one Lloyd iteration -> new (K,ndim) float64 codebook."""
    data, suf = _train_data(data)
    cb = np.ascontiguousarray(codebook, dtype=np.float64)
    K, nd = cb.shape
    out = np.zeros((K, nd), np.float64)
    idx = np.zeros(len(data), np.int32)
    counts = np.zeros(K, np.float64)
    stats = np.zeros(4, np.float64)
    rc = getattr(lib(), "orc_kmeans_update" + suf)(_p(data), ctypes.c_long(len(data)), _p(cb), ctypes.c_int(K), ctypes.c_int(nd),
                                                   _p(out), _p(idx), _p(counts), _p(stats))
    if rc:
        raise RuntimeError("orc_kmeans_update failed with status %d" % rc)
    if with_details:
        return out, idx, counts, stats
    return out


def kmeans_quantize(codebook, data):
    """cb_func.quantize (cb_func.py:103-112)."""
    data, suf = _train_data(data)
    cb = np.ascontiguousarray(codebook, dtype=np.float64)
    q = np.zeros((len(data), cb.shape[1]), np.float64)
    idx = np.zeros(len(data), np.int32)
    rc = getattr(lib(), "orc_kmeans_quantize" + suf)(_p(data), ctypes.c_long(len(data)), _p(cb), ctypes.c_int(len(cb)),
                                                     ctypes.c_int(cb.shape[1]), _p(q), _p(idx))
    if rc:
        raise RuntimeError("orc_kmeans_quantize failed with status %d" % rc)
    return q, idx


def vq_train(data, codebook, nb_entries, rng):
    """cb_func.vq_train (cb_func.py:28-54): grow-by-one LBG.  `rng` supplies the jitter the
    reference draws from numpy's global RNG (np.random.rand(e, ndims)), so a seeded
    np.random.RandomState reproduces a seeded reference run."""
    data, _ = _train_data(data)
    codebook = np.array(codebook, dtype=np.float64, copy=True)
    ndims = data.shape[1]
    codebook[0] = np.mean(data, 0)
    e = 1
    while e < nb_entries:
        codebook[e, :] = codebook[0, :]
        codebook[:e, :] += .001 * (rng.rand(e, ndims) / 2)
        e += 1
        for _ in range(4):
            codebook[:e, :] = kmeans_update(data, codebook[:e, :])
    for _ in range(10):
        codebook = kmeans_update(data, codebook)
    return codebook


def num_threads():
    return int(lib().orc_num_threads())
