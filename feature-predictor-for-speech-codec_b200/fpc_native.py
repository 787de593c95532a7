"""ctypes binding of csrc/libfpc_b200.so (the C ABI declared in include/fpc_b200.h).

PyTorch is used by the callers only for device memory and streams; what crosses this
boundary is raw device pointers, sizes and a cudaStream_t.  There is no CPU fallback: if the
shared library is missing and cannot be built, or a call returns a non-zero status, an
exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("FPC_B200_LIB") or os.path.join(_HERE, "csrc", "libfpc_b200.so")   # override: A/B builds
_lib = None

FPC_PREC_FP32, FPC_PREC_BF16 = 0, 1
FPC_F32, FPC_F64 = 0, 1
HIST_OFFSETS = (0, 256, 512, 1536, 2560)   # FPC_HIST_* in include/fpc_b200.h
HIST_TOTAL = 3584

EXPORTS = (
    "fpc_version", "fpc_status_string", "fpc_last_cuda_error", "fpc_launch_count",
    "fpc_packed_weights_bytes", "fpc_pack_weights", "fpc_packed_codebooks_bytes", "fpc_pack_codebooks",
    "fpc_encode_workspace_bytes", "fpc_encode", "fpc_encode_plan", "fpc_encode_host_workspace_bytes", "fpc_encode_host",
    "fpc_decode", "fpc_index_histogram",
    "fpc_vq_quantize_packed", "fpc_scl_quantize",
    "fpc_kmeans_workspace_bytes", "fpc_kmeans_assign_accumulate", "fpc_kmeans_finalize", "fpc_kmeans_finalize_acc", "fpc_kmeans_colsum_f32", "fpc_kmeans_colsum_f64", "fpc_kmeans_assign_accumulate_f64",
    "fpc_kmeans_stage_residual_f64", "fpc_kmeans_gather", "fpc_kmeans_ordered_workspace_bytes", "fpc_kmeans_accumulate_ordered",
    "fpc_selftest_umma", "fpc_selftest_tc_scores", "fpc_debug_set_phase_buffer", "fpc_ceps2lpc",
    "fpc_compact_workspace_bytes", "fpc_compact_rows", "fpc_kmeans_stage_residual",
    "fpc_dequantize", "fpc_pack_frames", "fpc_unpack_frames",
)


class FpcError(RuntimeError):
    pass


class Weights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "w_ih1", "w_hh1", "b_ih1", "b_hh1", "w_ih2", "w_hh2", "b_ih2", "b_hh2", "w_fc", "b_fc")]


class Codebooks(ctypes.Structure):
    _fields_ = [
        ("vq", ctypes.c_void_p), ("vq_dtype", ctypes.c_int), ("vq_stages", ctypes.c_int), ("vq_entries", ctypes.c_int),
        ("bl_vq", ctypes.c_void_p), ("bl_vq_dtype", ctypes.c_int), ("bl_vq_stages", ctypes.c_int),
        ("bl_vq_entries", ctypes.c_int),
        ("scl", ctypes.c_void_p), ("scl_dtype", ctypes.c_int), ("scl_entries", ctypes.c_int),
        ("bl_scl", ctypes.c_void_p), ("bl_scl_dtype", ctypes.c_int), ("bl_scl_entries", ctypes.c_int),
    ]


class EncodeIO(ctypes.Structure):
    _fields_ = [
        ("d_feat", ctypes.c_void_p), ("d_mask", ctypes.c_void_p), ("B", ctypes.c_int), ("L", ctypes.c_int),
        ("l1", ctypes.c_float), ("l2", ctypes.c_float), ("qtz", ctypes.c_int),
        ("d_c_in", ctypes.c_void_p), ("d_r", ctypes.c_void_p), ("d_r_qtz", ctypes.c_void_p),
        ("d_r_under", ctypes.c_void_p), ("d_ind1", ctypes.c_void_p), ("d_ind2", ctypes.c_void_p),
        ("d_idx", ctypes.c_void_p),
    ]


class EncodeHostIO(ctypes.Structure):
    _fields_ = [
        ("h_feat", ctypes.c_void_p), ("B", ctypes.c_int), ("L", ctypes.c_int),
        ("l1", ctypes.c_float), ("l2", ctypes.c_float), ("qtz", ctypes.c_int),
        ("h_c_in", ctypes.c_void_p), ("h_r", ctypes.c_void_p), ("h_r_qtz", ctypes.c_void_p),
        ("h_r_under", ctypes.c_void_p), ("h_ind1", ctypes.c_void_p), ("h_ind2", ctypes.c_void_p),
        ("h_idx", ctypes.c_void_p), ("h_hist", ctypes.c_void_p),
    ]


def lib_path():
    return _LIB_PATH


def _build_if_stale():
    """csrc/build.py compares the .so with every .cu / .cuh / header it depends on and returns at once when it is up to
    date, so a stale binary is never loaded after an edit.  Ranks of one job (torchrun) serialise on a lock file."""
    import fcntl
    import importlib.util
    spec = importlib.util.spec_from_file_location("_fpc_csrc_build", os.path.join(_HERE, "csrc", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(os.path.join(_HERE, "csrc", ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            mod.build()
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def lib():
    """Loads csrc/libfpc_b200.so, rebuilding it in-tree first when it is missing or older than one of its sources
    (FPC_B200_LIB names a prebuilt library instead: no build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("FPC_B200_LIB"):
        try:
            _build_if_stale()
        except Exception as exc:  # noqa: BLE001
            if not os.path.exists(_LIB_PATH):
                raise FpcError("libfpc_b200.so is missing and could not be built (%s); there is no CPU fallback. "
                               "Run `python __graft_entry__.py build`." % exc) from exc
            raise FpcError("libfpc_b200.so is older than its sources and could not be rebuilt (%s)" % exc) from exc
    L = ctypes.CDLL(_LIB_PATH)
    vp, ci, cl, cf, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float, ctypes.c_size_t
    L.fpc_version.restype = ci
    L.fpc_status_string.restype = ctypes.c_char_p
    L.fpc_status_string.argtypes = [ci]
    L.fpc_last_cuda_error.restype = ci
    L.fpc_launch_count.restype = ctypes.c_ulonglong
    L.fpc_packed_weights_bytes.restype = cs
    L.fpc_packed_weights_bytes.argtypes = [ci]
    L.fpc_pack_weights.argtypes = [ctypes.POINTER(Weights), ci, vp, cs, vp]
    L.fpc_packed_codebooks_bytes.restype = cs
    L.fpc_pack_codebooks.argtypes = [ctypes.POINTER(Codebooks), vp, cs, vp]
    L.fpc_encode_workspace_bytes.restype = cs
    L.fpc_encode_workspace_bytes.argtypes = [ci, ci, ci]
    L.fpc_encode.argtypes = [vp, vp, ctypes.POINTER(EncodeIO), ci, vp, cs, vp]
    L.fpc_encode_plan.argtypes = [ci, ci, ci, ctypes.POINTER(ci)]
    L.fpc_encode_host_workspace_bytes.restype = cs
    L.fpc_encode_host_workspace_bytes.argtypes = [ci, ci, ci]
    L.fpc_encode_host.argtypes = [vp, vp, ctypes.POINTER(EncodeHostIO), ci, ci, vp, cs, vp]
    L.fpc_decode.argtypes = [vp, vp, vp, ci, ci, vp, ci, vp, cs, vp]
    L.fpc_index_histogram.argtypes = [vp, cl, vp, vp]
    L.fpc_vq_quantize_packed.argtypes = [vp, cl, vp, ci, ci, ci, vp, vp, vp]
    L.fpc_scl_quantize.argtypes = [vp, cl, vp, ci, ci, vp, vp, vp]
    L.fpc_kmeans_workspace_bytes.restype = cs
    L.fpc_kmeans_workspace_bytes.argtypes = [cl, ci]
    L.fpc_kmeans_assign_accumulate.argtypes = [vp, cl, vp, ci, vp, vp, vp, vp, cs, vp]
    L.fpc_kmeans_finalize.argtypes = [vp, vp, ci, ctypes.c_double, vp, vp, vp]
    L.fpc_kmeans_finalize_acc.argtypes = [vp, ci, ctypes.c_double, vp, vp, vp]
    L.fpc_kmeans_colsum_f32.argtypes = [vp, cl, vp, vp]
    L.fpc_kmeans_colsum_f64.argtypes = [vp, cl, vp, vp]
    L.fpc_kmeans_assign_accumulate_f64.argtypes = L.fpc_kmeans_assign_accumulate.argtypes
    L.fpc_kmeans_stage_residual_f64.argtypes = [vp, ci, vp, vp, ci, cl, vp, vp]
    L.fpc_kmeans_gather.argtypes = [vp, ci, vp, cl, vp, vp]
    L.fpc_kmeans_ordered_workspace_bytes.restype = cs
    L.fpc_kmeans_ordered_workspace_bytes.argtypes = [cl, ci]
    L.fpc_kmeans_accumulate_ordered.argtypes = [vp, ci, cl, vp, ci, vp, vp, vp, cs, vp]
    L.fpc_selftest_umma.argtypes = [vp, vp, ci, ci, vp, vp]
    L.fpc_selftest_tc_scores.argtypes = [vp, ci, vp, ci, vp, vp, vp]
    L.fpc_debug_set_phase_buffer.argtypes = [vp]
    L.fpc_ceps2lpc.argtypes = [vp, cl, ci, vp, vp, vp, vp]
    L.fpc_compact_workspace_bytes.restype = cs
    L.fpc_compact_workspace_bytes.argtypes = [cl]
    L.fpc_compact_rows.argtypes = [vp, cl, ci, ci, ci, vp, vp, vp, cs, vp]
    L.fpc_kmeans_stage_residual.argtypes = [vp, ci, vp, vp, cl, vp, vp]
    L.fpc_dequantize.argtypes = [vp, vp, cl, vp, vp]
    L.fpc_pack_frames.argtypes = [vp, cl, vp, vp]
    L.fpc_unpack_frames.argtypes = [vp, vp, cl, vp, vp]
    _lib = L
    return L


def check(status, what):
    if status != 0:
        L = lib()
        msg = L.fpc_status_string(status).decode()
        extra = ""
        if status == 5:
            extra = " (cudaError %d)" % L.fpc_last_cuda_error()
        raise FpcError("%s failed: %s%s" % (what, msg, extra))


def encode_plan(n_utts, precision=FPC_PREC_FP32, sms=0):
    """[(tile height, first utterance, count), ...] -- the launches fpc_encode makes for a batch (fpc_encode_plan)."""
    seg = (ctypes.c_int * 9)()
    n = lib().fpc_encode_plan(int(n_utts), int(precision), int(sms), seg)
    if n < 0:
        check(-n, "fpc_encode_plan")
    return [(seg[3 * i], seg[3 * i + 1], seg[3 * i + 2]) for i in range(n)]


def launch_count():
    return int(lib().fpc_launch_count())


def current_stream(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise FpcError("no CUDA device: the fpc_b200 hot path has no CPU fallback")
