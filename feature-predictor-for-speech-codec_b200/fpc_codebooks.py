"""Device-resident codebook images.

The reference re-reads its codebook .npy files from disk on every single-vector quantiser call
(/root/reference/src/quantization/vq_func.py:141,171).  Here a file is read once, packed once
(fpc_pack_codebooks) and cached by (path, mtime, size, device).
"""
import os

import numpy as np

import fpc_native as N

import collections

_file_cache = {}
# packed device images (~10 MB each), least recently used first; a cfg whose files are rewritten gets a new
# key (mtime), so the table is bounded: the oldest images are dropped beyond _IMAGE_CACHE_MAX entries
_image_cache = collections.OrderedDict()
_IMAGE_CACHE_MAX = 8


def _image_get(key):
    img = _image_cache.get(key)
    if img is not None:
        _image_cache.move_to_end(key)
    return img


def _image_put(key, img):
    _image_cache[key] = img
    while len(_image_cache) > _IMAGE_CACHE_MAX:
        _image_cache.popitem(last=False)
    while len(_file_cache) > 4 * _IMAGE_CACHE_MAX:
        _file_cache.pop(next(iter(_file_cache)))


def load_codebook_file(path):
    """np.load with the reference's conventions; cached on (path, mtime, size)."""
    st = os.stat(path)   # FileNotFoundError propagates like np.load's (vq_func.py:141)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    arr = _file_cache.get(key)
    if arr is None:
        arr = np.load(path, allow_pickle=True)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        _file_cache[key] = arr
    return key, arr


def _check_vq(arr, path):
    if arr.ndim != 3:
        # vq_func.py:143-146 indexes the stage axis before expand_dims: 2-D files raise there too
        raise IndexError("VQ codebook %r must be 3-D (stages, entries, dims), got shape %s" % (path, arr.shape))
    if arr.shape[2] != 17:
        raise ValueError("VQ codebook %r must have 17 code dims, got %d" % (path, arr.shape[2]))
    if arr.shape[0] > 2:
        # vq_func.py:111 raises a broadcast ValueError for 3+ stages
        raise ValueError("VQ codebook %r has %d stages; the m-best search supports 1 or 2" % (path, arr.shape[0]))


class PackedCodebooks:
    """Packed device image of up to four codebooks + the metadata the host side needs."""

    def __init__(self, vq=None, bl_vq=None, scl=None, bl_scl=None, device=None):
        import torch
        N.require_cuda()
        L = N.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.arrays = {"vq": vq, "bl_vq": bl_vq, "scl": scl, "bl_scl": bl_scl}
        c = N.Codebooks()
        self.raw = {}    # device copies of the files in their on-disk layout (kept: the scalar
                         # quantiser entry point reads its code table straight from here)

        def dev(a, name):
            t = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
            self.raw[name] = t
            return t.data_ptr()

        for name in ("vq", "bl_vq"):
            a = self.arrays[name]
            if a is not None:
                _check_vq(a, name)
                setattr(c, name, dev(a, name))
                setattr(c, name + "_dtype", N.FPC_F32 if a.dtype == np.float32 else N.FPC_F64)
                setattr(c, name + "_stages", a.shape[0])
                setattr(c, name + "_entries", a.shape[1])
        for name in ("scl", "bl_scl"):
            a = self.arrays[name]
            if a is not None:
                a = a.reshape(-1)
                setattr(c, name, dev(a, name))
                setattr(c, name + "_dtype", N.FPC_F32 if a.dtype == np.float32 else N.FPC_F64)
                setattr(c, name + "_entries", a.shape[0])
        nbytes = L.fpc_packed_codebooks_bytes()
        self.image = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            N.check(L.fpc_pack_codebooks(c, self.image.data_ptr(), nbytes, N.current_stream(self.device)),
                    "fpc_pack_codebooks")
            torch.cuda.current_stream(self.device).synchronize()

    def ptr(self):
        return self.image.data_ptr()

    def scl_device_ptr(self, name="scl"):
        return self.raw[name].data_ptr()

    def hist_sizes(self):
        a = self.arrays
        return [0 if a["scl"] is None else a["scl"].size, 0 if a["bl_scl"] is None else a["bl_scl"].size,
                0 if a["vq"] is None else a["vq"].shape[1],
                a["vq"].shape[1] if (a["vq"] is not None and a["vq"].shape[0] > 1) else 0,
                0 if a["bl_vq"] is None else a["bl_vq"].shape[1]]


def from_cfg(cfg, device=None):
    """Packed image for the four paths Wavernn.encoder reads from cfg (wavernn.py:219-237);
    '' (or a missing key) means 'no such codebook'."""
    import torch
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    keys, arrs = [], {}
    for name, ck in (("vq", "cb_path"), ("bl_vq", "bl_cb_path"), ("scl", "scl_cb_path"), ("bl_scl", "bl_scl_cb_path")):
        path = cfg.get(ck, "") if hasattr(cfg, "get") else cfg[ck]
        if path:
            k, a = load_codebook_file(path)
            keys.append(k)
            arrs[name] = a
        else:
            keys.append(None)
            arrs[name] = None
    key = (tuple(keys), str(device))
    img = _image_get(key)
    if img is None:
        img = PackedCodebooks(device=device, **arrs)
        _image_put(key, img)
    return img


def single(path, slot, device=None):
    """Packed image holding one file in one slot ('vq' or 'scl'), for the stand-alone quantisers."""
    import torch
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    k, a = load_codebook_file(path)
    key = ((slot, k), str(device))
    img = _image_get(key)
    if img is None:
        img = PackedCodebooks(device=device, **{slot: a})
        _image_put(key, img)
    return img


def clear_cache():
    _file_cache.clear()
    _image_cache.clear()
