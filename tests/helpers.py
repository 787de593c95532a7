"""Shared helpers for the parity tests."""
import hashlib
import os

import numpy as np

from conftest import GOLDEN


def load_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    return dict(np.load(path, allow_pickle=False))


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        v = sd[k]
        v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        h.update(k.encode())
        h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest()


def golden_codebooks(synth, g):
    """Rebuilds the codebooks a golden encoder case was generated with (seeded generator)."""
    dt = np.float32 if str(g["cb_dtype"]) == "float32" else np.float64
    cbs = synth.make_codebooks(0, l1=synth.L1_README, dtype=dt, k_above=int(g["k_above"]), k_below=int(g["k_below"]))
    if not int(g["below"]):
        cbs["bl_cb_path"] = None
        cbs["bl_scl_cb_path"] = None
    return cbs


def oracle_codebooks(O, cbs):
    return O.Codebooks(cbs["cb_path"], cbs["scl_cb_path"], cbs["bl_cb_path"], cbs["bl_scl_cb_path"])


def index_agreement(idx_a, idx_b):
    """Fraction of frames on which all codebook indices and both branch flags agree."""
    same = np.all(np.asarray(idx_a) == np.asarray(idx_b), axis=-1)
    return float(same.mean()), same


def hist_equal(h_a, h_b):
    for a, b in zip(h_a, h_b):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        if a.shape != b.shape:
            if a.size <= 1 and b.size <= 1:
                if float(a.sum()) != float(b.sum()):
                    return False
                continue
            return False
        if not np.array_equal(a, b):
            return False
    return True
