"""Seeded synthetic workloads for the closed-loop encode and k-means paths.

There is no dataset, checkpoint or codebook in the reference tree (its .gitignore excludes
them), so BASELINE.json's configs are defined on synthetic inputs.  The seeds and shapes
here are part of the measurement contract: the golden fixtures under
tests/golden/, the parity tests and bench.py all draw from these generators, and features
are seeded per utterance id so a shard of utterances is identical no matter how many GPUs
the batch is split over.
"""
import numpy as np

IN_FEATURES = 20   # 18 cepstra + 2 pitch parameters (LPCNet features, wavernn.py:178,196)
N_CEPS = 18
CODE_DIMS = 17     # c1..c17 go to the VQ, c0 to the scalar quantizer (wavernn.py:219,230)
GRU1, GRU2 = 384, 128
L1_README, L2_README = 0.09, 0.28   # README.md:26,29,37,44


def make_state_dict(seed=0, in_features=IN_FEATURES, gru_units1=GRU1, gru_units2=GRU2, fc_units=N_CEPS):
    """Random-init predictor weights with the parameter names of wavernn.py:37-38,48-52.

    Built from torch.nn.GRU / Linear in the constructor order of the reference's Wavernn, so
    under the same torch.manual_seed the tensors equal those of Wavernn(20, 384, 128, 18)."""
    import torch
    from torch import nn
    torch.manual_seed(seed)
    rnn1 = nn.GRU(in_features, gru_units1, 1, batch_first=True)
    rnn2 = nn.GRU(gru_units1, gru_units2, 1, batch_first=True)
    fc = nn.Linear(gru_units2, fc_units)
    sd = {}
    for name, mod in (("rnn1", rnn1), ("rnn2", rnn2)):
        for k, v in mod.state_dict().items():
            sd["%s.%s" % (name, k)] = v.detach().clone()
    sd["dual_fc.0.weight"] = fc.weight.detach().clone()
    sd["dual_fc.0.bias"] = fc.bias.detach().clone()
    return sd


def make_features(n_utts, n_frames, first_utt=0, dtype=np.float32):
    """(n_utts, n_frames, 20) normalised LPCNet-like features.  Utterance u uses seed 1000+u:
    AR(1) x_t = 0.95 x_{t-1} + eps per cepstral dim with stationary std 0.3 (c0) / 0.1
    (c1..c17), pitch dims piecewise constant U(-1,1) with 20-frame segments."""
    from scipy.signal import lfilter
    out = np.empty((n_utts, n_frames, IN_FEATURES), dtype=np.float64)
    rho = 0.95
    std = np.full(N_CEPS, 0.1)
    std[0] = 0.3
    sig = std * np.sqrt(1.0 - rho * rho)
    nseg = (n_frames + 19) // 20
    for i in range(n_utts):
        g = np.random.Generator(np.random.Philox(key=1000 + first_utt + i))
        eps = g.standard_normal((n_frames, N_CEPS)) * sig
        eps[0] = g.standard_normal(N_CEPS) * std          # start in the stationary distribution
        out[i, :, :N_CEPS] = lfilter([1.0], [1.0, -rho], eps, axis=0)
        seg = g.uniform(-1.0, 1.0, (nseg, 2))
        out[i, :, N_CEPS:] = np.repeat(seg, 20, axis=0)[:n_frames]
    return out.astype(dtype)


def make_codebooks(seed=0, l1=L1_README, dtype=np.float32, k_above=1024, k_below=512,
                   n_scl=256, n_scl_below=16):
    """Random codebooks in the on-disk layout the reference expects:
    VQ above (2,K,17), VQ below (1,K,17), scalar (n,1).  Values are float32-representable;
    `dtype` selects the file dtype (train_cb.py writes float64, vq_func.py:18 then computes
    in float64)."""
    g = np.random.Generator(np.random.Philox(key=seed))
    vq = np.stack([g.standard_normal((k_above, CODE_DIMS)) * 0.1,
                   g.standard_normal((k_above, CODE_DIMS)) * 0.03]).astype(np.float32)
    bl = (g.standard_normal((1, k_below, CODE_DIMS)) * 0.01).astype(np.float32)
    scl = np.sort(g.standard_normal(n_scl) * 0.3).astype(np.float32)[:, None]
    bl_scl = np.linspace(-l1, l1, n_scl_below).astype(np.float32)[:, None]
    return {"cb_path": vq.astype(dtype), "bl_cb_path": bl.astype(dtype),
            "scl_cb_path": scl.astype(dtype), "bl_scl_cb_path": bl_scl.astype(dtype)}


def save_codebooks(cbs, directory, tag=""):
    """Writes the four .npy files and returns the cfg dict Wavernn.encoder reads
    (wavernn.py:219-237)."""
    import os
    os.makedirs(directory, exist_ok=True)
    cfg = {}
    for key, arr in cbs.items():
        if arr is None:
            cfg[key] = ""
            continue
        path = os.path.join(directory, "%s%s.npy" % (key.replace("_path", ""), tag))
        np.save(path, arr)
        cfg[key] = path
    return cfg


def make_kmeans_data(n, seed=0, n_components=2048, ndim=CODE_DIMS, chunk=1 << 20):
    """(n, 17) float32 residual-like vectors: mixture of `n_components` Gaussians (centres
    N(0, 0.1^2), spread 0.03), no all-zero rows (train_cb.py:187 drops those)."""
    g = np.random.Generator(np.random.Philox(key=seed))
    centres = g.standard_normal((n_components, ndim)) * 0.1
    out = np.empty((n, ndim), dtype=np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        comp = g.integers(0, n_components, e - s)
        out[s:e] = (centres[comp] + g.standard_normal((e - s, ndim)) * 0.03).astype(np.float32)
    return out
