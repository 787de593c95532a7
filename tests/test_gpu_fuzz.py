"""Seeded random sweep of the CUDA path against the CPU oracle: random batch shapes, thresholds, codebook sizes and
dtypes, optional below-threshold books, masks, host-buffer chunking, random weights -- every index and every float
must equal the oracle's.  Also the stand-alone quantisers and k-means assignment with adversarial inputs (duplicate
codewords, vectors that ARE codewords, zero vectors, tiny and huge scales)."""
import os
import tempfile

import numpy as np
import pytest

from helpers import oracle_codebooks
from test_gpu_encode import assert_bit_exact, run_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch


def random_state_dict(torch, seed, scale):
    """Reference-shaped parameters with a seed-dependent scale, so the recurrence visits other regimes
    than the default U(+-1/sqrt(H)) init (saturated gates, near-zero predictions)."""
    g = torch.Generator().manual_seed(seed)
    def u(*shape, k):
        return (torch.rand(*shape, generator=g) * 2 - 1) * k * scale
    return {
        "rnn1.weight_ih_l0": u(1152, 20, k=384 ** -0.5), "rnn1.weight_hh_l0": u(1152, 384, k=384 ** -0.5),
        "rnn1.bias_ih_l0": u(1152, k=384 ** -0.5), "rnn1.bias_hh_l0": u(1152, k=384 ** -0.5),
        "rnn2.weight_ih_l0": u(384, 384, k=128 ** -0.5), "rnn2.weight_hh_l0": u(384, 128, k=128 ** -0.5),
        "rnn2.bias_ih_l0": u(384, k=128 ** -0.5), "rnn2.bias_hh_l0": u(384, k=128 ** -0.5),
        "dual_fc.0.weight": u(18, 128, k=128 ** -0.5), "dual_fc.0.bias": u(18, k=128 ** -0.5),
    }


@pytest.mark.parametrize("seed", list(range(16)))
def test_encoder_random_configurations(torch_cuda, oracle, synth, seed):
    torch = torch_cuda
    from models.wavernn import Wavernn
    rng = np.random.RandomState(1000 + seed)
    B = int(rng.choice([1, 2, 7, 29, 57, 113, 150, 260]))
    L = int(rng.randint(1, 48))
    dtype = np.float64 if rng.rand() < 0.4 else np.float32
    k_above = int(rng.choice([1024, 1024, 512, 300, 64, 5]))
    k_below = int(rng.choice([512, 512, 100, 5]))     # < 5 entries fail in the reference itself (index[0] = curr_idx, vq_func.py:95)
    n_scl = int(rng.choice([256, 256, 17, 2]))
    n_scl_b = int(rng.choice([16, 16, 5, 1]))
    l1 = float(rng.choice([0.09, 0.25, 0.6, 0.0, 1e9]))
    l2 = float(rng.choice([0.28, 2.1, 4.0, 0.0, 1e9]))
    scale = float(rng.choice([1.0, 1.0, 3.0, 0.2]))
    sd = random_state_dict(torch, seed, scale)
    model = Wavernn(20, 384, 128, 18).eval()
    model.load_state_dict(sd)
    model = model.cuda()
    ow = oracle.weights_from_state_dict({k: v.numpy() for k, v in sd.items()})
    cbs = synth.make_codebooks(seed, l1=max(l1, 0.01) if l1 < 1e8 else 0.1, dtype=dtype, k_above=k_above, k_below=k_below,
                               n_scl=n_scl, n_scl_below=n_scl_b)
    feat = synth.make_features(B, L, first_utt=20000 + 300 * seed)
    if rng.rand() < 0.3:
        feat = (feat * rng.choice([0.0, 4.0])).astype(np.float32)          # silence / loud input
    mode = rng.choice(["qtz", "qtz", "qtz", "no_below", "residual", "mask"])
    with tempfile.TemporaryDirectory(prefix="fpc_fuzz_") as d:
        cfg = synth.save_codebooks(cbs, d)
        C = oracle_codebooks(oracle, cbs)
        mask = None
        qtz = True
        if mode == "no_below":
            cfg = dict(cfg, bl_cb_path="", bl_scl_cb_path="")
            C = oracle.Codebooks(cbs["cb_path"], cbs["scl_cb_path"], None, None)
        elif mode == "residual":
            qtz = False
        elif mode == "mask":
            mask = (rng.rand(B, L, 2) < 0.5).astype(np.float32)
        gpu = run_gpu(torch, model, cfg if qtz else {}, feat, l1, l2, mask, qtz)
        ora = oracle.encode(ow, C if qtz else None, feat, l1, l2, mask=mask, qtz=qtz)
        assert_bit_exact(gpu, ora, qtz)
        if mask is None:
            # the host-buffer call (time-chunked launches with carried state) gives the same bits
            host = model.encode_host(cfg if qtz else {}, torch.from_numpy(feat), l1, l2, qtz=qtz,
                                     chunks=int(rng.randint(1, 6)), want_under=True)
            torch.cuda.synchronize()
            for k in ("c_in", "r", "r_qtz", "r_under", "idx"):
                assert np.array_equal(host[k].numpy(), gpu[k]), "encode_host differs in %s" % k


@pytest.mark.parametrize("seed", list(range(8)))
def test_quantisers_adversarial(torch_cuda, oracle, synth, seed):
    torch = torch_cuda
    from quantization.vq_func import vq_quantize, scl_quantize
    rng = np.random.RandomState(2000 + seed)
    dtype = np.float64 if seed % 2 else np.float32
    stages = int(rng.choice([1, 2]))
    K = int(rng.choice([1024, 777, 64, 8, 5]))     # the reference needs K >= 5 (sorted()[:5] fills five survivor slots)
    cb = (rng.randn(stages, K, 17) * 0.1).astype(np.float32)
    # duplicate codewords (exact ties -> lowest index), a zero codeword
    cb[0, K // 2] = cb[0, 0]
    cb[-1, K - 1] = cb[-1, 1 % K]
    cb[0, (K // 3) % K] = 0.0
    cb = cb.astype(dtype)
    n = int(rng.choice([1, 33, 500]))
    x = (rng.randn(n, 17) * rng.choice([0.1, 1e-4, 30.0])).astype(np.float32)
    x[0] = cb[0, 0].astype(np.float32)                       # a vector that IS a (duplicated) codeword
    if n > 2:
        x[1] = 0.0
        x[2] = (cb[0, K // 2].astype(np.float64) + (cb[-1, 0].astype(np.float64) if stages == 2 else 0)).astype(np.float32)
    with tempfile.TemporaryDirectory(prefix="fpc_fuzzq_") as d:
        p = os.path.join(d, "cb.npy")
        np.save(p, cb)
        q, hist = vq_quantize(x, p)
        qo, io = oracle.vq_quantize(cb, x)
        assert np.array_equal(np.asarray(q), qo)
        ps = os.path.join(d, "scl.npy")
        codes = np.sort(rng.randn(int(rng.choice([256, 16, 3])), 1).astype(np.float32), 0).astype(dtype)
        codes[1 % len(codes)] = codes[0]                     # tie between two levels
        np.save(ps, codes)
        xs = (rng.randn(n, 1) * 0.3).astype(np.float32)
        xs[0, 0] = float(codes[0, 0])
        qs, hs = scl_quantize(xs, ps)
        qso, iso = oracle.scl_quantize(codes, xs[:, 0])
        assert np.array_equal(np.asarray(qs).reshape(-1), np.asarray(qso).reshape(-1))


@pytest.mark.parametrize("seed", list(range(8)))
def test_kmeans_assignment_adversarial(torch_cuda, oracle, seed):
    torch = torch_cuda
    from quantization import cb_func
    rng = np.random.RandomState(3000 + seed)
    K = int(rng.choice([1, 2, 7, 31, 33, 200, 511, 513, 1024]))
    n = int(rng.choice([1, 100, 5000, 40000]))
    scale = float(rng.choice([0.1, 1e-3, 50.0]))
    cb = rng.randn(K, 17) * scale
    if K > 3:
        cb[K - 1] = cb[0]                                    # duplicate centroids: first minimum wins
        cb[K // 2] = 0.0
    data = (rng.randn(n, 17) * scale).astype(np.float32)
    data[0] = cb[0].astype(np.float32)
    if n > 3:
        data[1] = 0.0
        data[2] = ((cb[0] + cb[min(1, K - 1)]) / 2).astype(np.float32)       # near-tie between two centroids
        data[3] = data[2]
    idx = np.asarray(cb_func.find_nearest(data, cb))
    want = oracle.find_nearest(data, cb)
    assert np.array_equal(idx.astype(np.int64), want.astype(np.int64))
    new = cb_func.update(data, cb, K, verbose=False)
    ref = oracle.kmeans_update(data, cb)
    np.testing.assert_allclose(new, ref, rtol=1e-11, atol=1e-300)
