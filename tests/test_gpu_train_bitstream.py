"""GPU tests of the rows either side of the closed loop: the training-set compaction and stage
residual (train_cb.py:177-211) against NumPy restatements of the reference lines, the receiver (dequantise + decode
from the 32-bit frame words) against the encoder's own outputs, the LPCNet layout against the reference's as_strided."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch


@pytest.fixture(scope="module")
def model(torch_cuda, state_dict):
    from models.wavernn import Wavernn
    m = Wavernn(20, 384, 128, 18).eval()
    m.load_state_dict(state_dict)
    return m.cuda()


def ref_training_sets(r, r_bl, code_dims=17):
    """train_cb.py:177-187 verbatim in NumPy."""
    scl = np.array([k for k in r[:, :, 0].flatten() if k != 0], dtype=np.float32)
    scl_bl = np.array([k for k in r_bl[:, :, 0].flatten() if k != 0], dtype=np.float32)
    va = r[:, :, -code_dims:].reshape(-1, code_dims)
    va = np.array([va[i] for i in range(len(va)) if sum(abs(va[i])) != 0], dtype=np.float32).reshape(-1, code_dims)
    vb = r_bl[:, :, -code_dims:].reshape(-1, code_dims)
    vb = np.array([vb[i] for i in range(len(vb)) if sum(abs(vb[i])) != 0], dtype=np.float32).reshape(-1, code_dims)
    return scl, scl_bl, va, vb


@pytest.mark.parametrize("B,L", [(3, 50), (37, 61), (130, 40)])
def test_training_sets_match_reference_lines(torch_cuda, model, synth, B, L):
    import fpc_train
    torch = torch_cuda
    feat = torch.from_numpy(synth.make_features(B, L, first_utt=9100)).cuda()
    with torch.no_grad():
        res = model.encode_device({}, feat, None, 0.25, 2.1, qtz=False)
    sets = fpc_train.training_sets(res.r, res.r_under)
    scl, scl_bl, va, vb = ref_training_sets(res.r.cpu().numpy(), res.r_under.cpu().numpy())
    assert len(va) > 0 and len(vb) > 0                          # both branches are exercised at these thresholds
    assert np.array_equal(sets["scl_above"].cpu().numpy(), scl)
    assert np.array_equal(sets["scl_below"].cpu().numpy(), scl_bl)
    assert np.array_equal(sets["vq_above"].cpu().numpy(), va)
    assert np.array_equal(sets["vq_below"].cpu().numpy(), vb)


def test_compact_rows_edges(torch_cuda):
    import fpc_train
    torch = torch_cuda
    g = torch.Generator(device="cpu").manual_seed(5)
    # nothing kept / everything kept / NaN rows are kept (sum(abs(row)) != 0 is True for NaN) / tile boundaries
    z = torch.zeros((3000, 18)).cuda()
    assert fpc_train.compact_rows(z, 1, 17).shape == (0, 17)
    x = torch.randn((5000, 18), generator=g)
    x[::3] = 0
    x[7, 5] = float("nan")
    x[1024, 1:] = 0
    x[1024, 0] = 1.0                                            # only the excluded column is non-zero -> dropped
    xd = x.cuda()
    got = fpc_train.compact_rows(xd, 1, 17).cpu().numpy()
    xn = x.numpy()
    want = np.array([xn[i, 1:] for i in range(len(xn)) if not (np.abs(xn[i, 1:]).sum(dtype=np.float32) == 0)])
    assert np.array_equal(got, want, equal_nan=True)
    assert fpc_train.compact_rows(torch.zeros((0, 18)).cuda(), 0, 1).shape == (0, 1)
    with pytest.raises(Exception):
        fpc_train.compact_rows(x, 1, 17)                        # CPU tensor: no fallback


def test_stage_residual_and_train_stages(torch_cuda, synth, oracle):
    import fpc_train
    from quantization import cb_func
    torch = torch_cuda
    data = synth.make_kmeans_data(20000, seed=3, n_components=64)
    d = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).cuda()
    rng = np.random.RandomState(11)
    cb = rng.randn(64, 17) * 0.1
    nxt = fpc_train.stage_residual(cb, d).cpu().numpy()
    # train_cb.py:199-200 in NumPy: qr = quantize(codebook, r); r = qr - r   (float64, as the reference leaves it)
    idx = oracle.find_nearest(data, cb)
    want = cb[idx] - data
    assert nxt.dtype == np.float64 and np.array_equal(nxt, want)
    assert np.array_equal(fpc_train.stage_residual(cb, d, keep_float64=False).cpu().numpy(), want.astype(np.float32))
    # ... and once more from the float64 vectors of a second stage
    cb2 = rng.randn(32, 17) * 0.03
    nxt2 = fpc_train.stage_residual(cb2, torch.from_numpy(want).cuda()).cpu().numpy()
    assert np.array_equal(nxt2, cb2[oracle.find_nearest(want, cb2)] - want)
    # two stages of 8 entries: stage 2 trains on the flipped residual of stage 1 and reduces the error
    cbs = fpc_train.train_stages(d, [8, 8], rng=np.random.RandomState(0))
    assert len(cbs) == 2 and cbs[0].shape == (8, 17) and cbs[1].shape == (8, 17)
    q1 = cb_func.quantize(cbs[0], data)
    r1 = q1 - data
    q2 = cb_func.quantize(cbs[1], r1)
    e1 = float((r1 ** 2).sum())
    e2 = float(((q2 - r1) ** 2).sum())
    assert e2 < e1


def test_train_stages_later_batches_vs_oracle(torch_cuda, synth, oracle):
    """train_cb.py:205-211 -- every batch after the first: for each stage ten `update` calls on the incoming
    codebook, then r = quantize(cb, r) - r (float64) feeds the next stage.  Restated here with the oracle's update /
    find_nearest and compared codebook by codebook."""
    import fpc_train
    torch = torch_cuda
    data = synth.make_kmeans_data(30000, seed=9, n_components=96)
    rng = np.random.RandomState(5)
    cb_in = [rng.randn(32, 17) * 0.1, rng.randn(16, 17) * 0.03]
    d = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).cuda()
    got = fpc_train.train_stages(d, [32, 16], codebooks=cb_in, first_batch=False)
    r = np.ascontiguousarray(data, dtype=np.float32)
    for i, K in enumerate((32, 16)):
        cb = np.array(cb_in[i], dtype=np.float64)
        for _ in range(10):
            cb = oracle.kmeans_update(r, cb)
        np.testing.assert_allclose(got[i], cb, rtol=1e-9, atol=1e-300, err_msg="stage %d" % i)
        idx = oracle.find_nearest(r, cb)
        r = cb[idx] - r
    # the incoming codebooks were used: a different start gives a different result
    other = fpc_train.train_stages(d, [32, 16], codebooks=[c[::-1].copy() * 1.5 for c in cb_in], first_batch=False)
    assert not np.allclose(other[0], got[0])


def test_two_stage_training_matches_the_reference(torch_cuda):
    """tests/golden/kmeans_stages.npz: the stage loop of train_cb.py:191-211 run with the reference's own cb_func
    (oracle/gen_golden.py).  The second stage trains on float64 vectors, as in the reference; what is left between the
    two runs is the order of the float64 additions inside `update` (1e-12 per iteration, chained over the schedule)."""
    import fpc_train
    from helpers import load_golden
    torch = torch_cuda
    g = load_golden("kmeans_stages")
    data = torch.from_numpy(g["data"]).cuda()
    n_entries = [int(k) for k in g["n_entries"]]
    first = fpc_train.train_stages(data, n_entries, rng=np.random.RandomState(int(g["train_seed"])))
    for i in range(2):
        np.testing.assert_allclose(first[i], g["first_cb%d" % i], rtol=1e-10, atol=1e-300, err_msg="first batch, stage %d" % i)
    # the float64 residual of the first stage, bit for bit, from the reference's codebook
    r1 = fpc_train.stage_residual(g["first_cb0"], data)
    assert r1.dtype == torch.float64 and np.array_equal(r1.cpu().numpy(), g["first_r1"])
    later = fpc_train.train_stages(data, n_entries, codebooks=[g["first_cb0"], g["first_cb1"]], first_batch=False)
    for i in range(2):
        np.testing.assert_allclose(later[i], g["later_cb%d" % i], rtol=1e-10, atol=1e-300, err_msg="later batch, stage %d" % i)


def test_kmeans_float64_vectors_vs_oracle(torch_cuda, synth, oracle):
    """The float64-vector entry points: indices bit-exact (incl. vectors whose float32 rounding would pick another
    centroid), sums / centroids to 1e-12, for small and large K, and the row-order float64 mean."""
    import fpc_native as N
    from quantization import cb_func
    torch = torch_cuda
    g = np.random.Generator(np.random.Philox(key=41))
    for n, K in ((5000, 7), (40000, 200), (30000, 1024)):
        data = synth.make_kmeans_data(n, seed=42, n_components=64).astype(np.float64)
        data += g.standard_normal(data.shape) * 1e-9                 # not float32-representable
        cb = g.standard_normal((K, 17)) * 0.1
        if K >= 7:
            cb[5] = cb[2]
            mid = 0.5 * (cb[0] + cb[1])
            data[:64] = mid + g.standard_normal((64, 17)) * 1e-12    # closer than float32 resolution to a bisector
        assert np.array_equal(cb_func.find_nearest(data, cb), oracle.find_nearest(data, cb).astype(np.int64))
        new = cb_func.update(data, cb, K, verbose=False)
        np.testing.assert_allclose(new, oracle.kmeans_update(data, cb), rtol=1e-12, atol=1e-300)
        d = torch.from_numpy(data).cuda()
        carry = torch.zeros(17, dtype=torch.float64, device="cuda")
        N.check(N.lib().fpc_kmeans_colsum_f64(d.data_ptr(), n, carry.data_ptr(), N.current_stream(d.device)), "colsum")
        assert np.array_equal(np.true_divide(carry.cpu().numpy(), n), np.mean(data, 0))


def test_scalar_codebook_is_a_lloyd_fixed_point(torch_cuda):
    import fpc_train
    torch = torch_cuda
    g = torch.Generator(device="cpu").manual_seed(2)
    x = (torch.randn(200000, generator=g) * 0.3).cuda()
    cb = fpc_train.scalar_codebook(x, 16, iters=200)
    assert cb.shape == (16, 1) and np.all(np.diff(cb[:, 0]) > 0)
    xs = x.double().cpu().numpy()
    near = np.abs(xs[:, None] - cb[None, :, 0]).argmin(1)
    means = np.array([xs[near == k].mean() for k in range(16)])
    assert np.allclose(means, cb[:, 0], atol=2e-4)               # every level is the mean of its cell
    # fewer distinct values than levels: still returns n_levels ascending-or-equal entries, no NaN
    cb2 = fpc_train.scalar_codebook(torch.tensor([0.5, -0.25, 0.5]).cuda(), 16)
    assert cb2.shape == (16, 1) and np.isfinite(cb2).all()


@pytest.mark.parametrize("variant", ["full", "no_below", "f64"])
def test_receiver_reproduces_encoder(torch_cuda, model, synth, variant):
    import fpc_bitstream
    torch = torch_cuda
    with tempfile.TemporaryDirectory(prefix="fpc_bs_") as d:
        cbs = synth.make_codebooks(0, dtype=np.float64 if variant == "f64" else np.float32)
        cfg = synth.save_codebooks(cbs, d)
        if variant == "no_below":
            cfg = dict(cfg, bl_cb_path="", bl_scl_cb_path="")
        B, L = 21, 60
        feat = torch.from_numpy(synth.make_features(B, L, first_utt=9300)).cuda()
        with torch.no_grad():
            res = model.encode_device(cfg, feat, None, 0.25, 2.1, qtz=True)
        torch.cuda.synchronize()
        idx = res.idx
        words = fpc_bitstream.pack_frames(idx)
        assert words.shape == (B, L) and words.dtype == torch.int32
        back = fpc_bitstream.unpack_frames(cfg, words.cpu())          # the words travelled through host memory
        assert torch.equal(back, idx)
        rq = fpc_bitstream.dequantize(cfg, back)
        assert torch.equal(rq, res.r_qtz)
        dec = fpc_bitstream.decode_indices(model, cfg, back, feat[:, :, -2:])
        assert torch.equal(dec, res.c_in)
        if variant == "no_below":
            assert (idx[..., 1] == -1).any()                          # some frame really coded nothing


def test_lpcnet_layout_matches_reference_as_strided(torch_cuda, model, synth):
    import fpc_features
    torch = torch_cuda
    L = 170
    feat = torch.from_numpy(synth.make_features(1, L, first_utt=9400)).cuda()
    feats = fpc_features.lpcnet_features(feat)
    assert feats.shape == (1, L, 36)
    assert torch.equal(feats[..., :20], feat * 24.1)
    chunks = fpc_features.lpcnet_chunks(feats[0], n_chunks=10)
    a = feats.cpu().numpy()
    s = a.strides[-1]
    want = np.lib.stride_tricks.as_strided(a.flatten(), shape=(10, 19, 36), strides=(15 * 36 * s, 36 * s, s))   # generate_qtz_features.py:66-70
    assert np.array_equal(chunks.cpu().numpy(), want)
    allc = fpc_features.lpcnet_chunks(feats)
    assert allc.shape == (1, (L - 19) // 15 + 1, 19, 36)
    with pytest.raises(ValueError):
        fpc_features.lpcnet_chunks(feats[0, :100], n_chunks=10)
