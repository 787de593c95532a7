"""GPU parity tests of the stand-alone quantisers (quantization/vq_func.py) and the k-means
codebook learner (quantization/cb_func.py) against the reference's golden vectors and the oracle.

Bars: codeword / centroid indices bit-exact (integer work); quantised values bit-exact (they are
gathered codewords); k-means centroids within 1e-12 relative of the reference (float64 sums are
accumulated in a different order than the reference's data-order Python loop)."""
import os
import tempfile

import numpy as np
import pytest

from helpers import load_golden

pytestmark = pytest.mark.gpu
CENTROID_RTOL = 1e-12


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; the fpc_b200 path has no CPU fallback")
    return torch


@pytest.fixture(scope="module")
def tmpd():
    with tempfile.TemporaryDirectory(prefix="fpc_q_") as d:
        yield d


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_quantizers_vs_reference_golden(torch_cuda, tmpd, tag):
    from quantization import vq_func
    g = load_golden("quantizers")
    p2, p1, ps = (os.path.join(tmpd, "%s_%s.npy" % (n, tag)) for n in ("cb2", "cb1", "scl"))
    np.save(p2, g["cb2_" + tag]); np.save(p1, g["cb1_" + tag]); np.save(ps, g["scl_" + tag])
    x, xs = g["x_" + tag], g["xs_" + tag]
    q2, tot2 = vq_func.vq_quantize(x, p2)
    assert q2.dtype == g["q2b_" + tag].dtype and np.array_equal(q2, g["q2b_" + tag])
    assert np.array_equal(tot2[0], g["hist2b0_" + tag]) and np.array_equal(tot2[1], g["hist2b1_" + tag])
    _, i2 = vq_func.vq_quantize_indices(x, p2)
    assert np.array_equal(i2, g["i2_" + tag])
    assert i2[0, 0] == 3 and i2[0, 1] == 20           # forced ties -> lowest index
    q1, i1 = vq_func.vq_quantize_indices(x, p1)
    assert np.array_equal(i1, g["i1_" + tag]) and np.array_equal(q1, g["q1_" + tag])
    assert i1[1, 0] == 7
    # one vector per call, as the encoder of the reference does (wavernn.py:230)
    for k in (0, 1, 5):
        q, tot = vq_func.vq_quantize(x[k:k + 1], p2)
        assert np.array_equal(q[0], g["q2_" + tag][k])
        assert int(np.argmax(tot[0])) == g["i2_" + tag][k, 0] and int(np.argmax(tot[1])) == g["i2_" + tag][k, 1]
    qs, tots = vq_func.scl_quantize(xs, ps)
    assert qs.shape == g["qs_" + tag].shape and np.array_equal(qs, g["qs_" + tag])
    assert np.array_equal(tots, g["hists_" + tag])
    # quantize_mstage on arrays
    cs, ix = vq_func.quantize_mstage(x[3], [64, 64], g["cb2_" + tag])
    assert np.array_equal(ix, g["i2_" + tag][3]) and np.array_equal(cs, g["q2_" + tag][3])


def test_vq_large_batch_vs_oracle(torch_cuda, oracle, synth, tmpd):
    from quantization import vq_func
    cbs = synth.make_codebooks(0)
    for name, key in (("above", "cb_path"), ("below", "bl_cb_path")):
        for dt in (np.float32, np.float64):
            p = os.path.join(tmpd, "%s_%s.npy" % (name, np.dtype(dt).name))
            np.save(p, cbs[key].astype(dt))
            g = np.random.Generator(np.random.Philox(key=11))
            x = (g.standard_normal((1000 + 37, 17)) * 0.1).astype(np.float32)
            q, idx = vq_func.vq_quantize_indices(x, p)
            qo, io = oracle.vq_quantize(cbs[key].astype(dt), x)
            assert np.array_equal(idx, io)
            assert np.array_equal(q, qo.astype(dt))
    # tensors in -> tensors out (device-resident use)
    xt = torch_cuda.from_numpy(x).cuda()
    qt, it = vq_func.vq_quantize_indices(xt, p)
    assert qt.is_cuda and np.array_equal(it.cpu().numpy(), io)


def test_kmeans_vs_reference_golden(torch_cuda, oracle):
    from quantization import cb_func
    g = load_golden("kmeans")
    data = g["data"]
    idx0 = cb_func.find_nearest(data, g["cb0"])
    assert idx0.dtype == np.int64 and np.array_equal(idx0, g["idx0"])
    cb1 = cb_func.update(data, g["cb0"], 64, verbose=False)
    assert cb1.dtype == np.float64
    np.testing.assert_allclose(cb1, g["cb1"], rtol=CENTROID_RTOL, atol=1e-300)
    assert np.all(cb1[60:] == 0.0)                    # empty clusters collapse to the zero vector
    cb2 = cb_func.update(data, g["cb1"], 64, verbose=False)
    np.testing.assert_allclose(cb2, g["cb2"], rtol=CENTROID_RTOL, atol=1e-300)
    q = cb_func.quantize(g["cb2"], data)
    assert np.array_equal(q, g["q"])
    # seeded grow-by-one LBG: same jitter stream as the reference's np.random.seed(1234) run
    np.random.seed(int(g["train_seed"]))
    cbt = cb_func.vq_train(g["train_data"], np.zeros((8, 17)), 8)
    # centroid 0 is seeded with NumPy's own float32 row-order mean (fpc_kmeans_colsum_f32): what is left between the two
    # runs is the order of the float64 additions inside `update` (atomics here, data order in the reference)
    np.testing.assert_allclose(cbt, g["train_cb"], rtol=CENTROID_RTOL, atol=1e-300)


def test_ordered_update_equals_the_reference_bit_for_bit(torch_cuda, oracle, synth):
    """ordered=True: per-centroid float64 sums in data order (cb_func.py:82-86) -> `update`, the chained `vq_train`
    schedule and the two-stage loop of train_cb.py return the reference's codebooks to the last bit."""
    from quantization import cb_func
    import fpc_train
    torch = torch_cuda
    g = load_golden("kmeans")
    cb1 = cb_func.update(g["data"], g["cb0"], 64, verbose=False, ordered=True)
    assert np.array_equal(cb1, g["cb1"])
    cb2 = cb_func.update(g["data"], cb1, 64, verbose=False, ordered=True)
    assert np.array_equal(cb2, g["cb2"])
    np.random.seed(int(g["train_seed"]))
    cbt = cb_func.vq_train(g["train_data"], np.zeros((8, 17)), 8, ordered=True)
    assert np.array_equal(cbt, g["train_cb"])
    # train_cb.py:191-211, first batch (vq_train per stage) and a later batch (10 x update per stage): the second stage
    # runs on the float64 residual quantize(cb, r) - r
    gs = load_golden("kmeans_stages")
    data = torch.from_numpy(gs["data"]).cuda()
    n_entries = [int(k) for k in gs["n_entries"]]
    first = fpc_train.train_stages(data, n_entries, rng=np.random.RandomState(int(gs["train_seed"])), ordered=True)
    later = fpc_train.train_stages(data, n_entries, codebooks=[gs["first_cb0"], gs["first_cb1"]], first_batch=False, ordered=True)
    for i in range(2):
        assert np.array_equal(first[i], gs["first_cb%d" % i]), "first batch, stage %d" % i
        assert np.array_equal(later[i], gs["later_cb%d" % i]), "later batch, stage %d" % i
    # larger seeded sets against the oracle: ragged sizes (tile = 2048 rows), tiny and full codebooks, float64 vectors,
    # heavy skew (one centroid owns half the set), empty centroids
    for n, K, seed in ((1, 1, 1), (2047, 3, 2), (2049, 7, 3), (300001, 1024, 4), (150000, 1000, 5), (70001, 512, 6)):
        data = synth.make_kmeans_data(n, seed=seed, n_components=max(2, min(K, 64)))
        if seed == 4:
            data[::2] = data[0] * 0.5                     # skew + exact duplicates (ties broken by the lower index)
        if seed == 5:
            # zeros of both signs, denormals, the smallest normal float and a tiny normal one
            data[7::101, 3] = 0.0
            data[11::103, 4] = -0.0
            data[13::107, 5] = 1e-40
            data[17::109, 6] = -1.4e-45
            data[19::113, 7] = np.float32(1.17549435e-38)
            data[23::127, 8] = np.float32(2.0 ** -100)
        if seed == 6:
            data = data.astype(np.float64) * (1.0 + 2.0 ** -30)          # a later stage's float64 vectors
        cb = np.random.RandomState(seed).randn(K, 17) * 0.1
        cb[K // 2] = 50.0                                 # nobody's nearest: an empty centroid -> the zero vector
        want = oracle.kmeans_update(data, cb)
        got = cb_func.update(data, cb, K, verbose=False, ordered=True)
        assert np.array_equal(got, want), (n, K)
        if K > 1:
            assert np.all(got[K // 2] == 0.0)
        again = cb_func.update(torch.from_numpy(data).cuda(), cb, K, verbose=False, ordered=True)
        assert np.array_equal(again, got)
        loose = cb_func.update(data, cb, K, verbose=False, ordered=False)
        # the atomics path: the same sums in another order (a coordinate that cancels to ~1e-7 keeps 1e-18 absolute)
        np.testing.assert_allclose(loose, want, rtol=CENTROID_RTOL, atol=1e-15)


def test_ordered_update_full_size_is_reproducible(torch_cuda, synth):
    """BASELINE's k-means size class (8 M vectors here, K = 1024): two ordered iterations give the same bits, the
    counts add up to N, and the atomics path agrees to 1e-12."""
    from quantization import cb_func
    torch = torch_cuda
    n, K = 8_000_000, 1024
    g = torch.Generator(device="cuda").manual_seed(7)
    data = torch.randn((n, 17), generator=g, device="cuda", dtype=torch.float32) * 0.1
    cb = torch.from_numpy(np.random.RandomState(7).randn(K, 17) * 0.1).cuda()
    a, sa, _ = cb_func.update_device(data, cb, ordered=True)
    b, sb, _ = cb_func.update_device(data, cb, ordered=True)
    c, sc, _ = cb_func.update_device(data, cb, ordered=False)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(sa, sb)
    assert float(sa[4]) == n and float(sc[4]) == n
    np.testing.assert_allclose(c.cpu().numpy(), a.cpu().numpy(), rtol=CENTROID_RTOL, atol=1e-15)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(5):
        cb_func.update_device(data, cb, ordered=True)
    ev[1].record()
    for _ in range(5):
        cb_func.update_device(data, cb, ordered=False)
    ev[2].record()
    torch.cuda.synchronize()
    print("update on %d vectors, K = %d: ordered %.2f ms, atomics %.2f ms per iteration"
          % (n, K, ev[0].elapsed_time(ev[1]) / 5, ev[1].elapsed_time(ev[2]) / 5))


def test_colsum_is_numpy_float32_mean(torch_cuda, synth):
    """np.mean(data, 0) of float32 data, bit for bit (cb_func.py:34): row-order float32 additions, float32 division."""
    import fpc_native as N
    torch = torch_cuda
    for n in (1, 17, 511, 512, 513, 100003):
        data = (synth.make_kmeans_data(n, seed=5, n_components=16) * 3.0 + 0.25).astype(np.float32)
        d = torch.from_numpy(data).cuda()
        carry = torch.zeros(17, dtype=torch.float32, device="cuda")
        N.check(N.lib().fpc_kmeans_colsum_f32(d.data_ptr(), n, carry.data_ptr(), N.current_stream(d.device)), "colsum")
        got = np.true_divide(carry.cpu().numpy(), n)
        want = np.mean(data, 0)
        assert got.dtype == want.dtype == np.float32
        assert np.array_equal(got, want), (n, got, want)
        # two shards continue one another's sums
        if n > 1:
            carry.zero_()
            h = n // 3
            N.check(N.lib().fpc_kmeans_colsum_f32(d.data_ptr(), h, carry.data_ptr(), N.current_stream(d.device)), "colsum")
            N.check(N.lib().fpc_kmeans_colsum_f32(d[h:].data_ptr(), n - h, carry.data_ptr(), N.current_stream(d.device)), "colsum")
            assert np.array_equal(np.true_divide(carry.cpu().numpy(), n), want)


@pytest.mark.parametrize("N,K", [(1, 1), (31, 1), (5000, 7), (20000, 33), (100003, 512), (60000, 1024)])
def test_kmeans_assign_vs_oracle(torch_cuda, oracle, synth, N, K):
    """Indices are bit-exact for every K of the grow-by-one schedule (arbitrary K, not only 1024),
    including near-ties produced by duplicated centroids and data lying exactly on centroids."""
    from quantization import cb_func
    data = synth.make_kmeans_data(N, seed=21, n_components=64)
    g = np.random.Generator(np.random.Philox(key=22))
    cb = g.standard_normal((K, 17)) * 0.1
    if K >= 7:
        cb[5] = cb[2]                                  # exact duplicate -> ties resolve to index 2
        cb[6] = data[min(N - 1, 3)].astype(np.float64)  # a centroid sitting exactly on a data point
        cb[3] = cb[4] * (1 + 2.0 ** -40)               # near-duplicate below fp32 resolution
    idx = cb_func.find_nearest(data, cb)
    ref = oracle.find_nearest(data, cb)
    assert np.array_equal(idx, ref)
    new = cb_func.update(data, cb, K, verbose=False)
    want, _, counts, stats = oracle.kmeans_update(data, cb, with_details=True)
    np.testing.assert_allclose(new, want, rtol=CENTROID_RTOL, atol=1e-300)


def test_kmeans_device_resident(torch_cuda, oracle, synth):
    from quantization import cb_func
    torch = torch_cuda
    data = synth.make_kmeans_data(30000, seed=23, n_components=32)
    d = torch.from_numpy(data).cuda()
    cb = np.random.Generator(np.random.Philox(key=24)).standard_normal((40, 17)) * 0.1
    cbd = torch.from_numpy(cb).cuda()
    for _ in range(3):
        cbd, stats, n = cb_func.update_device(d, cbd)
        cb = oracle.kmeans_update(data, cb)
        np.testing.assert_allclose(cbd.cpu().numpy(), cb, rtol=1e-11, atol=1e-300)
        cb = cbd.cpu().numpy()      # follow the device trajectory so rounding noise does not compound
    assert n == 30000


def _tc_scores(torch, x, cb):
    import fpc_native as N
    n, K = x.shape[0], cb.shape[0]
    Kp = (K + 127) // 128 * 128
    xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    cd = torch.from_numpy(np.ascontiguousarray(cb, np.float64)).cuda()
    out = torch.zeros((n, Kp), dtype=torch.float32, device="cuda")
    ws = torch.empty(N.lib().fpc_kmeans_workspace_bytes(n, K), dtype=torch.uint8, device="cuda")
    N.check(N.lib().fpc_selftest_tc_scores(xd.data_ptr(), n, cd.data_ptr(), K, out.data_ptr(), ws.data_ptr(),
                                           N.current_stream()), "fpc_selftest_tc_scores")
    torch.cuda.synchronize()
    return out.cpu().numpy()[:, :K]


@pytest.mark.parametrize("scale_x,scale_c", [(0.1, 0.1), (0.03, 0.1), (1.0, 0.05), (1e-4, 0.1), (3.0, 0.1)])
def test_tc_screen_score_error_budget(torch_cuda, scale_x, scale_c):
    """The tensor-core screen (fp16-pair operands, tcgen05, fp32 accumulation) against float64: the error of a score
    must stay far inside the decision slack 2^-15 R = 512 u R, R = (||x|| + Cmax)^2, u = 2^-24 (csrc/fpc_tc.cuh budgets
    <= 10 u R for the operand format; the rest is the tensor core's accumulation, which is what this test measures)."""
    g = np.random.Generator(np.random.Philox(key=31))
    K = 1024
    cb = g.standard_normal((K, 17)) * scale_c
    cb[7] = 0.0
    cb[8, :] = 1e-6 * scale_c                       # tiny entries: fp16 sub-normal hi / lo parts
    x = (g.standard_normal((128, 17)) * scale_x).astype(np.float32)
    x[3] = 0.0
    x[4, 1:] = 0.0
    s = _tc_scores(torch_cuda, x, cb).astype(np.float64)
    cf = cb.astype(np.float32).astype(np.float64)  # the screen sees the fp32 rounding of the book (budgeted separately)
    exact = (cf * cf).sum(1)[None, :] - 2.0 * x.astype(np.float64) @ cf.T
    cmax = np.sqrt((cb * cb).sum(1).max())
    R = (np.sqrt((x.astype(np.float64) ** 2).sum(1)) + cmax) ** 2
    err_u = np.abs(s - exact) / (R[:, None] * 2.0 ** -24)
    print("\ntensor-core screen, |x| ~ %g, |c| ~ %g: max score error %.2f u R, mean %.3f u R" % (
        scale_x, scale_c, err_u.max(), err_u.mean()))
    assert err_u.max() <= 64.0                     # a quarter of the slack per comparison would be 128 u R


@pytest.mark.parametrize("K", [128, 200, 513, 1024])
def test_kmeans_tc_adversarial_vs_oracle(torch_cuda, oracle, synth, K):
    """The tensor-core assignment against the oracle on data built to sit on decision boundaries: exact duplicates of
    centroids, midpoints of centroid pairs (exact distance ties up to rounding), vectors far outside the format's
    range (|x| >> |c|), all-zero vectors, and a book with a cluster of near-identical centroids."""
    from quantization import cb_func
    g = np.random.Generator(np.random.Philox(key=40 + K))
    cb = g.standard_normal((K, 17)) * 0.1
    cb[10:20] = cb[9] + g.standard_normal((10, 17)) * 1e-7      # near-identical cluster
    cb[21] = cb[20]
    base = synth.make_kmeans_data(20000, seed=33, n_components=64)
    mids = (0.5 * (cb[g.integers(0, K, 2000)] + cb[g.integers(0, K, 2000)])).astype(np.float32)
    on = cb[g.integers(0, K, 2000)].astype(np.float32)
    far = (g.standard_normal((500, 17)) * 300.0).astype(np.float32)           # leaves the fp16-pair range
    tiny = (g.standard_normal((500, 17)) * 1e-9).astype(np.float32)
    zero = np.zeros((37, 17), np.float32)
    data = np.ascontiguousarray(np.concatenate([base, mids, on, far, tiny, zero]))
    idx = cb_func.find_nearest(data, cb)
    ref = oracle.find_nearest(data, cb)
    bad = np.flatnonzero(idx != ref)
    assert len(bad) == 0, "%d of %d assignments differ, first at row %d: %d vs %d" % (len(bad), len(data), bad[0], idx[bad[0]], ref[bad[0]])
    new = cb_func.update(data, cb, K, verbose=False)
    want = oracle.kmeans_update(data, cb)
    np.testing.assert_allclose(new, want, rtol=CENTROID_RTOL, atol=1e-300)


def test_kmeans_tc_full_load_equals_cuda_core_path(torch_cuda, synth, monkeypatch):
    """Every SM busy for many tiles: the tensor-core assignment and the CUDA-core one (FPC_KMEANS_TC=0) must give the
    same index for every vector, the same counts, and sums within the float64 addition-order tolerance.  Timing-
    dependent faults (a ring slot released under a load in flight) only show at this size."""
    from quantization import cb_func
    torch = torch_cuda
    n, K = 6_000_000, 1024
    data = torch.from_numpy(synth.make_kmeans_data(n, seed=31, n_components=512)).cuda()
    g = np.random.Generator(np.random.Philox(key=32))
    cb = torch.from_numpy(g.standard_normal((K, 17)) * 0.1).cuda()
    s1, c1, i1 = cb_func.assign_accumulate(data, cb, want_idx=True)
    s1b, c1b, i1b = cb_func.assign_accumulate(data, cb, want_idx=True)
    monkeypatch.setenv("FPC_KMEANS_TC", "0")
    s0, c0, i0 = cb_func.assign_accumulate(data, cb, want_idx=True)
    monkeypatch.delenv("FPC_KMEANS_TC")
    assert torch.equal(i1, i0) and torch.equal(i1, i1b)
    assert torch.equal(c1, c0) and torch.equal(c1, c1b)
    np.testing.assert_allclose(s1.cpu().numpy(), s0.cpu().numpy(), rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(s1.cpu().numpy(), s1b.cpu().numpy(), rtol=1e-11, atol=1e-9)
