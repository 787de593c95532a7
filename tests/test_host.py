"""CPU-only tests: the C-ABI library loads and exports every symbol include/fpc_b200.h declares,
argument validation that needs no GPU, the host-side sharding logic, a world_size-2 gloo run of
the k-means reduce step, and the rule that the product never touches oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "fpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import fpc_native as N
    L = N.lib()
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "libfpc_b200.so does not export %s" % n
    assert set(N.EXPORTS) == set(names), "fpc_native.EXPORTS and the header disagree"
    assert L.fpc_version() == 100
    assert L.fpc_status_string(3).decode().startswith("bad codebook")


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device, and return (never throw)."""
    import fpc_native as N
    L = N.lib()
    assert L.fpc_encode(None, None, None, 0, None, 0, None) == 1                      # FPC_ERR_ARG
    io = N.EncodeIO()
    io.B, io.L = 0, 10
    dummy = ctypes.c_void_p(16)
    assert L.fpc_encode(dummy, None, ctypes.byref(io), 0, None, 0, None) == 0          # empty batch is a no-op
    io.B = -1
    assert L.fpc_encode(dummy, None, ctypes.byref(io), 0, None, 0, None) == 1
    assert L.fpc_pack_weights(None, 0, None, 0, None) == 1
    assert L.fpc_packed_weights_bytes(0) > 2_600_000
    assert L.fpc_packed_codebooks_bytes() > 0
    assert L.fpc_kmeans_assign_accumulate(None, 0, None, 4, None, None, None, None, 0, None) == 0   # N == 0
    assert L.fpc_kmeans_assign_accumulate(dummy, 5, dummy, 4096, dummy, dummy, None, None, 0, None) == 3
    assert L.fpc_scl_quantize(dummy, 3, dummy, 0, 1000, dummy, dummy, None) == 3       # > 256 levels
    # ordered sums (cb_func.py:82-86 in data order): sizes this build does not take, empty shard, short workspace
    assert L.fpc_kmeans_ordered_workspace_bytes(-1, 8) == 0 and L.fpc_kmeans_ordered_workspace_bytes(10, 4096) == 0
    need = L.fpc_kmeans_ordered_workspace_bytes(100_000, 1024)
    assert need >= 100_000 * 4 + 49 * 1024 * 4           # the row permutation + one count per (2048-row tile, centroid)
    assert L.fpc_kmeans_accumulate_ordered(dummy, 0, 5, dummy, 4096, dummy, dummy, dummy, need, None) == 2    # K > 2048
    assert L.fpc_kmeans_accumulate_ordered(None, 0, 0, None, 8, dummy, dummy, None, 0, None) == 0              # N == 0
    assert L.fpc_kmeans_accumulate_ordered(dummy, 0, 5, dummy, 8, None, None, dummy, need, None) == 1
    assert L.fpc_kmeans_accumulate_ordered(dummy, 0, 100_000, dummy, 1024, dummy, dummy, dummy, 16, None) == 4
    cb = N.Codebooks()
    cb.vq, cb.vq_dtype, cb.vq_stages, cb.vq_entries = 16, 0, 3, 64                      # 3 stages: vq_func.py:111 raises
    assert L.fpc_pack_codebooks(ctypes.byref(cb), dummy, L.fpc_packed_codebooks_bytes(), None) == 3
    cb.vq_stages, cb.vq_entries = 2, 4                                                 # fewer than SURVIVORS entries
    assert L.fpc_pack_codebooks(ctypes.byref(cb), dummy, L.fpc_packed_codebooks_bytes(), None) == 3
    assert L.fpc_pack_codebooks(ctypes.byref(cb), dummy, 16, None) == 4                # FPC_ERR_WORKSPACE


def test_no_cpu_fallback_and_module_keys():
    import torch
    from models.wavernn import Wavernn
    import fpc_native as N
    m = Wavernn(20, 384, 128, 18)
    assert sorted(m.state_dict()) == sorted([
        "rnn1.weight_ih_l0", "rnn1.weight_hh_l0", "rnn1.bias_ih_l0", "rnn1.bias_hh_l0",
        "rnn2.weight_ih_l0", "rnn2.weight_hh_l0", "rnn2.bias_ih_l0", "rnn2.bias_hh_l0",
        "dual_fc.0.weight", "dual_fc.0.bias"])
    assert sum(p.numel() for p in m.parameters()) == 667410
    y, h1, h2 = m(torch.zeros(2, 3, 20))           # teacher-forced forward stays plain torch
    assert y.shape == (2, 3, 18) and h1.shape == (1, 2, 384) and h2.shape == (1, 2, 128)
    if not torch.cuda.is_available():
        with pytest.raises(N.FpcError):
            m.encoder({}, torch.zeros(1, 2, 20), None, 0.1, 0.3, None, None, True)
        from quantization import vq_func, cb_func
        with pytest.raises(N.FpcError):
            vq_func.vq_quantize(np.zeros((1, 17), np.float32), "/nonexistent.npy")
        with pytest.raises(N.FpcError):
            cb_func.find_nearest(np.zeros((4, 17), np.float32), np.zeros((2, 17)))


def test_forward_matches_oracle_cpu(oracle, oracle_weights, state_dict, synth):
    """The drop-in module's torch forward and the oracle's canonical predictor agree to 2e-6
    (what separates them is summation order inside torch's GEMM and 2-ulp transcendentals)."""
    import torch
    from models.wavernn import Wavernn
    m = Wavernn(20, 384, 128, 18).eval()
    m.load_state_dict(state_dict)
    x = synth.make_features(3, 25, first_utt=77)
    with torch.no_grad():
        y, h1, h2 = m(torch.from_numpy(x))
    yo, h1o, h2o = oracle.forward(oracle_weights, x)
    assert np.abs(y.numpy() - yo).max() < 2e-6 and np.abs(h1.numpy()[0] - h1o).max() < 2e-6


def test_encode_launch_plan_covers_the_batch():
    """fpc_encode_plan (host arithmetic, no CUDA call when the SM count is given): consecutive utterance ranges that
    cover the batch exactly once with supported tile heights; the BASELINE.json configs[4] shard (12 500 utterances on
    148 SMs) must not be rounded up to whole waves of the tallest tile."""
    import fpc_native
    for prec, heights in ((fpc_native.FPC_PREC_FP32, (16, 24, 28, 32)), (fpc_native.FPC_PREC_BF16, (32, 64))):
        for B in (1, 3, 31, 148 * 32, 148 * 32 + 1, 4096, 12500, 16384, 50000, 100000):
            plan = fpc_native.encode_plan(B, prec, 148)
            assert 1 <= len(plan) <= 3
            pos = 0
            for h, first, count in plan:
                assert h in heights and first == pos and count > 0
                pos += count
            assert pos == B
    cost = lambda plan: sum(-(-(-(-c // h)) // 148) * (h + 6) for h, _, c in plan)   # waves x (tile height + fixed)
    assert cost(fpc_native.encode_plan(12500, fpc_native.FPC_PREC_FP32, 148)) <= 106    # 3 waves of 32 would be 114
    assert fpc_native.encode_plan(4096, fpc_native.FPC_PREC_FP32, 148) == [(28, 0, 4096)]
    with pytest.raises(fpc_native.FpcError):
        fpc_native.encode_plan(-1, fpc_native.FPC_PREC_FP32, 148)


def test_below_threshold_histogram_counts_the_last_stage(oracle):
    """cb_tot[4] += cb_t[-1] (wavernn.py:240): with a TWO-stage below-threshold book the table counts stage-2 indices."""
    cb2 = np.zeros((2, 8, 17), np.float32)
    C = oracle.Codebooks(cb2, np.zeros((4, 1), np.float32), cb2, np.zeros((4, 1), np.float32))
    idx = np.array([[[1, 5, 6, 3]], [[2, 3, 7, 1]], [[0, 2, 4, 0]]], np.int32)     # flags: above/above, c0 only, below/below
    h = oracle.histograms(idx, C)
    assert h[2][5] == 1 and h[3][6] == 1                  # above: both stages counted
    assert h[4][7] == 1 and h[4][4] == 1 and h[4].sum() == 2 and h[4][3] == 0 and h[4][2] == 0


def test_shard_range():
    import fpc_dist
    for n in (0, 1, 7, 4096, 100000):
        for w in (1, 2, 3, 8):
            got = [fpc_dist.shard_range(n, r, w) for r in range(w)]
            assert sum(c for _, c in got) == n
            assert got[0][0] == 0 and all(got[i][0] + got[i][1] == got[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(ValueError):
        fpc_dist.shard_range(10, 2, 2)


def test_product_never_touches_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                # (comments may cite the oracle as the arithmetic twin; code may not reach it)
                if re.search(r"import\s+oracle|from\s+oracle|libfpc_oracle|orc_[a-z_0-9]+\s*\(|oracle[/\\]|ref_shim", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, "product sources mention the oracle: %s" % bad


_GLOO_WORKER = r"""
import os, sys
sys.path[:0] = [%(pkg)r, %(orc)r]
import numpy as np, torch, torch.distributed as dist
import fpc_dist, fpc_synth, oracle as O
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
N, K = 4001, 24
data = fpc_synth.make_kmeans_data(N, seed=31, n_components=16)
cb = np.random.Generator(np.random.Philox(key=32)).standard_normal((K, 17)) * 0.1
first, cnt = fpc_dist.shard_range(N, rank, world)
shard = data[first:first + cnt]
# local assign + accumulate (the oracle stands in for the CUDA kernel on this CPU-only box)
idx = O.find_nearest(shard, cb)
sums = np.zeros((K, 17)); counts = np.zeros(K)
np.add.at(sums, idx, shard.astype(np.float64)); np.add.at(counts, idx, 1.0)
ts, tc = torch.from_numpy(sums), torch.from_numpy(counts)
n_total = fpc_dist.allreduce_kmeans(ts, tc, cnt)
new = ts.numpy() / (tc.numpy()[:, None] + 1e-20)
want = O.kmeans_update(data, cb)
assert n_total == N, n_total
assert np.allclose(new, want, rtol=1e-12, atol=0), np.abs(new - want).max()
# jitter broadcast: every rank ends with rank 0's draw
j = fpc_dist.broadcast_array(np.random.RandomState(100 + rank).rand(3, 17))
assert np.array_equal(j, np.random.RandomState(100).rand(3, 17))
# histogram merge incl. never-hit tables (the int 0 of wavernn.py:189)
h = [np.arange(4.0) * (rank + 1), 0, np.ones(3) if rank == 1 else 0, 0, 0]
m = fpc_dist.merge_histograms(h)
assert np.array_equal(m[0], np.arange(4.0) * 3) and m[1] == 0 and np.array_equal(m[2], np.ones(3)) and m[3] == 0
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_kmeans_reduce_world_size_2_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"pkg": PKG, "orc": os.path.join(ROOT, "oracle")})
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "ok" in o, "rank %d failed:\n%s" % (r, o)


def test_lpcnet_chunks_is_the_reference_strided_view():
    """fpc_features.lpcnet_chunks against the as_strided of generate_qtz_features.py:66-70 (pure tensor logic, CPU)."""
    import torch
    import fpc_features
    L = 200
    a = np.arange(L * 36, dtype=np.float32).reshape(1, L, 36)
    s = a.strides[-1]
    want = np.lib.stride_tricks.as_strided(a.flatten(), shape=(10, 19, 36), strides=(15 * 36 * s, 36 * s, s))
    got = fpc_features.lpcnet_chunks(torch.from_numpy(a)[0], n_chunks=10)
    assert np.array_equal(got.numpy(), want)
    every = fpc_features.lpcnet_chunks(torch.from_numpy(a))
    assert tuple(every.shape) == (1, (L - 19) // 15 + 1, 19, 36)
    assert np.array_equal(every[0, -1].numpy(), a[0, 15 * (every.shape[1] - 1):15 * (every.shape[1] - 1) + 19])
    with pytest.raises(ValueError):
        fpc_features.lpcnet_chunks(torch.from_numpy(a)[0, :150], n_chunks=10)     # the reference would read past the end
    assert tuple(fpc_features.lpcnet_chunks(torch.from_numpy(a)[0, :10]).shape) == (0, 19, 36)


def test_frame_word_layout_is_documented_and_disjoint():
    import fpc_bitstream
    bits = np.zeros(32, dtype=int)
    for name, (lo, n) in fpc_bitstream.WORD_BITS.items():
        bits[lo:lo + n] += 1
    assert bits.max() == 1 and bits[:30].min() == 1 and bits[30:].sum() == 0     # 30 payload bits, no overlap
    hdr = open(os.path.join(ROOT, "include", "fpc_b200.h")).read()
    assert "bits 2-9 scalar index" in hdr and "bits 20-29 VQ stage-2 index" in hdr


def test_no_cpu_fallback_in_new_host_modules():
    """CPU tensors are an error, not a slow path (no GPU needed to check that)."""
    import torch
    import fpc_native
    import fpc_train
    import fpc_features
    if torch.cuda.is_available():
        pytest.skip("checks the behaviour without a GPU")
    with pytest.raises(fpc_native.FpcError):
        fpc_train.compact_rows(torch.zeros(4, 18), 1, 17)
    with pytest.raises(fpc_native.FpcError):
        fpc_features.lpcnet_features(torch.zeros(1, 4, 20))
    with pytest.raises(fpc_native.FpcError):
        fpc_train.scalar_codebook(torch.zeros(10), 4)
