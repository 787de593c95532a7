"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle tooling).

Run in the build container, where /root/reference exists:

    python oracle/gen_golden.py [case ...]

The reference has no tests, fixtures or golden vectors of its own, so
these files -- inputs plus the outputs of the reference's own code on them -- are what pins
the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Each case stores the
inputs it cannot regenerate cheaply, the seeds of the ones it can (fpc_synth generators, with
a checksum), and everything the reference returned.  Per-frame codebook indices are not
returned by the reference; they are recovered by wrapping the injected quantizers and taking
the argmax of the one-hot histograms each single-vector call returns (wavernn.py:219-240).
"""
import hashlib
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "feature-predictor-for-speech-codec_b200"))

import fpc_synth as S  # noqa: E402
import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def run_encoder_case(name, B, L, l1, l2, cb_dtype=np.float32, below=True, qtz=True, mask=None,
                     first_utt=0, k_above=1024, k_below=512, compact=False):
    """compact=True (the >= 10^4-frame cases): the features are not stored -- tests regenerate them from
    fpc_synth.make_features(B, L, first_utt) and check feat_sha256 -- and neither are r_under (all zero when
    qtz, wavernn.py:245-249) and r (= feat - (c_in - r_qtz) only up to rounding, so the tests compare c_in and
    r_qtz); this keeps a 12 x 1000-frame fixture near 2 MB."""
    import torch
    W, vq_func, _ = ref_shim.load_reference()
    sd = S.make_state_dict(0)
    model = W.Wavernn(20, S.GRU1, S.GRU2, S.N_CEPS).eval()
    model.load_state_dict(sd)
    feat = S.make_features(B, L, first_utt=first_utt)
    cbs = S.make_codebooks(0, l1=S.L1_README, dtype=cb_dtype, k_above=k_above, k_below=k_below)
    if not below:
        cbs["bl_cb_path"] = None
        cbs["bl_scl_cb_path"] = None
    tmp = tempfile.mkdtemp(prefix="fpc_golden_")
    cfg = S.save_codebooks(cbs, tmp)

    scl_calls, vq_calls = [], []

    def scl_q(data, path):
        q, tot = vq_func.scl_quantize(data, path)
        scl_calls.append((path, int(np.argmax(tot))))
        return q, tot

    def vq_q(r, path):
        q, tot = vq_func.vq_quantize(r, path)
        vq_calls.append((path, [int(np.argmax(t)) for t in tot]))
        return q, tot

    t0 = time.time()
    with torch.no_grad():
        m = None if mask is None else torch.tensor(mask)
        c_in, r, r_qtz, r_under, i1, i2, cb_tot = model.encoder(
            cfg, torch.tensor(feat), m, l1, l2, vq_q, scl_q, qtz)
    dt = time.time() - t0
    i1n, i2n = i1.numpy(), i2.numpy()
    idx = np.full((B, L, 4), -1, np.int32)
    if mask is None:
        idx[:, :, 3] = (i1n[:, :, 0] != 0) * 1 + (i2n[:, :, 0] != 0) * 2
    else:   # the reference leaves ind*_mask zero when an external mask is given (wavernn.py:210-212)
        idx[:, :, 3] = (mask[:, :, 0] != 0) * 1 + (mask[:, :, 1] != 0) * 2
    if qtz:
        si = vi = 0
        for i in range(L):          # replay the reference's call order (wavernn.py:217,228)
            for k in range(B):
                if i1n[k, i, 0]:
                    assert scl_calls[si][0] == cfg["scl_cb_path"]
                    idx[k, i, 0] = scl_calls[si][1]; si += 1
                elif cfg["bl_scl_cb_path"]:
                    assert scl_calls[si][0] == cfg["bl_scl_cb_path"]
                    idx[k, i, 0] = scl_calls[si][1]; si += 1
            for k in range(B):
                if i2n[k, i, 0]:
                    assert vq_calls[vi][0] == cfg["cb_path"]
                    idx[k, i, 1], idx[k, i, 2] = vq_calls[vi][1]; vi += 1
                elif cfg["bl_cb_path"]:
                    assert vq_calls[vi][0] == cfg["bl_cb_path"]
                    idx[k, i, 1] = vq_calls[vi][1][0]; vi += 1
        assert si == len(scl_calls) and vi == len(vq_calls)
    hist = {}
    for j, h in enumerate(cb_tot):
        hist["hist%d" % j] = np.asarray(h, dtype=np.float64)
    out = dict(
        l1=np.float64(l1), l2=np.float64(l2), qtz=np.int32(qtz), below=np.int32(below),
        cb_dtype=np.array(np.dtype(cb_dtype).name), k_above=np.int32(k_above), k_below=np.int32(k_below),
        weights_sha256=np.array(sd_checksum(sd)),
        c_in=c_in.numpy(), r_qtz=r_qtz.numpy(),
        ind1=i1n, ind2=i2n, idx=idx, ref_seconds=np.float64(dt), **hist)
    if compact:
        assert qtz and not r_under.numpy().any()
        out.update(B=np.int32(B), L=np.int32(L), first_utt=np.int32(first_utt),
                   feat_sha256=np.array(hashlib.sha256(np.ascontiguousarray(feat).tobytes()).hexdigest()))
        out["idx"] = idx.astype(np.int16)
        out["ind1"] = i1n.astype(np.uint8)
        out["ind2"] = i2n.astype(np.uint8)
    else:
        out.update(feat=feat, r=r.numpy(), r_under=r_under.numpy())
    if mask is not None:
        out["mask"] = mask
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("%-16s B=%d L=%d  %.1fs  (%.1f frames/s)  above: %.2f / %.2f" % (
        name, B, L, dt, B * L / dt, i1n.mean(), i2n.mean()), flush=True)


def run_forward_case():
    import torch
    W, _, _ = ref_shim.load_reference()
    sd = S.make_state_dict(0)
    model = W.Wavernn(20, S.GRU1, S.GRU2, S.N_CEPS).eval()
    model.load_state_dict(sd)
    x = S.make_features(4, 120, first_utt=500)
    with torch.no_grad():
        y, h1, h2 = model(torch.tensor(x))
        # second call continuing from the returned state, as the encoder loop does
        y2, h1b, h2b = model(torch.tensor(x[:, :7]), h1, h2)
    np.savez_compressed(os.path.join(GOLDEN, "forward.npz"), x=x, y=y.numpy(), h1=h1.numpy(), h2=h2.numpy(),
                        y2=y2.numpy(), h1b=h1b.numpy(), h2b=h2b.numpy(), weights_sha256=np.array(sd_checksum(sd)))
    print("forward          done", flush=True)


def run_quantizer_case():
    """Stand-alone vq_quantize / scl_quantize vectors incl. forced exact ties (duplicate codewords,
    a vector equidistant from two scalar levels)."""
    _, vq_func, _ = ref_shim.load_reference()
    g = np.random.Generator(np.random.Philox(key=77))
    tmp = tempfile.mkdtemp(prefix="fpc_golden_q_")
    out = {}
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        cb2 = np.stack([g.standard_normal((64, 17)) * 0.1, g.standard_normal((64, 17)) * 0.03]).astype(np.float32)
        cb2[0, 40] = cb2[0, 3]      # duplicate stage-0 codewords -> exact distance ties
        cb2[0, 41] = cb2[0, 3]
        cb2[1, 20] = cb2[1, 50]     # duplicate stage-1 codewords
        cb1 = (g.standard_normal((1, 48, 17)) * 0.05).astype(np.float32)
        cb1[0, 30] = cb1[0, 7]
        x = (g.standard_normal((40, 17)) * 0.1).astype(np.float32)
        x[0] = cb2[0, 3] + cb2[1, 50]       # exactly representable sum of tied entries
        x[1] = cb1[0, 7]
        x[2] = 0.0
        scl = np.sort(g.standard_normal(32) * 0.3).astype(np.float32)[:, None]
        scl[11] = scl[10]                   # duplicate level
        xs = (g.standard_normal((40, 1)) * 0.3).astype(np.float32)
        xs[0, 0] = np.float32(0.5) * (scl[4, 0] + scl[5, 0])   # midpoint (tie if exactly representable)
        xs[1, 0] = scl[10, 0]
        p2 = os.path.join(tmp, "cb2_%s.npy" % tag); np.save(p2, cb2.astype(dt))
        p1 = os.path.join(tmp, "cb1_%s.npy" % tag); np.save(p1, cb1.astype(dt))
        ps = os.path.join(tmp, "scl_%s.npy" % tag); np.save(ps, scl.astype(dt))
        q2, i2 = [], []
        q1, i1 = [], []
        for v in x:   # the encoder always calls with one vector (wavernn.py:230)
            q, tot = vq_func.vq_quantize(v[None], p2); q2.append(q[0]); i2.append([int(np.argmax(t)) for t in tot])
            q, tot = vq_func.vq_quantize(v[None], p1); q1.append(q[0]); i1.append([int(np.argmax(t)) for t in tot])
        qb, totb = vq_func.vq_quantize(x, p2)   # batched call form
        qs, tots = vq_func.scl_quantize(xs, ps)
        si = []
        for v in xs:
            q, tot = vq_func.scl_quantize(v[None], ps); si.append(int(np.argmax(tot)))
        out.update({
            "cb2_" + tag: cb2.astype(dt), "cb1_" + tag: cb1.astype(dt), "scl_" + tag: scl.astype(dt),
            "x_" + tag: x, "xs_" + tag: xs,
            "q2_" + tag: np.array(q2), "i2_" + tag: np.array(i2, np.int32),
            "q1_" + tag: np.array(q1), "i1_" + tag: np.array(i1, np.int32),
            "q2b_" + tag: qb, "hist2b0_" + tag: totb[0], "hist2b1_" + tag: totb[1],
            "qs_" + tag: qs, "hists_" + tag: tots, "is_" + tag: np.array(si, np.int32)})
    np.savez_compressed(os.path.join(GOLDEN, "quantizers.npz"), **out)
    print("quantizers       done", flush=True)


def run_kmeans_case():
    import contextlib
    import io
    _, _, cb_func = ref_shim.load_reference()
    data = S.make_kmeans_data(3000, seed=5, n_components=40)
    g = np.random.Generator(np.random.Philox(key=6))
    cb0 = (g.standard_normal((64, 17)) * 0.1)
    cb0[60:] = 10.0     # far-away centroids -> empty clusters -> zero vectors (cb_func.py:88)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        idx0 = cb_func.find_nearest(data, cb0)
        cb1 = cb_func.update(data, cb0, 64)
        cb2 = cb_func.update(data, cb1, 64)
        q = cb_func.quantize(cb2, data)
        # seeded grow-by-one LBG (cb_func.py:28-54 draws from numpy's global RNG)
        small = data[:400]
        np.random.seed(1234)
        cbt = cb_func.vq_train(small, np.zeros((8, 17)), 8)
    np.savez_compressed(os.path.join(GOLDEN, "kmeans.npz"), data=data, cb0=cb0, idx0=idx0.astype(np.int32),
                        cb1=cb1, cb2=cb2, q=q, train_data=small, train_seed=np.int32(1234), train_cb=cbt,
                        update_log=np.array(sink.getvalue().splitlines()[:2]))
    print("kmeans           done", flush=True)


def run_kmeans_stages_case():
    """The stage loop of train_cb.py:191-211 with the reference's own cb_func: the second stage trains on
    r = quantize(cb, r) - r, which is FLOAT64 (float64 codebook minus float32 residual, :200)."""
    import contextlib
    import io
    _, _, cb_func = ref_shim.load_reference()
    r0 = S.make_kmeans_data(1500, seed=9, n_components=24)
    n_entries = [8, 8]
    out = {"data": r0, "n_entries": np.array(n_entries, np.int32), "train_seed": np.int32(4321)}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        # first batch (:193-200)
        np.random.seed(4321)
        r = r0
        first = []
        for i in range(2):
            cb = cb_func.vq_train(r, np.zeros((n_entries[i], 17)), n_entries[i])
            qr = cb_func.quantize(cb, r)
            r = qr - r
            first.append(cb)
            out["first_r%d" % (i + 1)] = r
        # a later batch (:205-211): ten updates per stage from the codebooks of the first
        r = r0
        later = []
        for i in range(2):
            cb = first[i]
            for _ in range(10):
                cb = cb_func.update(r, cb, n_entries[i])
            qr = cb_func.quantize(cb, r)
            r = qr - r
            later.append(cb)
    assert out["first_r1"].dtype == np.float64
    out.update(first_cb0=first[0], first_cb1=first[1], later_cb0=later[0], later_cb1=later[1])
    np.savez_compressed(os.path.join(GOLDEN, "kmeans_stages.npz"), **out)
    print("kmeans_stages    done", flush=True)


def run_ceps2lpc_case():
    """ceps2lpc_v (ceps2lpc/ceps2lpc_vct.py:122-162) on de-normalised synthetic cepstra (x 24.1, as the callers do at
    synthesis_qtz.py:158) incl. an all-zero frame and a ramp."""
    import importlib.util
    import torch
    ref_shim.load_reference()
    spec = importlib.util.spec_from_file_location("ref_ceps2lpc", os.path.join(ref_shim.REFERENCE_SRC, "ceps2lpc", "ceps2lpc_vct.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    x = (S.make_features(4, 100, first_utt=300) * 24.1).reshape(-1, 20).astype(np.float32)
    x[5] = 0
    x[17, :18] = np.linspace(-3, 3, 18)
    e, lpc, rc = mod.ceps2lpc_v(torch.tensor(x))
    np.savez_compressed(os.path.join(GOLDEN, "ceps2lpc.npz"), x=x, lpc=lpc.numpy(), e_last=np.float32(float(e)),
                        rc_last=rc.numpy())
    print("ceps2lpc         done", flush=True)


CASES = {
    "forward": run_forward_case,
    "ceps2lpc": run_ceps2lpc_case,
    "quantizers": run_quantizer_case,
    "kmeans": run_kmeans_case,
    "kmeans_stages": run_kmeans_stages_case,
    # BASELINE.json configs[0]: 3 utterances x 3 s, README thresholds
    "cfg1_readme": lambda: run_encoder_case("cfg1_readme", 3, 300, S.L1_README, S.L2_README),
    # calibrated thresholds (about half the frames below) so both branches are exercised
    "calibrated": lambda: run_encoder_case("calibrated", 3, 200, 0.25, 2.1, first_utt=10),
    "f64cb": lambda: run_encoder_case("f64cb", 2, 100, 0.25, 2.1, cb_dtype=np.float64, first_utt=20),
    "no_below": lambda: run_encoder_case("no_below", 2, 80, 0.25, 2.1, below=False, first_utt=30),
    "qtz0": lambda: run_encoder_case("qtz0", 3, 100, 0.25, 2.1, qtz=False, first_utt=40),
    "mask_b1": lambda: run_encoder_case(
        "mask_b1", 1, 60, 0.25, 2.1, qtz=False, first_utt=50,
        mask=(np.random.Generator(np.random.Philox(key=9)).uniform(0, 1, (1, 60, 2)) > 0.5).astype(np.float32)),
    "smallcb": lambda: run_encoder_case("smallcb", 4, 150, 0.25, 2.1, first_utt=60, k_above=32, k_below=16),
    # >= 10^4 coded frames of the unmodified reference per threshold pair (one mismatch = 0.008 %):
    # 12 utterances x 10 s, the frame count of BASELINE.json configs[1]
    "long_readme": lambda: run_encoder_case("long_readme", 12, 1000, S.L1_README, S.L2_README, first_utt=100, compact=True),
    "long_calibrated": lambda: run_encoder_case("long_calibrated", 12, 1000, 0.25, 2.1, first_utt=200, compact=True),
}

if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        CASES[n]()
