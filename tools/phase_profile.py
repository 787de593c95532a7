#!/usr/bin/env python
"""Per-phase cycle breakdown of the frame-step kernels (all CTAs), via fpc_debug_set_phase_buffer.
    python tools/phase_profile.py [utts] [frames] [l1 l2] [bf16]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feature-predictor-for-speech-codec_b200"))
import numpy as np, torch
import fpc_native as N, fpc_synth as S
from models.wavernn import Wavernn
U = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 200
l1, l2 = (float(sys.argv[3]), float(sys.argv[4])) if len(sys.argv) > 4 else (0.09, 0.28)
bf16 = len(sys.argv) > 5 and sys.argv[5] == "bf16"
qtz = os.environ.get("FPC_PROF_QTZ", "1") != "0"      # 0: residual mode (no quantiser in the loop)
m = Wavernn(20, 384, 128, 18).eval(); m.load_state_dict(S.make_state_dict(0)); m = m.cuda()
if bf16:
    m.precision = N.FPC_PREC_BF16
d = tempfile.mkdtemp(); cfg = S.save_codebooks(S.make_codebooks(0), d)
base = S.make_features(min(U, 256), L)
feat = torch.from_numpy(np.tile(base, ((U + len(base) - 1) // len(base), 1, 1))[:U]).cuda()
buf = torch.zeros(16 * 1024, dtype=torch.int64, device="cuda")     # one row of 32 counters per CTA
with torch.no_grad():
    m.encode_device(cfg, feat, None, l1, l2, qtz=qtz); torch.cuda.synchronize()
    N.lib().fpc_debug_set_phase_buffer(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); res = m.encode_device(cfg, feat, None, l1, l2, qtz=qtz); e1.record(); torch.cuda.synchronize()
    N.lib().fpc_debug_set_phase_buffer(None)
c = buf.cpu().numpy().astype(np.float64)[:8192].reshape(-1, 32)
c = c[c[:, 5] > 0]
fr = c[:, 5:6]
names = ["gru", "fc+residual", "thresholds+scalar", "vq", "feedback/out"]
per = c[:, :5] / fr
tot = per.sum(1)
print("kernel %.2f ms for %d x %d frames (%.2f M frames/s); %d CTAs; cycles/frame min %.0f mean %.0f max %.0f" % (
    e0.elapsed_time(e1), U, L, U * L / e0.elapsed_time(e1) / 1e3, len(c), tot.min(), tot.mean(), tot.max()))
for i, n_ in enumerate(names):
    print("  %-18s min %8.0f  mean %8.0f  max %8.0f   %5.1f%% of mean" % (n_, per[:, i].min(), per[:, i].mean(), per[:, i].max(), 100 * per[:, i].mean() / tot.mean()))
print("vq rows %d, sent to the exact search %d (%.2f %%)" % (c[:, 6].sum(), c[:, 7].sum(), 100 * c[:, 7].sum() / max(c[:, 6].sum(), 1)))
for i, n_ in enumerate(["margins", "stage-0 screen", "stage-0 select", "last-stage screen", "merge+gather", "exact fallback"]):
    print("     vq/%-18s mean %8.0f cycles/frame" % (n_, (c[:, 8 + i] / fr[:, 0]).mean()))
print("     vq/wait for MMA units: stage 0 %8.0f, last stage %8.0f cycles/frame (thread 0)" % ((c[:, 14] / fr[:, 0]).mean(), (c[:, 15] / fr[:, 0]).mean()))
if c[:, 16].sum() > 0 or c[:, 17].sum() > 0:
    print("  fp32 roles: GEMM warps busy %8.0f, waiting for the tail %8.0f; tail waiting for h2 %8.0f cycles/frame" % (
        (c[:, 0] / fr[:, 0]).mean(), (c[:, 16] / fr[:, 0]).mean(), (c[:, 17] / fr[:, 0]).mean()))
print("above-threshold fractions: c0 %.3f  c1..17 %.3f" % (res.ind1.mean().item(), res.ind2.mean().item()))

if bf16:
    t = buf.cpu().numpy()[8192:8192 + 512].astype(np.int64)
    if t[0]:
        t0 = t[0]
        print("trace CTA 0 (cycles from the first codebook copy): chunk: copy issued | data landed | units free | MMAs issued | commits issued")
        for c in range(0, 34):
            print("  %2d: %7d %7d %7d %7d %7d" % (c, t[c] - t0, t[64 + c] - t0, t[192 + c] - t0, t[320 + c] - t0, t[128 + c] - t0))
