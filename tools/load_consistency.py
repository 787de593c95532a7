#!/usr/bin/env python
"""Every copy of an utterance in a large batch (all SMs busy) against a small reference run of the same utterances:
per tile, the first frame and the rows whose residual differs.  Used to find timing-dependent faults of the fp32
frame-step kernel (a weight-ring stage released under a load in flight showed up as whole row groups of a tile being
off by ~1e-4 under load only).
    [FPC_FP32_TILE=16|24|28|32] [QTZ=0] python tools/load_consistency.py <utterances> <frames> <tile height> [l1 l2]"""
import os, sys, numpy as np, torch, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feature-predictor-for-speech-codec_b200")); sys.path.insert(0, ROOT)
import fpc_native as N, fpc_synth as S
from models.wavernn import Wavernn
m = Wavernn(20, 384, 128, 18).eval(); m.load_state_dict(S.make_state_dict(0)); m = m.cuda()
d = tempfile.mkdtemp(); cfg = S.save_codebooks(S.make_codebooks(0), d)
B = int(sys.argv[1]); L = int(sys.argv[2]); MT = int(sys.argv[3])
l1, l2 = (0.25, 2.1) if len(sys.argv) < 5 else (float(sys.argv[4]), float(sys.argv[5]))
base = S.make_features(64, L)
feat = torch.from_numpy(np.tile(base, ((B + 63) // 64, 1, 1))[:B]).cuda()
qtz = os.environ.get("QTZ", "1") != "0"
print("plan", N.encode_plan(B), "ref plan", N.encode_plan(64), "qtz", qtz)
with torch.no_grad():
    ref = m.encode_device(cfg, feat[:64].contiguous(), None, l1, l2, qtz=qtz)
    ref_r = ref.r.cpu().numpy().view(np.int32); ref_i = ref.idx.cpu().numpy()
    big = m.encode_device(cfg, feat, None, l1, l2, qtz=qtz)
    br = big.r.cpu().numpy().view(np.int32); bi = big.idx.cpu().numpy()
want_r = np.tile(ref_r, ((B + 63) // 64, 1, 1))[:B]
want_i = np.tile(ref_i, ((B + 63) // 64, 1, 1))[:B]
dif = (br != want_r).any(-1)            # (B, L)
ntile = (B + MT - 1) // MT
bad_tiles = 0
for t in range(ntile):
    dd = dif[t * MT:(t + 1) * MT]
    if not dd.any():
        continue
    bad_tiles += 1
    f0 = int(np.argmax(dd.any(0)))
    rows = np.nonzero(dd[:, f0])[0]
    if bad_tiles <= 12:
        u = t * MT + rows[0]
        ulp = (br[u, f0] - want_r[u, f0])
        print("tile %d: first diff frame %d rows %s | row %d ulp %s | idx same: %s" % (t, f0, rows.tolist(), rows[0], ulp[:6].tolist(), bool((bi[u, f0] == want_i[u, f0]).all())))
print("tiles with any difference: %d of %d" % (bad_tiles, ntile))
