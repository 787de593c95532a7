import os, sys, numpy as np, torch, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feature-predictor-for-speech-codec_b200")); sys.path.insert(0, ROOT)
import fpc_native as N, fpc_synth as S
from models.wavernn import Wavernn
m = Wavernn(20, 384, 128, 18).eval(); m.load_state_dict(S.make_state_dict(0)); m = m.cuda()
d = tempfile.mkdtemp(); cfg = S.save_codebooks(S.make_codebooks(0), d)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
l1, l2 = 0.25, 2.1
base = S.make_features(64, L)
feat = torch.from_numpy(np.tile(base, ((B + 63) // 64, 1, 1))[:B]).cuda()
print('plan', N.encode_plan(B) if hasattr(N, 'encode_plan') else None)
with torch.no_grad():
    m.encode_device(cfg, feat, None, l1, l2); ref = m.last_result.idx.clone() if hasattr(m, "last_result") and m.last_result is not None else None
    r1 = m.encode_device(cfg, feat, None, l1, l2)
    idx1 = r1.idx.cpu().numpy()
    rr = r1.r.cpu().numpy().view(np.int32)
    rg0 = rr[:64]
    # the same 64 utterances repeat: every group of 64 must be identical
    g0 = idx1[:64]
    bad = 0
    for s in range(64, B - 63, 64):
        ne = np.argwhere((idx1[s:s + 64] != g0).any(-1))
        if len(ne):
            bad += 1
            if bad <= 6:
                u, f = ne[0]
                dr = np.argwhere((rr[s:s + 64][u] != rg0[u]).any(-1))
                f0 = dr[0][0] if len(dr) else -1
                if f0 >= 0:
                    a, b = rr[s + u, f0], rg0[u, f0]
                    print("   residual first differs at frame", f0, "coeffs", np.nonzero(a != b)[0], "ulp diff", (a - b)[a != b][:6], "values", r1.r[s + u, f0].cpu().numpy()[(a != b)][:3])
                print("group", s // 64, "first diff utt", s + u, "tile", (s + u) // 28, "row", (s + u) % 28, "frame", f, "got", idx1[s + u, f], "want", g0[u, f], "ndiff-utts", len(set(ne[:, 0])))
    print("groups differing from group 0:", bad, "of", B // 64 - 1)
    # small run of group 0 alone
    r2 = m.encode_device(cfg, feat[:64].contiguous(), None, l1, l2)
    idx2 = r2.idx.cpu().numpy()
    ne = np.argwhere((idx2 != g0).any(-1))
    print("group 0 big vs alone: diffs", len(ne), (ne[0], idx2[ne[0][0], ne[0][1]], g0[ne[0][0], ne[0][1]]) if len(ne) else "")
