import sys, os
sys.path.insert(0, "/root/repo/feature-predictor-for-speech-codec_b200")
import numpy as np, torch
import fpc_synth as S
from quantization import cb_func
n = 2_000_000
d = torch.from_numpy(S.make_kmeans_data(n, seed=0)).cuda()
for K in (1, 2, 4, 8, 16, 32, 33, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024):
    cb = torch.from_numpy(np.random.RandomState(1).randn(K, 17) * 0.1).cuda()
    for _ in range(3): cb, _, _ = cb_func.update_device(d, cb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): cb, _, _ = cb_func.update_device(d, cb)
    e1.record(); torch.cuda.synchronize()
    print("K=%4d  %.3f ms per update" % (K, e0.elapsed_time(e1) / 10))
