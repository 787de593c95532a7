#!/usr/bin/env python
"""Times the pieces of an ordered k-means iteration (cb_func.update_device(ordered=True)) with CUDA events:
the index-only assignment, fpc_kmeans_accumulate_ordered, and the default assign+atomics pass.
    python tools/kmeans_ordered_probe.py [n_vectors] [K]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feature-predictor-for-speech-codec_b200"))
import fpc_native as N  # noqa: E402
from quantization import cb_func  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
g = torch.Generator(device="cuda").manual_seed(0)
centres = torch.randn((2048, 17), generator=g, device="cuda") * 0.1
data = torch.empty((n, 17), device="cuda", dtype=torch.float32)
for s in range(0, n, 1 << 22):
    e = min(n, s + (1 << 22))
    comp = torch.randint(0, 2048, (e - s,), generator=g, device="cuda")
    data[s:e] = centres[comp] + torch.randn((e - s, 17), generator=g, device="cuda") * 0.03
cb = centres[:K].double().contiguous()
dev = data.device
L = N.lib()
st = N.current_stream(dev)
ws = torch.empty(max(L.fpc_kmeans_workspace_bytes(n, K), 1), dtype=torch.uint8, device=dev)
ows = torch.empty(L.fpc_kmeans_ordered_workspace_bytes(n, K), dtype=torch.uint8, device=dev)
idx = torch.empty(n, dtype=torch.int32, device=dev)
acc = torch.zeros(K * 18, dtype=torch.float64, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def assign_idx():
    N.check(L.fpc_kmeans_assign_accumulate(data.data_ptr(), n, cb.data_ptr(), K, None, None, idx.data_ptr(), ws.data_ptr(), ws.numel(), st), "assign")


def assign_sums():
    N.check(L.fpc_kmeans_assign_accumulate(data.data_ptr(), n, cb.data_ptr(), K, acc.data_ptr(), acc.data_ptr() + K * 17 * 8, None, ws.data_ptr(), ws.numel(), st), "assign")


def ordered():
    N.check(L.fpc_kmeans_accumulate_ordered(data.data_ptr(), 0, n, idx.data_ptr(), K, acc.data_ptr(), acc.data_ptr() + K * 17 * 8, ows.data_ptr(), ows.numel(), st), "ordered")


print("n = %d, K = %d" % (n, K))
print("assign, indices only      %.3f ms" % timed(assign_idx))
print("assign + float64 atomics  %.3f ms" % timed(assign_sums))
t = timed(ordered)
print("accumulate_ordered        %.3f ms  (%.0f GB/s of 84 B/vector)" % (t, n * 84 / t / 1e6))
sizes = torch.bincount(idx.long(), minlength=K)
print("centroid sizes: min %d max %d" % (int(sizes.min()), int(sizes.max())))
