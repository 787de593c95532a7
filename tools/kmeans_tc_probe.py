"""Timing / statistics probe of the k-means assignment kernels on the device (development aid).

    python tools/kmeans_tc_probe.py [n_vectors]

Prints, for K = 1024, 512 and 256: ms per fpc_kmeans_assign_accumulate with the sums and for indices only, and the share
of vectors the tensor-core screen handed to its exact fallback.  FPC_KMEANS_TC=0 selects the CUDA-core screen."""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feature-predictor-for-speech-codec_b200"))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import fpc_native as N
from bench import make_kmeans_shard


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    dev = torch.device("cuda", 0)
    data = make_kmeans_shard(torch, dev, n, 0, 1)
    L = N.lib()
    st = N.current_stream(dev)
    for K in (1024, 512, 256):
        cb = torch.from_numpy(np.random.Generator(np.random.Philox(key=7)).standard_normal((K, 17)) * 0.1).to(dev)
        sums = torch.zeros((K, 17), dtype=torch.float64, device=dev)
        counts = torch.zeros((K,), dtype=torch.float64, device=dev)
        idx = torch.empty((n,), dtype=torch.int32, device=dev)
        need = L.fpc_kmeans_workspace_bytes(n, K)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        for _ in range(2):          # move the centroids off the random init (two Lloyd steps)
            sums.zero_(); counts.zero_()
            N.check(L.fpc_kmeans_assign_accumulate(data.data_ptr(), n, cb.data_ptr(), K, sums.data_ptr(), counts.data_ptr(), None,
                                                   ws.data_ptr(), need, st), "assign")
            cb = (sums / (counts[:, None] + 1e-20)).contiguous()
        for what, s_ptr, c_ptr, i_ptr in (("sums", sums.data_ptr(), counts.data_ptr(), None), ("indices only", None, None, idx.data_ptr())):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                N.check(L.fpc_kmeans_assign_accumulate(data.data_ptr(), n, cb.data_ptr(), K, s_ptr, c_ptr, i_ptr, ws.data_ptr(), need, st), "assign")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            rows = int(os.environ.get("FPC_KMEANS_REPL_ROWS", "512"))
            off = 0 if K >= rows else ((-(-rows // K)) * K * 18 * 8 + 255) // 256 * 256
            beta, cmax, k_, kp, nfb, _ = struct.unpack("ffiiII", bytes(ws[off:off + 24].cpu().numpy()))
            prof = struct.unpack("8Q", bytes(ws[off + 24:off + 88].cpu().numpy()))
            if prof[7]:
                print("      cycles per tile (CTA 0, thread 0): scan %d, wait MMA %d, finish %d, barrier A %d, decide %d, barrier B %d, "
                      "sums %d  (%d tiles)" % tuple([p // prof[7] for p in prof[:7]] + [prof[7]]))
            print("K=%4d %-13s %.2f ms  (TC=%s) fallback vectors %d = %.4f %%  beta=%g cmax=%.3f" % (
                K, what, ms, os.environ.get("FPC_KMEANS_TC", "1"), nfb, 100.0 * nfb / n, beta, cmax), flush=True)


if __name__ == "__main__":
    main()
