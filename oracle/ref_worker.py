"""One worker process of the CPU reference timing (measurement infrastructure; run by bench.py only).

    python oracle/ref_worker.py encode <first_utt> <frames> <l1> <l2> <steps> <warmup> <seconds_per_step>
    python oracle/ref_worker.py kmeans <n_vectors> <K>

`encode` runs the UNMODIFIED reference Wavernn.encoder (models/wavernn.py:165-256 with the reference's own
vq_quantize / scl_quantize injected, codebooks np.load-ed per call as the reference does) on ONE synthetic utterance
on ONE core (torch.set_num_threads(1)); bench.py starts one worker per host core, each with its own utterance id
(SURVEY.md 8d "CPU baseline timing").  The reference codes ~10-80 frames/s per process, so a step is a bounded
prefix of the utterance: after a 12-frame probe the worker picks the prefix length that takes about
<seconds_per_step>.  Prints one JSON line.  The reference tree is the staged copy (FPC_REFERENCE_SRC =
baseline/_ref/src, oracle/stage_ref.py) or /root/reference/src in the build container.
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "feature-predictor-for-speech-codec_b200"))


def main():
    mode = sys.argv[1]
    import numpy as np
    import torch
    torch.set_num_threads(1)
    import fpc_synth as S
    import ref_shim
    W, vq_func, cb_func = ref_shim.load_reference()
    if mode == "kmeans":
        n, K = int(sys.argv[2]), int(sys.argv[3])
        data = S.make_kmeans_data(n, seed=0)
        cb = np.random.Generator(np.random.Philox(key=7)).standard_normal((K, 17)) * 0.1
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            cb_func.update(data, cb, K)
            dt = time.perf_counter() - t0
        print(json.dumps({"mode": "kmeans", "vectors": n, "K": K, "seconds": dt}), flush=True)
        return
    first, frames, l1, l2 = int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5])
    steps, warmup, per_step = int(sys.argv[6]), int(sys.argv[7]), float(sys.argv[8])
    model = W.Wavernn(20, S.GRU1, S.GRU2, S.N_CEPS).eval()
    model.load_state_dict(S.make_state_dict(0))
    feat = torch.tensor(S.make_features(1, frames, first_utt=first))
    tmp = tempfile.mkdtemp(prefix="fpc_refw_")
    cfg = S.save_codebooks(S.make_codebooks(0), tmp)

    def run(n):
        t0 = time.perf_counter()
        with torch.no_grad():
            model.encoder(cfg, feat[:, :n], None, l1, l2, vq_func.vq_quantize, vq_func.scl_quantize, True)
        return time.perf_counter() - t0

    probe = min(12, frames)
    rate = probe / run(probe)
    n = max(4, min(frames, int(rate * per_step)))
    for _ in range(warmup):
        run(n)
    ts = [run(n) for _ in range(steps)]
    print(json.dumps({"mode": "encode", "first_utt": first, "frames_per_step": n, "seconds": ts,
                      "probe_frames_per_s": rate}), flush=True)


if __name__ == "__main__":
    main()
