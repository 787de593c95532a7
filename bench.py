#!/usr/bin/env python
"""bench.py -- coded frames/s of the closed-loop encode (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: Wavernn.encoder on U utterances x L frames per
GPU (default: BASELINE.json configs[1], 4096 utterances x 10 s = 4 096 000 frames, README
thresholds l1=0.09 l2=0.28, random-init weights + random codebooks, synthetic features).  One
process per GPU; with N > 1 (torchrun) every rank encodes its own 4096 utterances (weak scaling,
utterance ids offset by rank, no collective on the data path).

Timed regions (CUDA events on the launching stream, barrier + synchronize on both sides, max
over ranks):
  value  inputs already in HBM; the step is exactly one launch of the fused frame-step kernel.
  e2e    the public call with HOST buffers (Wavernn.encode_host -> C ABI fpc_encode_host): pinned feat -> H2D ->
         closed loop -> D2H of every output the reference returns (c_in, r, r_qtz, ind1, ind2) + the index record,
         every step.  The call cuts the utterances along time and overlaps the copies with the kernel.
The 328 MB input and 1.3 GB of outputs per step are larger than the 126 MB L2, so nothing is
served from cache between steps.

--impl reference times the reference's algorithm on the host cores: the reference itself is pure
Python (12 frames/s per process) and cannot travel to the GPU box, so this arm runs
its C restatement (oracle/, OpenMP over utterances, all host threads) on a bounded sample of the
same workload.  It is the only place besides cpu_baseline where bench.py executes oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

# stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "feature-predictor-for-speech-codec_b200")
for p in (PKG,):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "coded frames/sec closed-loop encode"
UNIT = "frames/s"
F_GRU = 1328640.0                      # FLOP per frame, predictor
F_VQ_ABOVE, F_VQ_BELOW = 313344.0, 26112.0
F_SCL_ABOVE, F_SCL_BELOW = 768.0, 48.0
HBM_BYTES_PER_FRAME = 320.0
# dram__bytes_read.sum + dram__bytes_write.sum of one fpc::encode_fp32_kernel launch over 4096 x 50 frames, ncu --set full
# (profiles/r1_final_encode_fp32_ncu_raw.txt): 42.56 MB + 25.26 MB = 331 B per coded frame
NCU_DRAM_BYTES_PER_FRAME = (42.560000e6 + 25.260800e6) / (4096 * 50)


def flops_per_frame(p1, p2):
    return F_GRU + p2 * F_VQ_ABOVE + (1 - p2) * F_VQ_BELOW + p1 * F_SCL_ABOVE + (1 - p1) * F_SCL_BELOW


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def make_inputs(S, first_utt, n_utts, n_frames):
    import numpy as np
    # 512 distinct seeded utterances per rank, tiled: generation is host-side scipy work that is not
    # part of the measurement, and the branch statistics do not depend on the tiling
    distinct = min(n_utts, 512)
    base = S.make_features(distinct, n_frames, first_utt=first_utt)
    reps = (n_utts + distinct - 1) // distinct
    return np.ascontiguousarray(np.tile(base, (reps, 1, 1))[:n_utts])


def thresholds(name):
    return (0.09, 0.28) if name == "readme" else (0.25, 2.1)


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_encode_rate(S, n_frames, l1, l2, target_seconds, first_utt=0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    sd = S.make_state_dict(0)
    w = O.weights_from_state_dict(sd)
    cbs = S.make_codebooks(0)
    C = O.Codebooks(cbs["cb_path"], cbs["scl_cb_path"], cbs["bl_cb_path"], cbs["bl_scl_cb_path"])
    # every host thread this process may use -- not OMP_NUM_THREADS, which torchrun pins to 1 for its children
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    probe = S.make_features(threads, min(n_frames, 50), first_utt=first_utt)
    t0 = time.perf_counter()
    O.encode(w, C, probe, l1, l2, nthreads=threads)
    rate = probe.shape[0] * probe.shape[1] / max(time.perf_counter() - t0, 1e-6)
    n_utts = max(threads, int(rate * target_seconds / n_frames) // threads * threads)
    feat = S.make_features(min(n_utts, 64), n_frames, first_utt=first_utt)
    import numpy as np
    feat = np.ascontiguousarray(np.tile(feat, ((n_utts + feat.shape[0] - 1) // feat.shape[0], 1, 1))[:n_utts])

    def step():
        t = time.perf_counter()
        O.encode(w, C, feat, l1, l2, nthreads=threads)
        return time.perf_counter() - t
    return step, n_utts, threads


def run_reference(args):
    import fpc_synth as S
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    l1, l2 = thresholds(args.thresholds)
    budget = 150.0 / max(1, args.steps + args.warmup)
    step, n_utts, threads = cpu_encode_rate(S, args.frames, l1, l2, min(20.0, budget))
    for _ in range(args.warmup):
        step()
    ts = [step() for _ in range(args.steps)]
    frames = n_utts * args.frames
    val = frames * len(ts) / sum(ts)
    sample = "%d utterances x %d frames per step (same generator and codebooks as the GPU arm)" % (n_utts, args.frames)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(ts) / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "thresholds": args.thresholds, "l1": l1, "l2": l2},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference (oracle/), OpenMP over utterances; the Python reference "
                                 "itself ran at 12.3 frames/s per process in the build container (BASELINE.md)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    return "closed-loop encode, %d utterances x %d frames per GPU (BASELINE.json configs[1]), fp32 predictor" % (
        args.utts, args.frames)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import fpc_native
    import fpc_synth as S
    from models.wavernn import Wavernn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the closed-loop path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l1, l2 = thresholds(args.thresholds)
    U, L = args.utts, args.frames
    model = Wavernn(20, 384, 128, 18).eval()
    model.load_state_dict(S.make_state_dict(0))
    model = model.to(dev)
    tmp = tempfile.TemporaryDirectory(prefix="fpc_bench_")
    cfg = S.save_codebooks(S.make_codebooks(0), tmp.name)
    feat_h = torch.from_numpy(make_inputs(S, rank * U, U, L)).pin_memory()
    feat_d = feat_h.to(dev, non_blocking=True)
    out = {"c_in": torch.empty((U, L, 20), device=dev), "r": torch.empty((U, L, 18), device=dev),
           "r_qtz": torch.empty((U, L, 18), device=dev), "ind1": torch.empty((U, L, 1), device=dev),
           "ind2": torch.empty((U, L, 1), device=dev), "idx": torch.empty((U, L, 4), dtype=torch.int32, device=dev)}
    host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
    stream = torch.cuda.current_stream(dev)

    def step_resident():
        return model.encode_device(cfg, feat_d, None, l1, l2, qtz=True, want_under=False, out=out)

    def step_e2e():
        # the reference-facing call with HOST buffers: upload, closed loop and download of every result inside
        # (C ABI fpc_encode_host: time-chunked, the copies overlap the kernel)
        return model.encode_host(cfg, feat_h, l1, l2, qtz=True, out=host_out, chunks=args.e2e_chunks)

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        n0 = fpc_native.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t_all0.record(stream)
        for a, b in evs:
            a.record(stream)
            res = step_resident()
            b.record(stream)
        t_all1.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        launches = fpc_native.launch_count() - n0
        clocks = sampler.stop() if rank == 0 else None
        ms_total = t_all0.elapsed_time(t_all1)
        kernel_ms = [a.elapsed_time(b) for a, b in evs]
        p1 = float(res.ind1.mean().item())
        p2 = float(res.ind2.mean().item())

        # ---- end to end through the public call with host buffers ----
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_e2e()
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        checksum = float(host_out["c_in"][0, -1].sum())   # the result really is on the host

    d2h_bytes = int(sum(v.numel() * v.element_size() for v in out.values()))
    del host_out
    bf16 = None
    if args.workload in ("both", "bf16"):
        del feat_d
        for k in list(out):
            out[k] = out[k][:0]
        torch.cuda.empty_cache()
        bf16 = bench_bf16(args, torch, dist, dev, rank, world, barrier, model, cfg, S, l1, l2)
    kmeans = None
    if args.workload in ("both", "kmeans"):
        torch.cuda.empty_cache()
        kmeans = bench_kmeans(args, torch, dist, dev, rank, world, barrier)

    t = torch.tensor([ms_total, e2e_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, wall_ms = [float(x) for x in t.tolist()]
    frames_all = float(U) * L * world
    value = frames_all * args.steps / (ms_total * 1e-3)
    e2e_value = frames_all * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        pk, pk_src = peaks()
        k_ms = sum(kernel_ms) / len(kernel_ms)
        fpf = flops_per_frame(p1, p2)
        fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12     # FFMA lanes x 2 FLOP x max SM clock
        achieved = U * L * fpf / (k_ms * 1e-3) / 1e12
        hbm_achieved = U * L * HBM_BYTES_PER_FRAME / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "utterances_per_gpu": U, "frames": L, "thresholds": args.thresholds,
                       "l1": l1, "l2": l2, "above_threshold_fraction": {"c0": p1, "c1_17": p2},
                       "l2_policy": "inputs (328 MB/step) and outputs (1.3 GB/step) larger than the 126 MB L2",
                       "parallelism": "utterance shards, %d rank(s), no collective" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(feat_h.numel() * 4),
                    "d2h_bytes_per_step": d2h_bytes,
                    "call": "Wavernn.encode_host (fpc_encode_host), chunks=%s" % (args.e2e_chunks or "auto"),
                    "checksum": checksum},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak, "traffic": NCU_DRAM_BYTES_PER_FRAME * U * L,
                         "traffic_note": "DRAM bytes of this launch scaled from the ncu capture of a 4096 x 50 frame launch "
                                         "(331 B per frame; algorithmic 320 B per frame)",
                         "kernel": "fpc::encode_fp32_kernel", "kernel_ms": k_ms,
                         "flop_per_frame": fpf,
                         "peak_source": "148 SM x 128 FFMA lanes x 2 x sm_max_mhz of MEASURED_PEAKS.json (%s); the fp32 "
                                        "predictor + direct-form VQ run on the FP32 pipe, SURVEY.md 8(d)" % pk_src,
                         "hbm": {"achieved": hbm_achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": hbm_achieved / pk["hbm_gbs"], "bytes_per_frame": HBM_BYTES_PER_FRAME}},
            "clocks": clocks,
            "wall_ms_per_step": wall_ms / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            step, n_utts, threads = cpu_encode_rate(S, L, l1, l2, args.cpu_seconds)
            dt = step()
            line["cpu_baseline"] = {
                "value": n_utts * L / dt, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "%d utterances x %d frames, one pass (%.1f s)" % (n_utts, L, dt),
                "note": "C restatement of the reference (oracle/), OpenMP over utterances; the Python reference itself "
                        "ran at 12.3 frames/s per process in the build container (BASELINE.md)"}
        if bf16 is not None:
            line["bf16"] = bf16
        if kmeans is not None:
            line["kmeans"] = kmeans
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    tmp.cleanup()



# ---------------------------------------------------------------------------------------------
# second metric: k-means iters/s (one iter = one cb_func.update over the whole residual set)
# ---------------------------------------------------------------------------------------------
def make_kmeans_shard(torch, dev, n_total, rank, world, chunk=1 << 20):
    """This rank's shard of the N x 17 synthetic residual set, generated on the device: mixture of
    2048 Gaussians (centres N(0, 0.1^2) from a Philox seed, spread 0.03).  Chunks of 2^20 vectors are
    seeded by chunk id, so the global data set does not depend on the number of ranks."""
    import numpy as np
    import fpc_dist
    centres = torch.from_numpy((np.random.Generator(np.random.Philox(key=0)).standard_normal((2048, 17)) * 0.1)
                               .astype(np.float32)).to(dev)
    first, cnt = fpc_dist.shard_range(n_total, rank, world)
    out = torch.empty((cnt, 17), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    pos = first
    while pos < first + cnt:
        c = pos // chunk
        lo, hi = c * chunk, min((c + 1) * chunk, n_total)
        g.manual_seed(1000003 + c)
        comp = torch.randint(0, 2048, (hi - lo,), generator=g, device=dev)
        block = centres[comp] + 0.03 * torch.randn((hi - lo, 17), generator=g, device=dev)
        a, b = max(lo, first), min(hi, first + cnt)
        out[a - first:b - first] = block[a - lo:b - lo]
        pos = hi
    return out


def bench_kmeans(args, torch, dist, dev, rank, world, barrier):
    import numpy as np
    import fpc_native
    from quantization import cb_func
    n_total = args.kmeans_vectors
    data = make_kmeans_shard(torch, dev, n_total, rank, world)
    res = {}
    for K in (1024, 512):
        cb0 = np.random.Generator(np.random.Philox(key=7)).standard_normal((K, 17)) * 0.1
        cb = torch.from_numpy(cb0).to(dev)
        for _ in range(2):                       # warm-up; also moves centroids off the random init
            cb, _, _ = cb_func.update_device(data, cb)
        barrier()
        stream = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = fpc_native.launch_count()
        iters = args.kmeans_iters
        e0.record(stream)
        for _ in range(iters):
            cb, stats, n_seen = cb_func.update_device(data, cb)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / iters
        flops = 3.0 * n_total * K * 17
        res["K%d" % K] = {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "gpu_launches": int(fpc_native.launch_count() - n0),
                          "fp32_direct_form_tflops": flops / (ms * 1e-3) / 1e12,
                          "hbm_gbs": n_total * 68.0 / world / (ms * 1e-3) / 1e9,
                          "empty_clusters": float(stats[2].item()), "vectors_seen": int(n_seen)}
    res["vectors"] = n_total
    res["sharding"] = "%d vectors per rank, all-reduce of (K,17) sums + (K) counts per iteration" % data.shape[0]
    res["scaling"] = "strong"
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # e2e: the reference-facing call with a HOST array (H2D of the data inside the timed region)
        nh = min(n_total, 8_000_000)
        host = data[:nh].cpu().pin_memory()
        cbh = cb.cpu().numpy()
        cb_func.update(host, cbh, 512, verbose=False)
        t0 = time.perf_counter()
        cb_func.update(host, cbh, 512, verbose=False)
        dt = time.perf_counter() - t0
        res["e2e_host_call"] = {"call": "cb_func.update(host array, K=512)", "vectors": nh, "seconds": dt,
                                "h2d_bytes": nh * 68, "iters_per_s_scaled_to_all_vectors": nh / dt / n_total}
        # the whole grow-by-one LBG schedule of cb_func.vq_train (cb_func.py:28-54): 4 (K - 1) + 10 Lloyd iterations with K
        # growing from 1 to 1024, on a bounded sample so the default run stays short
        nv = min(n_total, args.vq_train_vectors)
        if nv > 0:
            sub = data[:nv].contiguous()
            torch.cuda.synchronize()
            n0 = fpc_native.launch_count()
            t0 = time.perf_counter()
            cb_func.vq_train(sub, np.zeros((1024, 17)), 1024, rng=np.random.RandomState(0))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            res["vq_train"] = {"call": "cb_func.vq_train(data, codebook, 1024)", "vectors": nv, "updates": 4 * 1023 + 10,
                               "seconds": dt, "gpu_launches": int(fpc_native.launch_count() - n0),
                               "note": "wall clock incl. the host loop; the reference runs the same schedule with "
                                       "cb_func.update at 0.39 iters/s for N=20000, K=1024 (BASELINE.md)"}
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        ns = 200_000
        sample = data[:ns].cpu().numpy()
        t0 = time.perf_counter()
        O.kmeans_update(sample, cbh)
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"kind": "port", "cores": O.num_threads(), "sample": "%d vectors, K=512, one update" % ns,
                               "iters_per_s_scaled_to_all_vectors": ns / dt / n_total,
                               "note": "reference cb_func.update measured 0.39 iters/s at N=20000, K=1024 (BASELINE.md) "
                                       "= 1.6e-4 iters/s scaled to 50 M vectors"}
    del data
    return res


# ---------------------------------------------------------------------------------------------
# bf16 leg (BASELINE.json configs[2]): GRU gate GEMMs on tcgen05 at a 16k-utterance batch per GPU
# ---------------------------------------------------------------------------------------------
def bench_bf16(args, torch, dist, dev, rank, world, barrier, model, cfg, S, l1, l2):
    import fpc_native
    U, L = args.bf16_utts, args.frames
    base = torch.from_numpy(make_inputs(S, rank * U, min(U, 512), L)).to(dev)
    feat = base.repeat((U + base.shape[0] - 1) // base.shape[0], 1, 1)[:U].contiguous()
    out = {"c_in": torch.empty((U, L, 20), device=dev), "r": torch.empty((U, L, 18), device=dev),
           "r_qtz": torch.empty((U, L, 18), device=dev), "ind1": torch.empty((U, L, 1), device=dev),
           "ind2": torch.empty((U, L, 1), device=dev), "idx": torch.empty((U, L, 4), dtype=torch.int32, device=dev)}
    prev = model.precision
    model.precision = fpc_native.FPC_PREC_BF16
    try:
        with torch.no_grad():
            for _ in range(2):
                res = model.encode_device(cfg, feat, None, l1, l2, qtz=True, want_under=False, out=out)
            barrier()
            stream = torch.cuda.current_stream(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = fpc_native.launch_count()
            e0.record(stream)
            for _ in range(args.bf16_steps):
                res = model.encode_device(cfg, feat, None, l1, l2, qtz=True, want_under=False, out=out)
            e1.record(stream)
            barrier()
            launches = fpc_native.launch_count() - n0
            p1, p2 = float(res.ind1.mean().item()), float(res.ind2.mean().item())
            ms_res = e0.elapsed_time(e1)
            # end to end with host buffers (Wavernn.encode_host): upload, closed loop, download of every output
            shapes = {k: (tuple(v.shape), v.dtype) for k, v in out.items()}
            del out, res
            feat_h = feat.cpu().pin_memory()
            host_out = {k: torch.empty(shp, dtype=dt).pin_memory() for k, (shp, dt) in shapes.items()}
            model.encode_host(cfg, feat_h, l1, l2, qtz=True, out=host_out, chunks=args.e2e_chunks)
            barrier()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record(stream)
            for _ in range(args.bf16_steps):
                model.encode_host(cfg, feat_h, l1, l2, qtz=True, out=host_out, chunks=args.e2e_chunks)
            e3.record(stream)
            barrier()
            ms_e2e = e2.elapsed_time(e3)
            d2h = int(sum(v.numel() * v.element_size() for v in host_out.values()))
            h2d = int(feat_h.numel() * 4)
            del host_out, feat_h
    finally:
        model.precision = prev
    t = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0].item()) / args.bf16_steps
    ms_e2e = float(t[1].item()) / args.bf16_steps
    pk, _ = peaks()
    fq = flops_per_frame(p1, p2) - F_GRU
    fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    per_gpu = U * L / (ms * 1e-3)
    return {"workload": "closed-loop encode, %d utterances x %d frames per GPU (BASELINE.json configs[2]), bf16 predictor on "
                        "tcgen05, bf16 recurrent state, exact fp32 quantisers" % (U, L),
            "value": per_gpu * world, "unit": UNIT, "ms_per_step": ms, "steps": args.bf16_steps, "gpu_launches": int(launches),
            "e2e": {"value": U * L * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "call": "Wavernn.encode_host (fpc_encode_host)"},
            "above_threshold_fraction": {"c0": p1, "c1_17": p2},
            "roofline": {"bound": "fp32", "what": "quantiser work (direct-form VQ + scalar) on the FP32 pipe",
                         "achieved": per_gpu * fq / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": per_gpu * fq / 1e12 / fp32_peak, "flop_per_frame": fq},
            "tensor": {"achieved": per_gpu * F_GRU / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                       "frac": per_gpu * F_GRU / 1e12 / pk["bf16_tflops_sustained"], "flop_per_frame": F_GRU},
            "tolerance": "tests/test_gpu_bf16.py: predictor within 3e-2 abs of the fp32 oracle before the first index "
                         "divergence; decode(encode(x)) bit-exact; quantisers exact on the residual the kernel saw"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--utts", type=int, default=4096, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=1000, help="frames per utterance (10 ms each)")
    ap.add_argument("--thresholds", choices=("readme", "calibrated"), default="readme")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", choices=("both", "encode", "kmeans", "bf16"), default="both",
                    help="the fp32 encode is always the headline line; 'both' appends the bf16 and k-means objects")
    ap.add_argument("--bf16-utts", type=int, default=16384)
    ap.add_argument("--bf16-steps", type=int, default=3)
    ap.add_argument("--kmeans-vectors", type=int, default=50_000_000, help="total residual vectors (all ranks)")
    ap.add_argument("--kmeans-iters", type=int, default=3)
    ap.add_argument("--vq-train-vectors", type=int, default=2_000_000, help="sample for the full vq_train schedule (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--e2e-chunks", type=int, default=0, help="frame ranges of the host-buffer call (0 = library default)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
