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

Every line also carries `scaling100k`: BASELINE.json configs[4], the closed-loop encode of --total-utts (100 000)
utterances x L frames sharded by utterance over the N ranks (strong scaling, no collective, features generated on the
device from per-chunk Philox seeds so the data set does not depend on N), with a digest of the index record and
decoded features of 64 fixed utterance ids: equal digests across the N = 1/2/4/8 lines -- and against rank 0 encoding
those 64 utterances alone -- show that the sharded result is bit-identical.  `calibrated` repeats the headline with
thresholds that put about half of the frames below threshold (both codebook branches are then timed).

--impl reference times the reference on the host cores: the UNMODIFIED Python reference staged under
baseline/_ref (oracle/stage_ref.py), one process per host core through oracle/ref_worker.py (`kind: "reference"`), and
beside it the C restatement (oracle/, OpenMP over utterances, all host threads; `port`), each on a bounded sample of
the same workload.  When no staged reference travelled with the tree the port alone is reported (`kind: "port"`).
These legs are the only places where bench.py executes oracle/.
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
# (profiles/r2_encode_fp32_ncu_raw.txt, the role-split kernel of round 2): 41.46 MB + 20.99 MB = 305 B per coded frame
# (round 1: 331; a launch this short leaves part of its 248 B/frame of output in the 126 MB L2)
NCU_DRAM_BYTES_PER_FRAME = (41.458688e6 + 20.985600e6) / (4096 * 50)


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


def make_features_device(torch, dev, first, count, n_frames, chunk=1024):
    """(count, n_frames, 20) features of utterances [first, first + count) generated on the device.  Same process as
    fpc_synth.make_features (AR(1) cepstra with rho 0.95 and stationary std 0.3 / 0.1, piecewise-constant pitch), but
    drawn from torch's Philox generator seeded per CHUNK of 1024 utterance ids, so utterance u has the same features
    whatever range, rank or GPU count it is generated for (100 000 distinct utterances; the scipy generator would
    need minutes for them)."""
    out = torch.empty((count, n_frames, 20), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    rho = 0.95
    std = torch.full((18,), 0.1, device=dev)
    std[0] = 0.3
    sig = std * (1.0 - rho * rho) ** 0.5
    nseg = (n_frames + 19) // 20
    pos = first
    while pos < first + count:
        c = pos // chunk
        lo = c * chunk
        g.manual_seed(7_000_003 + c)
        eps = torch.randn((chunk, n_frames, 18), generator=g, device=dev)
        seg = torch.rand((chunk, nseg, 2), generator=g, device=dev) * 2.0 - 1.0
        a, b = max(lo, first), min(lo + chunk, first + count)
        blk = out[a - first:b - first]
        blk[:, :, :18] = eps[a - lo:b - lo] * sig
        blk[:, 0, :18] = eps[a - lo:b - lo, 0] * std
        blk[:, :, 18:] = seg[a - lo:b - lo].repeat_interleave(20, dim=1)[:, :n_frames]
        pos = lo + chunk
    for t in range(1, n_frames):
        out[:, t, :18].add_(out[:, t - 1, :18], alpha=rho)
    return out


def host_threads():
    # every host thread this process may use -- not OMP_NUM_THREADS, which torchrun pins to 1 for its children
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def staged_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import stage_ref
    return stage_ref.stage()        # copies from /root/reference when present (build container); else what travelled


def python_reference_encode(n_frames, l1, l2, steps, warmup, seconds_per_step, max_procs=32):
    """The unmodified Python reference (baseline/_ref), one single-threaded process per host core, each on its own
    utterance.  Returns (frames/s over all processes per step list, description dict) or None."""
    src = staged_reference()
    if src is None:
        return None
    procs = max(1, min(host_threads(), max_procs))
    env = dict(os.environ, FPC_REFERENCE_SRC=src, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_worker.py"), "encode"]
    ps = [subprocess.Popen(cmd + [str(k), str(n_frames), repr(l1), repr(l2), str(steps), str(warmup), repr(seconds_per_step)],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for k in range(procs)]
    outs = []
    for pr in ps:
        try:
            o, e = pr.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            pr.kill()
            return None
        if pr.returncode != 0:
            sys.stderr.write("ref_worker failed: %s\n" % e[-400:])
            return None
        outs.append(json.loads(o.strip().splitlines()[-1]))
    # the processes run side by side: a step's rate is the frames all of them coded over the slowest one's time
    rates = []
    for k in range(steps):
        rates.append(sum(o["frames_per_step"] for o in outs) / max(o["seconds"][k] for o in outs))
    frames = [o["frames_per_step"] for o in outs]
    return rates, {"cores": procs, "frames_per_process_per_step": [min(frames), max(frames)], "frames_total": sum(frames),
                   "per_process_frames_per_s": sum(o["probe_frames_per_s"] for o in outs) / len(outs), "src": os.path.relpath(src, ROOT)}


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
    threads = host_threads()
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
    nsteps = max(1, args.steps + args.warmup)
    # the C port first (short), then the real reference with what is left of a ~4 minute budget
    step, n_utts, threads = cpu_encode_rate(S, args.frames, l1, l2, min(10.0, 60.0 / nsteps))
    for _ in range(args.warmup):
        step()
    ts = [step() for _ in range(args.steps)]
    port_val = n_utts * args.frames * len(ts) / sum(ts)
    port = {"value": port_val, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": 1e3 * sum(ts) / len(ts),
            "sample": "%d utterances x %d frames per step (same generator and codebooks as the GPU arm)" % (n_utts, args.frames),
            "note": "C restatement of the reference (oracle/), OpenMP over utterances"}
    ref = python_reference_encode(args.frames, l1, l2, args.steps, args.warmup, min(8.0, 150.0 / nsteps))
    if ref is not None:
        rates, info = ref
        val = len(rates) / sum(1.0 / r for r in rates)        # frames over total time of the timed steps
        fr = info["frames_per_process_per_step"]
        cpu = {"value": val, "unit": UNIT, "cores": info["cores"], "kind": "reference",
               "sample": "%d processes (one per host core, torch.set_num_threads(1)), each the first %d-%d frames of its own "
                         "utterance per step" % (info["cores"], fr[0], fr[1]),
               "note": "UNMODIFIED Python reference (%s, sha256 manifest beside it): Wavernn.encoder with its own vq_quantize / "
                       "scl_quantize, codebooks np.load-ed per call" % info["src"],
               "per_process_frames_per_s": info["per_process_frames_per_s"], "port": port}
        ms = 1e3 * info["frames_total"] / val
    else:
        val, cpu, ms = port_val, dict(port), port["ms_per_step"]
        cpu["note"] += "; no staged Python reference travelled with the tree (baseline/_ref absent)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "thresholds": args.thresholds, "l1": l1, "l2": l2},
        "cpu_baseline": cpu,
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
        return model.encode_host(cfg, feat_h, l1, l2, qtz=True, out=host_out, chunks=args.e2e_chunks, want_hist=True)

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
            hres = step_e2e()
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        checksum = float(host_out["c_in"][0, -1].sum())   # the result really is on the host
        cb_tot = Wavernn.host_cb_tot(hres)                # ... and so is the 7th element of the reference's tuple
        e2e_hist_frames = [float(np.sum(h)) for h in cb_tot]

        # ---- the same step at calibrated thresholds: about half of the frames take the below-threshold books ----
        calibrated = None
        if args.thresholds == "readme" and not args.no_calibrated:
            cl1, cl2 = thresholds("calibrated")
            for _ in range(2):
                model.encode_device(cfg, feat_d, None, cl1, cl2, qtz=True, want_under=False, out=out)
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for _ in range(args.calibrated_steps):
                cres = model.encode_device(cfg, feat_d, None, cl1, cl2, qtz=True, want_under=False, out=out)
            c1.record(stream)
            barrier()
            calibrated = {"ms": c0.elapsed_time(c1) / args.calibrated_steps, "l1": cl1, "l2": cl2,
                          "p1": float(cres.ind1.mean().item()), "p2": float(cres.ind2.mean().item())}

    d2h_bytes = int(sum(v.numel() * v.element_size() for v in out.values())) + 8 * fpc_native.HIST_TOTAL
    del host_out
    scaling = None
    feat_d = res = cres = None          # free the headline's device buffers before the other workloads
    out.clear()
    model.last_result = None
    torch.cuda.empty_cache()
    if args.total_utts > 0:
        scaling = bench_scaling(args, torch, dist, dev, rank, world, barrier, model, cfg, l1, l2)
    bf16 = None
    if args.workload in ("both", "bf16"):
        bf16 = bench_bf16(args, torch, dist, dev, rank, world, barrier, model, cfg, S, l1, l2)
    kmeans = None
    if args.workload in ("both", "kmeans"):
        torch.cuda.empty_cache()
        kmeans = bench_kmeans(args, torch, dist, dev, rank, world, barrier)

    t = torch.tensor([ms_total, e2e_ms, wall * 1e3, calibrated["ms"] if calibrated else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, wall_ms, cal_ms = [float(x) for x in t.tolist()]
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
                       "inputs": "%d distinct seeded utterances per rank (fpc_synth.make_features, seed 1000 + id), tiled %dx to "
                                 "fill the batch" % (min(U, 512), (U + min(U, 512) - 1) // min(U, 512)),
                       "launch_plan": fpc_native.encode_plan(U, fpc_native.FPC_PREC_FP32),
                       "l2_policy": "inputs (328 MB/step) and outputs (1.3 GB/step) larger than the 126 MB L2",
                       "parallelism": "utterance shards, %d rank(s), no collective" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(feat_h.numel() * 4),
                    "d2h_bytes_per_step": d2h_bytes,
                    "call": "Wavernn.encode_host (fpc_encode_host), chunks=%s; returns c_in, r, r_qtz, ind1, ind2, the index "
                            "record and cb_tot" % (args.e2e_chunks or "auto"),
                    "checksum": checksum, "cb_tot_frames": e2e_hist_frames},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak, "traffic": NCU_DRAM_BYTES_PER_FRAME * U * L,
                         "traffic_note": "DRAM bytes of this launch scaled from the ncu capture of a 4096 x 50 frame launch "
                                         "(305 B per frame, part of the output still in L2; algorithmic 320 B per frame)",
                         "kernel": "fpc::encode_fp32_kernel", "kernel_ms": k_ms,
                         "flop_per_frame": fpf,
                         "peak_source": "148 SM x 128 FFMA lanes x 2 x sm_max_mhz of MEASURED_PEAKS.json (%s); the fp32 "
                                        "predictor + direct-form VQ run on the FP32 pipe, SURVEY.md 8(d)" % pk_src,
                         "hbm": {"achieved": hbm_achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": hbm_achieved / pk["hbm_gbs"], "bytes_per_frame": HBM_BYTES_PER_FRAME}},
            "clocks": clocks,
            "wall_ms_per_step": wall_ms / args.steps,
        }
        if calibrated is not None:
            cf = flops_per_frame(calibrated["p1"], calibrated["p2"])
            cach = U * L * cf / (cal_ms * 1e-3) / 1e12
            line["calibrated"] = {
                "workload": "the headline step at calibrated thresholds l1=%g l2=%g (SURVEY.md 8d)" % (calibrated["l1"], calibrated["l2"]),
                "value": U * L * world / (cal_ms * 1e-3), "unit": UNIT, "ms_per_step": cal_ms, "steps": args.calibrated_steps,
                "above_threshold_fraction": {"c0": calibrated["p1"], "c1_17": calibrated["p2"]},
                "roofline": {"bound": "fp32", "achieved": cach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": cach / fp32_peak,
                             "flop_per_frame": cf}}
        if scaling is not None:
            line["scaling100k"] = scaling
        if world == 1 and not args.no_cpu_baseline:
            step, n_utts, threads = cpu_encode_rate(S, L, l1, l2, args.cpu_seconds)
            dt = step()
            port = {"value": n_utts * L / dt, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": "%d utterances x %d frames, one pass (%.1f s)" % (n_utts, L, dt),
                    "note": "C restatement of the reference (oracle/), OpenMP over utterances"}
            ref = python_reference_encode(L, l1, l2, 1, 0, args.cpu_seconds)
            if ref is not None:
                rates, info = ref
                fr = info["frames_per_process_per_step"]
                line["cpu_baseline"] = {
                    "value": rates[0], "unit": UNIT, "cores": info["cores"], "kind": "reference",
                    "sample": "%d processes (one per host core, torch.set_num_threads(1)), each the first %d-%d frames of its own "
                              "utterance, one pass" % (info["cores"], fr[0], fr[1]),
                    "note": "UNMODIFIED Python reference (%s): Wavernn.encoder with its own vq_quantize / scl_quantize, codebooks "
                            "np.load-ed per call" % info["src"],
                    "per_process_frames_per_s": info["per_process_frames_per_s"], "port": port}
            else:
                port["note"] += "; no staged Python reference travelled with the tree (baseline/_ref absent)"
                line["cpu_baseline"] = port
        if bf16 is not None:
            line["bf16"] = bf16
        if kmeans is not None:
            line["kmeans"] = kmeans
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    tmp.cleanup()



# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: >= 100k utterances sharded by utterance over the ranks (strong scaling)
# ---------------------------------------------------------------------------------------------
DIGEST_IDS = 64


def _utt_digest(idx_u, cin_u):
    import hashlib
    h = hashlib.sha256()
    h.update(idx_u.contiguous().cpu().numpy().tobytes())
    h.update(cin_u.contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


def bench_scaling(args, torch, dist, dev, rank, world, barrier, model, cfg, l1, l2):
    import hashlib
    import fpc_dist
    import fpc_native
    total, L = args.total_utts, args.frames
    first, count = fpc_dist.shard_range(total, rank, world)
    feat = make_features_device(torch, dev, first, count, L)
    out = {"c_in": torch.empty((count, L, 20), device=dev), "r": torch.empty((count, L, 18), device=dev),
           "r_qtz": torch.empty((count, L, 18), device=dev), "ind1": torch.empty((count, L, 1), device=dev),
           "ind2": torch.empty((count, L, 1), device=dev), "idx": torch.empty((count, L, 4), dtype=torch.int32, device=dev)}
    stream = torch.cuda.current_stream(dev)
    with torch.no_grad():
        for _ in range(args.scaling_warmup):
            res = model.encode_device(cfg, feat, None, l1, l2, qtz=True, want_under=False, out=out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = fpc_native.launch_count()
        e0.record(stream)
        for _ in range(args.scaling_steps):
            res = model.encode_device(cfg, feat, None, l1, l2, qtz=True, want_under=False, out=out)
        e1.record(stream)
        barrier()
        launches = fpc_native.launch_count() - n0
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.scaling_steps
        # digest of fixed utterance ids, gathered from the ranks that own them
        ids = [7 + k * (total // DIGEST_IDS) for k in range(DIGEST_IDS)] if total >= 8 * DIGEST_IDS else list(range(total))
        mine = {u: _utt_digest(res.idx[u - first], res.c_in[u - first]) for u in ids if first <= u < first + count}
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, mine)
        else:
            parts = [mine]
        p2 = float(res.ind2.mean().item())
        del res
        model.last_result = None
        single = None
        if rank == 0:
            merged = {}
            for part in parts:
                merged.update(part)
            digest = hashlib.sha256("".join(merged[u] for u in ids).encode()).hexdigest()
            # the same ids encoded ALONE by this rank (one small batch: other tiles, other plan)
            sub = torch.cat([make_features_device(torch, dev, u, 1, L) for u in ids])
            r1 = model.encode_device(cfg, sub, None, l1, l2, qtz=True, want_under=False)
            torch.cuda.synchronize()
            alone = {u: _utt_digest(r1.idx[k], r1.c_in[k]) for k, u in enumerate(ids)}
            single = hashlib.sha256("".join(alone[u] for u in ids).encode()).hexdigest()
    del feat, out
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"workload": "closed-loop encode, %d utterances x %d frames in total (BASELINE.json configs[4]), sharded by utterance "
                        "over %d rank(s), fp32 predictor" % (total, L, world),
            "scaling": "strong", "value": total * L / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": args.scaling_steps,
            "warmup": args.scaling_warmup, "utterances_total": total, "utterances_rank0": count, "gpu_launches": int(launches),
            "launch_plan_rank0": fpc_native.encode_plan(count, fpc_native.FPC_PREC_FP32),
            "above_threshold_fraction_c1_17_rank0": p2,
            "inputs": "device Philox generator seeded per chunk of 1024 utterance ids (bench.make_features_device): utterance u "
                      "is the same whatever the rank count; inputs resident in HBM, no e2e leg for this workload",
            "digest": {"ids": "%d utterance ids %d + k * %d" % (len(ids), ids[0], (ids[1] - ids[0]) if len(ids) > 1 else 0),
                       "sharded": digest, "encoded_alone_on_rank0": single, "bit_identical": digest == single,
                       "what": "sha256 over the (L,4) int32 index record and the (L,20) float32 decoded features of each id"}}


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
        pk, _ = peaks()
        kp = (K + 127) // 128 * 128
        res["K%d" % K] = {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "gpu_launches": int(fpc_native.launch_count() - n0),
                          "roofline": {"bound": "tensor", "kernel": "fpc::kmeans_assign_tc_kernel",
                                       "what": "distance screen as a tcgen05 GEMM (vectors x centroids, fp16-pair operands, K = 64); ncu (profiles/r2_kmeans_tc_ncu_raw.txt): tensor pipe 33 % active, the ALU pipe of the accumulator scan 67 % -- that is the pipe that binds",
                                       "achieved": 2.0 * n_total / world * K * 17 / (ms * 1e-3) / 1e12, "peak": pk["bf16_tflops_sustained"],
                                       "unit": "TFLOP/s", "frac": 2.0 * n_total / world * K * 17 / (ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                                       "per": "GPU (the iteration includes the all-reduce and the finalize)",
                                       "algorithmic_flop_per_iter": 2.0 * n_total * K * 17,
                                       "executed_tflops": 2.0 * n_total / world * kp * 64 / (ms * 1e-3) / 1e12},
                          "fp32_direct_form_tflops": flops / (ms * 1e-3) / 1e12,
                          "hbm_gbs": n_total * 68.0 / world / (ms * 1e-3) / 1e9,
                          "empty_clusters": float(stats[2].item()), "vectors_seen": int(n_seen)}
        if K == 1024:
            # opt-in exact mode: sums in data order like cb_func.py:82-86 (fpc_kmeans_accumulate_ordered) -- the codebook
            # of the reference bit for bit on one GPU; a stable counting sort of the rows + one warp per centroid
            co = cb
            for _ in range(2):
                co, _, _ = cb_func.update_device(data, co, ordered=True)
            barrier()
            e0.record(stream)
            for _ in range(iters):
                co, _, _ = cb_func.update_device(data, co, ordered=True)
            e1.record(stream)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res["K1024"]["ordered_sums"] = {"ms_per_iter": float(t.item()) / iters, "iters_per_s": 1e3 * iters / float(t.item()),
                                            "what": "update_device(ordered=True): index-only assignment + sums in data order (bit-exact vs the reference on one GPU)"}
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
            "roofline": {"bound": "tensor",
                         "what": "gate GEMMs and the VQ distance screen both run on tcgen05; achieved = ALGORITHMIC FLOP per frame "
                                 "(1 328 640 predictor + direct-form quantiser, SURVEY.md 8d) / time, peak = sustained bf16 of "
                                 "MEASURED_PEAKS.json",
                         "achieved": per_gpu * (F_GRU + fq) / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": per_gpu * (F_GRU + fq) / 1e12 / pk["bf16_tflops_sustained"], "flop_per_frame": F_GRU + fq,
                         "executed_mma_flop_per_frame": F_GRU + 16 * 4 * (1 + 3) * 262144.0 / 64 * p2,
                         "executed_tflops": per_gpu * (F_GRU + 16 * 4 * (1 + 3) * 262144.0 / 64 * p2) / 1e12,
                         "ncu": "profiles/r2_encode_bf16_ncu_raw.txt: tensor pipe 13.0 % active, ALU pipe 25.5 %"},
            "fp32_equivalent": {"what": "round-1 yardstick: the quantiser's direct-form FLOPs against the FP32 pipe it used to run on",
                                "achieved": per_gpu * fq / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                                "frac": per_gpu * fq / 1e12 / fp32_peak, "flop_per_frame": fq},
            "tolerance": "tests/test_gpu_bf16.py (asserted): predictor within 3e-3 abs of the fp32 oracle before the first index "
                         "divergence, index agreement with the fp32 oracle >= 0.95 (calibrated) / 0.93 (README thresholds) of "
                         "frames, decoded features within 0.05 rms; decode(encode(x)) bit-exact; quantisers exact on the residual "
                         "the kernel saw"}


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
    ap.add_argument("--total-utts", type=int, default=100_000,
                    help="utterances of the strong-scaling leg, sharded over the ranks (BASELINE.json configs[4]); 0 = skip")
    ap.add_argument("--scaling-steps", type=int, default=2)
    ap.add_argument("--scaling-warmup", type=int, default=1)
    ap.add_argument("--no-calibrated", action="store_true")
    ap.add_argument("--calibrated-steps", type=int, default=3)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
