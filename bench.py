#!/usr/bin/env python
"""bench.py — frames/sec of SeqFormer-R50 inference on synthetic clips (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload 720p|480p]

One "step" = one clip of T=5 frames through the whole hot path (normalise/pad -> ResNet-50 ->
Deformable-DETR encoder/decoder with MSDeformAttn -> heads -> dynamic mask head -> video-level
post-processing).  `value` = frames/s with the uint8 clip already resident in HBM (CUDA-graph replay,
CUDA-event timed, max over ranks); `e2e` = frames/s through the reference-facing call
`SeqFormer.forward(batched_inputs)` with host frames in, host masks out, every step.
Multi-GPU: clips are independent — each rank runs its own clips, no collective on the data
path; weak scaling.  `--impl reference` times the CPU oracle port of the same path on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {"720p": (5, 720, 1280), "480p": (5, 480, 864), "swinl_720p": (5, 720, 1280),
             "idol_train": (8, 480, 864),          # BASELINE configs[2]: 4 videos x (key, ref) frames per GPU
             "seqformer_train": (10, 480, 864)}    # 2 clips x 5 frames per GPU (INPUT.SAMPLING_FRAME_NUM: 5)
METRIC = "frames/sec SeqFormer-R50 T=5 720p synthetic"


def env_int(name, default):
    return int(os.environ.get(name, default))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


def make_weights(seed=0, swin=None):
    from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
    from vnext_b200.synth import fill_state_dict
    return fill_state_dict(seqformer_r50_state_spec(swin=swin), seed)


def make_clip_u8(T, H, W, seed=0):
    from vnext_b200.synth import synthetic_clip
    return torch.stack(synthetic_clip(T, H, W, seed)).to(torch.uint8)


# ---------------------------------------------------------------------------------------------- reference arm
def cpu_oracle_fps(T, H, W, steps, warmup, budget_s=240.0):
    """The reference's CPU path for this workload = the oracle port (plain PyTorch, grid_sample MSDA),
    all host threads.  Bounded: the per-step sample shrinks to fewer frames when a full clip would not
    fit `budget_s` for steps+warmup runs.  Returns (frames/s, description, cores)."""
    from oracle import seqformer_oracle as SO
    cores = min(os.cpu_count() or 1, 32)        # beyond ~32 threads the small ATen ops of the path stop scaling
    torch.set_num_threads(cores)
    sd = make_weights(0)
    frames = [f.float() for f in make_clip_u8(T, H, W, 0)]
    t_frames = T
    with torch.no_grad():
        t0 = time.perf_counter()
        out = SO.inference(sd, frames[:1])
        SO.whole_video_postprocess(out, (H, W))
        t1 = time.perf_counter() - t0                       # one-frame clip (also warms the allocator)
        est_clip = t1 * T * 0.9
        if est_clip * (steps + warmup) > budget_s:
            t_frames = max(1, min(T, int(budget_s / ((steps + warmup) * t1))))
        for _ in range(max(0, warmup - 1)):
            SO.whole_video_postprocess(SO.inference(sd, frames[:t_frames]), (H, W))
        t0 = time.perf_counter()
        for _ in range(steps):
            SO.whole_video_postprocess(SO.inference(sd, frames[:t_frames]), (H, W))
        dt = time.perf_counter() - t0
    fps = steps * t_frames / dt
    desc = (f"oracle port (oracle/seqformer_oracle.py, torch {torch.__version__} CPU, {cores} threads): "
            f"{steps} step(s) of a {t_frames}-frame clip at {H}x{W} ({dt / steps:.2f} s/step)")
    return fps, desc, cores, dt / steps * 1e3


def reference_arm(args, rank, world):
    if rank != 0:
        return
    T, H, W = WORKLOADS[args.workload]
    if args.workload == "seqformer_train":
        print(json.dumps({"impl": "reference", "unavailable": "no CPU leg for the optional seqformer_train workload; "
                          "use --workload idol_train (BASELINE configs[2]) or the default inference workload"}), flush=True)
        return
    if args.workload == "idol_train":
        threads = min(32, os.cpu_count() or 1)
        t0 = time.perf_counter()
        fps = cpu_train_sample(H, W, threads)
        desc = "1 (key, ref) pair 480x864, forward+backward once, oracle MSDA on host threads"
        print(json.dumps({"impl": "reference", "metric": TRAIN_METRIC, "value": round(fps, 4), "unit": "frames/s",
                          "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": round(2e3 / fps, 1),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": "IDOL-R50 training step", "where": "host CPU"},
                          "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": threads, "kind": "port",
                                           "sample": desc},
                          "e2e": {"value": round(fps, 4), "unit": "frames/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    fps, desc, cores, ms = cpu_oracle_fps(T, H, W, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC if args.workload == "720p" else METRIC.replace("720p", "480x864"),
        "value": round(fps, 4), "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SeqFormer-R50 inference, T={T}, {H}x{W}, 300 queries, random-init weights",
                   "where": "host CPU"},
        "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": round(fps, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- own arm
def own_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    from vnext_b200 import _cabi
    from vnext_b200.seqformer.meta_arch import SeqFormer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    T, H, W = WORKLOADS[args.workload]
    swin = None
    if args.workload.startswith("swinl"):           # BASELINE config 4's backbone (tf32 here; bf16 is a later round)
        from vnext_b200.swin_spec import SWIN_L
        swin = dict(SWIN_L)
    model = SeqFormer(device=str(dev), precision=args.precision)
    model.swin = swin
    model.load_reference_state_dict(make_weights(0, swin))
    model.eval()
    clip = make_clip_u8(T, H, W, rank)
    run = model.runner(T, H, W, (H, W))                    # captures the CUDA graph
    run.frames_dev.copy_(clip.to(dev))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput: W warm-ups, exactly K timed steps, CUDA events on the launch stream
    for _ in range(max(3, args.warmup)):
        run.run_device()
    run.stream.synchronize()
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(run.stream)
    for _ in range(args.steps):
        run.run_device()
    ev1.record(run.stream)
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.result()
    from vnext_b200.dist import max_over_ranks
    ms_total = max_over_ranks(ev0.elapsed_time(ev1), dev)
    fps = world * args.steps * T / (ms_total / 1e3)

    # ---- end to end through the reference-facing API: host frames in, host masks out, every step
    clip_pin = clip.pin_memory()                     # "host→device copy of that step's inputs from pinned host memory"
    batched = [{"image": [f for f in clip_pin], "height": H, "width": W, "file_names": [f"f{i}" for i in range(T)]}]
    for _ in range(max(3, args.warmup)):
        model(batched)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = model(batched)
    torch.cuda.synchronize()
    e2e_fps = world * args.steps * T / (max_over_ranks((time.perf_counter() - t0) * 1e3, dev) / 1e3)
    # same API, two clips in flight (model.stream): transfers + host list building overlap the next clip's compute
    for _ in model.stream([batched] * 4):
        pass
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for res in model.stream([batched] * args.steps):
        pass
    torch.cuda.synchronize()
    e2e_stream_fps = world * args.steps * T / (max_over_ranks((time.perf_counter() - t0) * 1e3, dev) / 1e3)

    # ---- the reference's own CPU-runnable geometry (BASELINE configs[1]: 480x864), same model, device-resident
    also = None
    if args.workload == "720p":
        T2, H2, W2 = WORKLOADS["480p"]
        run2 = model.runner(T2, H2, W2, (H2, W2))
        run2.frames_dev.copy_(make_clip_u8(T2, H2, W2, rank).to(dev))
        for _ in range(3):
            run2.run_device()
        barrier()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(run2.stream)
        for _ in range(args.steps):
            run2.run_device()
        a1.record(run2.stream)
        torch.cuda.synchronize()
        ms2 = max_over_ranks(a0.elapsed_time(a1), dev)
        also = {"workload": "SeqFormer-R50 T=5 480x864", "value": round(world * args.steps * T2 / (ms2 / 1e3), 2),
                "unit": "frames/s", "ms_per_step": round(ms2 / args.steps, 4)}
    if rank != 0:
        return
    # ---- per-kernel live timing on eager steps (same inputs, same stream discipline), launch count
    hbm, tflops, peak_src = load_peaks()
    eager = model.engine
    frames_f = run.frames_dev.float()
    n0 = _cabi.launch_count()
    eager.infer_video_device(frames_f, (H, W))
    launches_per_step = _cabi.launch_count() - n0
    with _cabi.KernelTimer() as kt:
        for _ in range(3):
            eager.infer_video_device(frames_f, (H, W))
    agg = kt.summary()
    # group call sites into kernel families (one __global__ each) and rank by time inside a step
    fam = {}
    for name, d in agg.items():
        f = ("gemm_tf32_kernel" if name.startswith("linear_tf32") else
             "msda_fwd_tile_kernel" if name.startswith("msda_") else name)
        g = fam.setdefault(f, {"calls": 0, "us": 0.0, "bytes": 0.0, "flops": 0.0})
        for k in g:
            g[k] += d[k]
    kernels = {f: {"calls_per_step": d["calls"] // 3, "us_per_step": round(d["us"] / 3, 1),
                   "GBps": round(d["bytes"] / d["us"] / 1e3, 1),
                   **({"TFLOPs": round(d["flops"] / d["us"] / 1e6, 1)} if d["flops"] else {})}
               for f, d in sorted(fam.items(), key=lambda kv: -kv[1]["us"])}
    roofline = None
    if fam:
        top = max(fam, key=lambda k: fam[k]["us"])
        d = fam[top]
        if top == "gemm_tf32_kernel":
            # one __global__, ~180 launches of very different shapes: split its call sites at the TF32 ridge point
            # (peak TF32 flop/s / HBM B/s; TF32 peak = half the measured bf16 rate) and quote each group against
            # the roof that bounds it; the headline entry is the group with more time in the step.
            tf32_peak = tflops / 2.0
            ridge = tf32_peak * 1e12 / (hbm * 1e9)
            grp = {"hbm": {"calls": 0, "us": 0.0, "bytes": 0.0, "flops": 0.0},
                   "tensor": {"calls": 0, "us": 0.0, "bytes": 0.0, "flops": 0.0}}
            for name, c in agg.items():
                if name.startswith("linear_tf32") or name.startswith("conv3x3_tf32"):
                    g = grp["tensor" if c["flops"] / max(c["bytes"], 1.0) > ridge else "hbm"]
                    for k in g:
                        g[k] += c[k]
            def entry(kind, g):
                if not g["us"]:
                    return None
                if kind == "hbm":
                    a, pk, unit = g["bytes"] / g["us"] / 1e3, hbm, "GB/s"
                else:
                    a, pk, unit = g["flops"] / g["us"] / 1e6, tf32_peak, "TFLOP/s"
                return {"bound": kind, "achieved": round(a, 1), "peak": pk, "unit": unit, "frac": round(a / pk, 4),
                        "launches_per_step": g["calls"] // 3, "us_per_step": round(g["us"] / 3, 1)}
            main_kind = max(grp, key=lambda k: grp[k]["us"])
            roofline = {"kernel": top, **entry(main_kind, grp[main_kind]), "traffic": None, "peak_source": peak_src,
                        "ridge_flop_per_byte": round(ridge, 1),
                        "other_group": entry("tensor" if main_kind == "hbm" else "hbm",
                                             grp["tensor" if main_kind == "hbm" else "hbm"]),
                        "note": "algorithmic bytes (or flops) summed over the group's launches in a step / their summed "
                                "CUDA-event time; TF32 peak = measured bf16 peak / 2"}
        else:
            ach = d["bytes"] / d["us"] / 1e3
            roofline = {"kernel": top, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s",
                        "frac": round(ach / hbm, 4), "traffic": None, "peak_source": peak_src,
                        "launches_per_step": d["calls"] // 3, "avg_us": round(d["us"] / d["calls"], 2),
                        "algorithmic_bytes_per_launch": round(d["bytes"] / d["calls"]),
                        "note": "sum of algorithmic bytes over the family's launches in a step / their summed CUDA-event time"}
        if "msda_fwd_tile_kernel" in fam:
            m = fam["msda_fwd_tile_kernel"]
            roofline["msda"] = {"achieved": round(m["bytes"] / m["us"] / 1e3, 1), "frac": round(m["bytes"] / m["us"] / 1e3 / hbm, 4),
                                "us_per_step": round(m["us"] / 3, 1)}
            if args.workload == "720p":        # dram__bytes_read+write of one encoder launch (350.5 MB algorithmic), ncu --set full
                roofline["msda"].update(traffic=341668096, traffic_source="profiles/r01_msda_enc_v3_ncu_full.txt")

    cpu = None
    if args.gpus == 1 and not args.no_cpu_baseline and swin is None:
        cfps, desc, cores, _ = cpu_oracle_fps(T, H, W, 1, 0, budget_s=20.0)
        cpu = {"value": round(cfps, 4), "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc}

    line = {
        "metric": {"720p": METRIC, "480p": METRIC.replace("720p", "480x864"),
                   "swinl_720p": METRIC.replace("R50", "SwinL")}[args.workload],
        "value": round(fps, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else args.precision, "data": "synthetic",
        "config": {"workload": f"SeqFormer-R50 inference, T={T}, {H}x{W} (padded to /32), 300 queries, 40 classes, "
                               "random-init weights, one clip per step per GPU",
                   "l2": "no flush: per-step working set (fp32 weights ~165 MB + activations > 1 GB) exceeds the 126 MB L2",
                   "parallelism": f"{world} independent replicas, no data-path collective",
                   "precision": args.precision},
        "clocks": clocks,
        "e2e": {"value": round(e2e_fps, 2), "unit": "frames/s", "h2d_bytes_per_step": run.h2d_bytes,
                "d2h_bytes_per_step": run.d2h_bytes, "instances_returned": len(res["pred_scores"]),
                "api": "SeqFormer.forward(batched_inputs), one synchronous call per clip",
                "pipelined_value": round(e2e_stream_fps, 2),
                "pipelined_api": "SeqFormer.stream(iterable): same inputs/outputs, 2 clips in flight"},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "also": also,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- IDOL training step
TRAIN_METRIC = "frames/sec IDOL-R50 two-frame training step (fwd+bwd+grad all-reduce+AdamW) 480x864 synthetic"


def cpu_train_sample(H, W, threads):
    """CPU leg of the training workload: ONE (key, ref) pair, forward + backward of the same model code with the
    oracle's MSDA (the reference's ms_deform_attn_core_pytorch restated) substituted for the CUDA op."""
    from oracle.msda_oracle import forward_torch as msda_core
    from vnext_b200.idol import train as T
    from vnext_b200.idol.state_spec import idol_r50_state_spec
    from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_targets

    class OracleMSDA:
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, im2col_step):
            return msda_core(value, shapes, loc, attn)
    real = T.MSDeformAttnFunction
    T.MSDeformAttnFunction = OracleMSDA
    try:
        torch.set_num_threads(threads)
        model = T.IDOLTrainModel(fill_state_dict(idol_r50_state_spec(), 0), device="cpu", dropout=0.1).train()
        frames = torch.stack(synthetic_clip(2, H, W, 0))
        det, ref = synthetic_targets(1, H, W, 0)
        t0 = time.perf_counter()
        loss, _ = model(frames, det, ref)
        sum(model.weighted(loss).values()).backward()
        return 2 / (time.perf_counter() - t0)
    finally:
        T.MSDeformAttnFunction = real


def train_arm(args, rank, world, local_rank):
    import random

    import torch.distributed as dist
    from vnext_b200 import _cabi
    from vnext_b200.dist import max_over_ranks
    from vnext_b200.idol import train as T
    from vnext_b200.idol.state_spec import idol_r50_state_spec
    from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_targets

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    seq = args.workload == "seqformer_train"
    NF, H, W = WORKLOADS[args.workload]
    host = torch.stack(synthetic_clip(NF, H, W, rank)).to(torch.uint8).pin_memory()
    if seq:
        from vnext_b200.seqformer import train as ST
        from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
        from vnext_b200.synth import synthetic_clip_targets
        model = ST.SeqFormerTrainModel(fill_state_dict(seqformer_r50_state_spec(), 0), 5, device=dev, dropout=0.1,
                                       precision=args.precision).train()
        trainer = ST.SeqFormerTrainer(model)
        tg = [{k: v.to(dev) for k, v in t.items()} for t in synthetic_clip_targets(NF // 5, 5, H, W, rank)]
        step_args = (tg,)
    else:
        model = T.IDOLTrainModel(fill_state_dict(idol_r50_state_spec(), 0), device=dev, dropout=0.1,
                                 precision=args.precision).train()
        trainer = T.IDOLTrainer(model)
        det, ref = synthetic_targets(NF // 2, H, W, rank)
        step_args = ([{k: v.to(dev) for k, v in t.items()} for t in det],
                     [{k: v.to(dev) for k, v in t.items()} for t in ref])
    frames = host.to(dev).float()
    random.seed(rank)

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(3, args.warmup)):
        trainer.step(frames, *step_args)
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    n0 = _cabi.lib().vnext_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        trainer.step(frames, *step_args)
    ev1.record()
    torch.cuda.synchronize()
    launches = _cabi.lib().vnext_launch_count() - n0
    barrier()
    clocks = sampler.result()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1), dev)
    fps = world * args.steps * NF / (ms_total / 1e3)

    # end to end: uint8 frames from pinned host memory every step, the loss read back every step
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, total, _ = trainer.step(host.to(dev, non_blocking=True).float(), *step_args)
        total.item()
    torch.cuda.synchronize()
    e2e = world * args.steps * NF / (max_over_ranks((time.perf_counter() - t0) * 1e3, dev) / 1e3)

    with _cabi.KernelTimer() as kt:
        trainer.step(frames, *step_args)
    agg = kt.summary()
    hbm_peak, _, peak_src = load_peaks()
    kern = {k: {"calls": v["calls"], "us_per_step": round(v["us"], 1),
                "GBps": round(v["bytes"] / (v["us"] * 1e-6) / 1e9, 1) if v["us"] else None} for k, v in agg.items()}
    dom = agg.get("msda_bwd_enc")
    roof = None
    if dom:
        ach = dom["bytes"] / (dom["us"] * 1e-6) / 1e9
        roof = {"bound": "hbm", "kernel": f"msda_bwd_tile_kernel (encoder shape, {NF} frames)", "achieved": round(ach, 1),
                "peak": hbm_peak, "unit": "GB/s", "frac": round(ach / hbm_peak, 4), "traffic": None,
                "peak_source": peak_src, "own_kernels": kern}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not seq:
        threads = min(32, os.cpu_count() or 1)
        cpu = {"value": round(cpu_train_sample(H, W, threads), 4), "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "1 (key, ref) pair 480x864, forward+backward once (no optimiser), oracle MSDA on host threads"}
    if rank == 0:
        print(json.dumps({
            "metric": TRAIN_METRIC.replace("IDOL-R50 two-frame", "SeqFormer-R50 five-frame clip") if seq else TRAIN_METRIC,
            "value": round(fps, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 GEMMs/convs, f32 MSDA + accumulate" if args.precision == "tf32" else "f32 (cuDNN default TF32 convs)",
            "data": "synthetic",
            "config": {"workload": (f"SeqFormer-R50 training step, {NF // 5} clips x 5 frames {H}x{W} per GPU, 300 queries, "
                                    "dropout 0.1, Hungarian + 5 loss terms x 6 levels, AdamW, clip 0.01" if seq else
                                    f"IDOL-R50 training step, {NF // 2} videos x (key, ref) frames {H}x{W} per GPU, 300 "
                                    "queries, dropout 0.1, simOTA + 7 loss terms x 6 levels, AdamW lr 1e-4 (backbone x0.1), "
                                    "clip 0.01") + "; activations >> L2 so no flush", "parallelism": f"dp{world}"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": round(e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": int(host.numel()),
                    "d2h_bytes_per_step": 4},
            "roofline": roof, "cpu_baseline": cpu}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="720p", choices=list(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the CPU oracle baseline)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        (train_arm if args.workload.endswith("_train") else own_arm)(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
