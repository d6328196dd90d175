"""Where the end-to-end (host frames in -> host masks out) milliseconds go: per-stage host timers around
ClipRunner.submit()/collect() and SeqFormerEngine.finish_video(), 720p T=5.  Diagnostic, not a bench."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vnext_b200.seqformer.engine import SeqFormerEngine  # noqa: E402
from vnext_b200.seqformer.meta_arch import SeqFormer  # noqa: E402


def main():
    T, H, W = 5, 720, 1280
    precision = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    if os.environ.get("VNEXT_TUNE_THREADS", "1") == "1":
        from vnext_b200.hostcfg import tune_host_threads
        tune_host_threads()
    model = SeqFormer(device="cuda", precision=precision)
    model.load_reference_state_dict(bench.make_weights(0))
    model.eval()
    clip = bench.make_clip_u8(T, H, W, 0).pin_memory()
    frames = [f for f in clip]
    run = model.runner(T, H, W, (H, W))
    batched = [{"image": frames, "height": H, "width": W, "file_names": [str(i) for i in range(T)]}]
    for _ in range(5):
        model(batched)
    acc = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    n = 20
    for _ in range(n):
        t0 = time.perf_counter()
        for i, f in enumerate(frames):
            run.frames_pin[i].copy_(f)
        tick("stage_host_copy", t0)
        t0 = time.perf_counter()
        sp = torch.empty(run.scores.shape, dtype=run.scores.dtype, pin_memory=True)
        mp = torch.empty(run.masks.shape, dtype=run.masks.dtype, pin_memory=True)
        tick("pinned_alloc", t0)
        t0 = time.perf_counter()
        with torch.cuda.stream(run.stream):
            run.frames_dev.copy_(run.frames_pin, non_blocking=True)
        tick("h2d", t0)
        t0 = time.perf_counter()
        run.run_device()
        tick("graph", t0)
        t0 = time.perf_counter()
        with torch.cuda.stream(run.stream):
            sp.copy_(run.scores, non_blocking=True)
            mp.copy_(run.masks, non_blocking=True)
        tick("d2h", t0)
        t0 = time.perf_counter()
        res = SeqFormerEngine.finish_video(sp, mp, (H, W))
        tick("finish_video", t0)
        del sp, mp, res
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        model(batched)
    torch.cuda.synchronize()
    whole = (time.perf_counter() - t0) * 1e3 / n
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        res = model(batched)                      # as bench.py does: the previous clip's result stays alive during the call
    torch.cuda.synchronize()
    whole_retain = (time.perf_counter() - t0) * 1e3 / n
    import gc
    events = []
    gc.callbacks.append(lambda phase, info: events.append((phase, info.get("generation"), time.perf_counter())))
    per_call = []
    for _ in range(24):
        t1 = time.perf_counter()
        res = model(batched)
        per_call.append(round((time.perf_counter() - t1) * 1e3, 2))
    gc_starts = [(g, t) for ph, g, t in events if ph == "start"]
    gc_stops = [(g, t) for ph, g, t in events if ph == "stop"]
    gc_ms = [(g0, round((t1_ - t0_) * 1e3, 2)) for (g0, t0_), (_, t1_) in zip(gc_starts, gc_stops)]
    gc.callbacks.clear()
    gc.collect()
    gc.freeze()
    frozen = []
    for _ in range(24):
        t1 = time.perf_counter()
        res = model(batched)
        frozen.append(round((time.perf_counter() - t1) * 1e3, 2))
    print(json.dumps({"gc_collections_during_24_calls (generation, ms)": gc_ms, "per_call_ms_after_gc_freeze": frozen}))
    print(json.dumps({"precision": precision, "forward_call_retaining_previous_result_ms": round(whole_retain, 3),
                      "per_call_ms": per_call, "per_stage_ms": {k: round(v / n, 3) for k, v in acc.items()},
                      "sum_ms": round(sum(acc.values()) / n, 3), "forward_call_ms": round(whole, 3),
                      "cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()}))


if __name__ == "__main__":
    main()
