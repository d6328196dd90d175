"""MSDA microbenchmark sweep (BASELINE config 5): this repo's kernels vs the reference's stock CUDA op
(oracle/_ref/msda_stock_ref.so, comparator only) on identical inputs.

    python tools/msda_sweep.py [--quick] [--out gpurun_out/msda_sweep.jsonl]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/msda_sweep.py --quick
        (N > 1 ranks: the batch N is split evenly over the ranks, no communication on the data path; every time is the
         max over ranks, configs with N < world are skipped — SURVEY §8d cfg 5 "1 GPU and 8 GPUs")

Timing: CUDA events on the current stream, 20 warm-ups + `iters` timed launches, L2 flushed between
launches with a 256 MiB memset unless --no-flush (then the value tensor is L2-resident when it fits).
"""
import argparse
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vnext_b200.synth import make_msda_inputs as make_inputs  # noqa: E402
from vnext_b200 import msda  # noqa: E402

PYR = {"480x864": [(60, 108), (30, 54), (15, 27), (8, 14)], "736x1280": [(92, 160), (46, 80), (23, 40), (12, 20)]}


def load_stock():
    path = os.path.join(ROOT, "oracle", "_ref", "msda_stock_ref.so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("msda_stock_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def fwd_bytes(e, N, S, M, D, Lq, L, K):
    return e * N * (S * M * D + 3 * Lq * M * L * K + Lq * M * D)


def bwd_bytes(e, N, S, M, D, Lq, L, K):
    return e * N * (2 * S * M * D + Lq * M * D + 6 * Lq * M * L * K)


def time_fn(fn, iters, flush):
    for _ in range(20 if iters >= 20 else 3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "msda_sweep.jsonl"))
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")

    def over_ranks(us):
        if world == 1:
            return us
        t = torch.tensor([us], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    stock = load_stock()
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cases = []
    if args.quick:
        cases = [("480x864", 5, None, 4, 4), ("480x864", 5, 300, 4, 4), ("480x864", 1, None, 4, 4),
                 ("480x864", 64, 300, 4, 4), ("736x1280", 5, None, 4, 4), ("480x864", 8, None, 4, 4),
                 ("736x1280", 8, None, 4, 4), ("480x864", 64, 10000, 4, 4), ("736x1280", 64, 10000, 4, 8)]
    else:
        for pyr in PYR:
            for N in (1, 2, 4, 8, 16, 32, 64):
                for Lq in (300, 1000, 10000, None):
                    for L, K in ((4, 4), (4, 8), (1, 4), (1, 8)):
                        if Lq is None and (N > 8 or (L, K) != (4, 4)):
                            continue
                        cases.append((pyr, N, Lq, L, K))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fout = open(args.out, "a") if rank == 0 else open(os.devnull, "w")
    M, D = 8, 32
    for pyr, N_total, Lq, L, K in cases:
        if N_total % world:
            continue
        N = N_total // world                       # this rank's share of the batch
        shapes = PYR[pyr][:L]
        S = sum(h * w for h, w in shapes)
        lq = S if Lq is None else Lq
        value, sh, lsi, loc, attn = make_inputs(N, lq, M, D, shapes, K, hash((pyr, N, lq, L, K)) % 10000)
        dv, dsh, dlsi, dloc, dattn = [t.cuda() for t in (value, sh, lsi, loc, attn)]
        gout = torch.rand(N, lq, M * D, device="cuda")
        rec = {"pyr": pyr, "N": N_total, "n_gpus": world, "Lq": lq, "L": L, "K": K, "S": S,
               "fwd_MB": fwd_bytes(4, N_total, S, M, D, lq, L, K) / 1e6,
               "bwd_MB": bwd_bytes(4, N_total, S, M, D, lq, L, K) / 1e6}
        for dtype, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            v = dv.to(dtype)
            med, best = time_fn(lambda: msda.ms_deform_attn_forward(v, dsh, dlsi, dloc, dattn, 64), args.iters, flush)
            med = over_ranks(med)
            e_bytes = rec["fwd_MB"] * 1e6 if dtype == torch.float32 else (
                N_total * (2 * S * M * D + 12 * lq * M * L * K + 2 * lq * M * D))
            rec[f"fwd_{tag}_us"] = round(med, 2)
            rec[f"fwd_{tag}_GBps"] = round(e_bytes / med / 1e3, 1)
            rec[f"fwd_{tag}_frac"] = round(e_bytes / med / 1e3 / (hbm * world), 4)
        med, best = time_fn(lambda: msda.ms_deform_attn_backward(dv, dsh, dlsi, dloc, dattn, gout, 64), args.iters, flush)
        med = over_ranks(med)
        rec["bwd_f32_us"] = round(med, 2)
        rec["bwd_f32_GBps"] = round(rec["bwd_MB"] * 1e3 / med, 1)
        if stock is not None:
            step = 64
            if N % min(N, step) == 0:
                med, _ = time_fn(lambda: stock.ms_deform_attn_forward(dv, dsh, dlsi, dloc, dattn, step), args.iters, flush)
                med = over_ranks(med)
                rec["stock_fwd_us"] = round(med, 2)
                rec["speedup_fwd"] = round(med / rec["fwd_f32_us"], 2)
                med, _ = time_fn(lambda: stock.ms_deform_attn_backward(dv, dsh, dlsi, dloc, dattn, gout, step), args.iters, flush)
                med = over_ranks(med)
                rec["stock_bwd_us"] = round(med, 2)
                rec["speedup_bwd"] = round(med / rec["bwd_f32_us"], 2)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        fout.write(json.dumps(rec) + "\n")
        fout.flush()


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", 1)) > 1:
        torch.distributed.destroy_process_group()
