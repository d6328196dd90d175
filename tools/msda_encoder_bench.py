"""Encoder-shape MSDA forward (queries = the pixels of the pyramid): window kernel vs tile kernel vs the reference's
stock CUDA op, on encoder-like sampling locations (pixel centre + N(0, sigma) pixel offsets) and on uniform random ones.

    python tools/msda_encoder_bench.py [--out gpurun_out/msda_encoder.jsonl] [--iters 50]

CUDA events on the current stream, 10 warm-ups, median of `iters`; L2 flushed between launches (256 MiB memset).
Algorithmic bytes = SURVEY §8d fwd_bytes with e = element size of value/out and 4-byte loc/attn."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from msda_sweep import load_stock, time_fn  # noqa: E402
from test_msda_window_gpu import encoder_inputs  # noqa: E402
from vnext_b200 import _cabi, msda  # noqa: E402

PYR = {"480x864": [(60, 108), (30, 54), (15, 27), (8, 14)], "736x1280": [(92, 160), (46, 80), (23, 40), (12, 20)]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "msda_encoder.jsonl"))
    ap.add_argument("--cases", default="all")
    args = ap.parse_args()
    stock = load_stock()
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cases = [("736x1280", 5, 4, 1.6), ("480x864", 5, 4, 1.6), ("736x1280", 1, 4, 1.6), ("736x1280", 8, 4, 1.6),
             ("736x1280", 5, 4, 3.0), ("736x1280", 5, 4, None), ("736x1280", 5, 8, 1.6), ("480x864", 8, 4, 1.6)]
    if args.cases == "quick":
        cases = cases[:2]
    M, D = 8, 32
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "a") as fout:
        for pyr, N, K, sigma in cases:
            shapes = PYR[pyr]
            S = sum(h * w for h, w in shapes)
            if sigma is None:            # no locality at all: uniform locations
                from vnext_b200.synth import make_msda_inputs
                value, sh, lsi, loc, attn = make_msda_inputs(N, S, M, D, shapes, K, 1)
            else:
                value, sh, lsi, loc, attn = encoder_inputs(N, M, shapes, K, 1, sigma, 0.0)
            dv, dsh, dlsi, dloc, dattn = [t.cuda() for t in (value, sh, lsi, loc, attn)]
            rec = {"pyr": pyr, "N": N, "Lq": S, "L": 4, "K": K, "sigma_px": sigma if sigma is not None else "uniform"}
            for dtype, tag, e in ((torch.float32, "f32", 4), (torch.bfloat16, "bf16", 2)):
                v = dv.to(dtype)
                alg = N * (e * S * M * D + 4 * 3 * S * M * 4 * K + e * S * M * D)
                for mode in ("cached", "auto", "off"):
                    msda.WINDOW_MODE = mode
                    med, best = time_fn(lambda: msda.ms_deform_attn_forward(v, dsh, dlsi, dloc, dattn, 64), args.iters, flush)
                    name = {"auto": "window", "off": "tile", "cached": "cached"}[mode]
                    assert _cabi.last_variant().startswith("fwd_window") == (mode == "auto"), _cabi.last_variant()
                    assert _cabi.last_variant().startswith("fwd_cached") == (mode == "cached"), _cabi.last_variant()
                    rec[f"{name}_{tag}_us"] = round(med, 1)
                    rec[f"{name}_{tag}_GBps"] = round(alg / med / 1e3, 1)
                    rec[f"{name}_{tag}_frac"] = round(alg / med / 1e3 / hbm, 4)
                msda.WINDOW_MODE = "off"
                rec[f"alg_MB_{tag}"] = round(alg / 1e6, 1)
            if stock is not None:
                med, _ = time_fn(lambda: stock.ms_deform_attn_forward(dv, dsh, dlsi, dloc, dattn, 64), args.iters, flush)
                rec["stock_f32_us"] = round(med, 1)
                rec["cached_f32_vs_stock"] = round(med / rec["cached_f32_us"], 2)
                rec["cached_bf16_vs_stock_f32"] = round(med / rec["cached_bf16_us"], 2)
                rec["window_f32_vs_stock"] = round(med / rec["window_f32_us"], 2)
                rec["window_bf16_vs_stock_f32"] = round(med / rec["window_bf16_us"], 2)
                rec["tile_f32_vs_stock"] = round(med / rec["tile_f32_us"], 2)
            print(json.dumps(rec), flush=True)
            fout.write(json.dumps(rec) + "\n")
            fout.flush()


if __name__ == "__main__":
    main()
