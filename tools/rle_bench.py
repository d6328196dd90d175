"""Device RLE encoder (csrc/rle_encode.cu) at the BASELINE output size — 50 masks (top-10 instances x 5 frames) of 720x1280 —
timed with CUDA events, with the CPU oracle (pycocotools algorithm, numpy) timed beside it on a sample of the same masks.
  python tools/rle_bench.py > profiles/r02_rle_bench.json
Roofline: HBM; algorithmic bytes = one read of every mask byte + the string bytes written."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rle_oracle as R  # noqa: E402  (tool: the oracle is the CPU leg)
from vnext_b200.rle import RleEncoder  # noqa: E402


def blobs(n, h, w, seed, smooth):
    g = torch.Generator().manual_seed(seed)
    f = torch.randn((n, 1, h // smooth + 2, w // smooth + 2), generator=g)
    return torch.nn.functional.interpolate(f, size=(h, w), mode="bilinear", align_corners=False)[:, 0] > 0.3


def main():
    n, h, w = 50, 720, 1280
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6575.4))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = []
    for name, smooth in (("blobs_40px", 40), ("blobs_9px", 9), ("noise_2px", 2)):
        masks = blobs(n, h, w, 1, smooth)
        dm = masks.cuda()
        enc = RleEncoder(n, h, w, "cuda", capacity=n * h * w)
        for _ in range(3):
            enc.encode(dm)
        torch.cuda.synchronize()
        times = []
        for cold in (True, False):
            ts = []
            for _ in range(10):
                if cold:
                    flush.zero_()                         # masks out of the L2 (in the engine they were just written: L2 hits)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                enc.encode(dm)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            times.append(sorted(ts)[len(ts) // 2])
        total = int(enc.char_offsets[-1])
        offs = enc.char_offsets.cpu().tolist()
        runs = 10 * sum(len(R.from_string(bytes(enc.chars[offs[i]:offs[i + 1]].cpu().numpy()))) for i in range(0, n, 10))
        alg = n * h * w + total
        t0 = time.perf_counter()
        sample = 5
        for i in range(sample):
            R.encode(masks[i].numpy())
        cpu_ms_per_mask = (time.perf_counter() - t0) * 1e3 / sample
        out.append({"masks": f"{n} x {h} x {w}, {name}", "string_bytes": total, "approx_runs": runs,
                    "us_l2_cold": round(times[0], 1), "us_l2_warm": round(times[1], 1),
                    "algorithmic_bytes": alg, "achieved_gbps_cold": round(alg / times[0] / 1e3, 1),
                    "frac_of_hbm_peak_cold": round(alg / times[0] / 1e3 / hbm, 3), "hbm_peak_gbps": hbm,
                    "cpu_oracle_ms_per_mask": round(cpu_ms_per_mask, 2), "cpu_oracle_ms_for_all": round(cpu_ms_per_mask * n, 1),
                    "d2h_bytes_instead_of": n * h * w})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
