"""Where one IDOL training step spends its time: torch.profiler kernel totals vs wall (run on the GPU box)."""
import os
import random
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vnext_b200.idol import train as T  # noqa: E402
from vnext_b200.idol.state_spec import idol_r50_state_spec  # noqa: E402
from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_targets  # noqa: E402

if os.environ.get("CUDNN_BENCH"):
    torch.backends.cudnn.benchmark = True
NF, H, W = 8, 480, 864
PREC = os.environ.get("PREC", "fp32")
dev = torch.device("cuda")
model = T.IDOLTrainModel(fill_state_dict(idol_r50_state_spec(), 0), device=dev, dropout=0.1, precision=PREC, channels_last=os.environ.get("CL", "1") == "1").train()
trainer = T.IDOLTrainer(model)
frames = torch.stack(synthetic_clip(NF, H, W, 0)).to(dev)
det, ref = synthetic_targets(NF // 2, H, W, 0)
det = [{k: v.to(dev) for k, v in t.items()} for t in det]
ref = [{k: v.to(dev) for k, v in t.items()} for t in ref]
random.seed(0)
for _ in range(3):
    trainer.step(frames, det, ref)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    trainer.step(frames, det, ref)
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 3 * 1e3)

# phase timing with syncs
def phase():
    torch.cuda.synchronize(); return time.perf_counter()
t = phase(); model.flat.g.zero_(); d = model.dense(frames); t1 = phase()
print("dense forward ms", (t1 - t) * 1e3)
loss, _ = model(frames, det, ref); t2 = phase()
print("full forward (dense+heads+matching+losses) ms", (t2 - t1) * 1e3)
tot = sum(model.weighted(loss).values()); tot.backward(); t3 = phase()
print("backward ms", (t3 - t2) * 1e3)
trainer.apply_gradients(); t4 = phase()
print("optimizer ms", (t4 - t3) * 1e3)

from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    trainer.step(frames, det, ref)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
