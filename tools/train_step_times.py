import sys, time, random, torch
sys.path.insert(0, "/root/repo")
import os
os.chdir("/root/repo")
from vnext_b200.hostcfg import tune_host_threads
if os.environ.get("TUNE", "1") == "1":
    print("threads", tune_host_threads())
from vnext_b200 import msda
msda.WINDOW_MODE = os.environ.get("MSDA_MODE", "cached")
from vnext_b200.idol import train as T
from vnext_b200.idol.state_spec import idol_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_targets
dev = torch.device("cuda", 0)
NF, H, W = 8, 480, 864
model = T.IDOLTrainModel(fill_state_dict(idol_r50_state_spec(), 0), device=dev, dropout=0.1, precision="tf32").train()
trainer = T.IDOLTrainer(model)
det, ref = synthetic_targets(NF // 2, H, W, 0)
args = ([{k: v.to(dev) for k, v in t.items()} for t in det], [{k: v.to(dev) for k, v in t.items()} for t in ref])
frames = torch.stack(synthetic_clip(NF, H, W, 0)).to(dev).float()
random.seed(0)
ts = []
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    trainer.step(frames, *args)
    torch.cuda.synchronize(); ts.append(round((time.perf_counter() - t0) * 1e3, 1))
print(os.environ.get("MSDA_MODE", "cached"), "step ms:", ts)
