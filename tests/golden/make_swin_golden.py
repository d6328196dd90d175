"""Golden outputs of the REFERENCE's own SwinTransformer (projects/IDOL/idol/backbone/swin.py), CPU.
Stubs: timm.models.layers (DropPath -> identity in eval, to_2tuple, trunc_normal_) and detectron2.modeling names
imported at module level.  Run only in the build container."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from vnext_b200.swin_spec import swin_state_spec  # noqa: E402
from vnext_b200.synth import fill_state_dict  # noqa: E402

CASES = {  # name: (cfg, B, H, W, seed)
    "w12": (dict(embed_dim=96, depths=(2, 2, 2, 2), num_heads=(3, 6, 12, 24), window_size=12), 2, 160, 224, 0),
    "w7": (dict(embed_dim=64, depths=(2, 2, 2, 2), num_heads=(2, 4, 8, 16), window_size=7), 1, 96, 130, 1),
}


def mount():
    timm = types.ModuleType("timm"); tm = types.ModuleType("timm.models"); tl = types.ModuleType("timm.models.layers")

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x
    tl.DropPath, tl.to_2tuple, tl.trunc_normal_ = DropPath, (lambda x: (x, x)), (lambda t, std=0.02: t)
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    d2 = types.ModuleType("detectron2"); dm = types.ModuleType("detectron2.modeling")

    class _Reg:
        def register(self):
            return lambda c: c
    dm.BACKBONE_REGISTRY, dm.Backbone, dm.ShapeSpec = _Reg(), nn.Module, dict
    sys.modules.update({"detectron2": d2, "detectron2.modeling": dm})
    spec = importlib.util.spec_from_file_location("ref_swin", "/root/reference/projects/IDOL/idol/backbone/swin.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    mod = mount()
    out = {}
    for name, (cfg, B, H, W, seed) in CASES.items():
        m = mod.SwinTransformer(pretrain_img_size=224, patch_size=4, in_chans=3, embed_dim=cfg["embed_dim"],
                                depths=list(cfg["depths"]), num_heads=list(cfg["num_heads"]),
                                window_size=cfg["window_size"], drop_path_rate=0.3)
        m.eval()                      # (the reference's train() override returns None, so no chaining)
        spec = swin_state_spec("", **cfg)
        assert sorted(spec) == sorted(m.state_dict()), set(spec) ^ set(m.state_dict())
        sd = fill_state_dict(spec, seed)
        for k, v in m.state_dict().items():
            if not v.dtype.is_floating_point:
                assert torch.equal(v, sd[k]), k
        m.load_state_dict(sd)
        x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(seed + 50))
        with torch.no_grad():
            o = m(x)
        for k, v in o.items():
            out[f"{name}_{k}"] = v.numpy()[:, ::4]
            print(name, k, tuple(v.shape), float(v.abs().mean()))
    path = os.path.join(HERE, "swin_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
