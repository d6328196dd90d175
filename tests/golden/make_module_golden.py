"""Golden outputs of the REFERENCE's own MSDeformAttn nn.Modules (IDOL and SeqFormer variants), CPU path.
Run only in the build container: python tests/golden/make_module_golden.py"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from vnext_b200.synth import fill_state_dict  # noqa: E402

SHAPES = [(12, 16), (6, 8), (3, 4), (2, 2)]


def mount(name, path):
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    pkg = types.ModuleType(name)
    pkg.__path__ = [path]
    sys.modules[name] = pkg
    func = importlib.import_module(name + ".functions.ms_deform_attn_func")
    mod = importlib.import_module(name + ".modules.ms_deform_attn")

    class CpuMSDA:
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, im2col_step):
            return func.ms_deform_attn_core_pytorch(value, shapes, loc, attn)
    mod.MSDeformAttnFunction = CpuMSDA
    return mod


def inputs(seed, N, nf, Lq):
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(SHAPES, dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    return dict(shapes=shapes, lsi=lsi, S=S,
                src=torch.randn(N, nf, S, 256, generator=g), query=torch.randn(N, nf, Lq, 256, generator=g),
                ref2=torch.rand(N, nf, Lq, 4, 2, generator=g), ref4=torch.rand(N, nf, Lq, 4, 4, generator=g) * 0.5 + 0.2,
                mask=torch.rand(N, nf, S, generator=g) < 0.1)


def main():
    out = {}
    idol = mount("refidol_ops", "/root/reference/projects/IDOL/idol/models/ops")
    seq = mount("refseq_ops", "/root/reference/projects/SeqFormer/seqformer/models/ops")
    with torch.no_grad():
        m = idol.MSDeformAttn(256, 4, 8, 4).eval()
        m.load_state_dict(fill_state_dict({k: v.shape for k, v in m.state_dict().items()}, 11))
        d = inputs(1, 2, 1, 50)
        for tag, ref in (("r2", d["ref2"][:, 0]), ("r4", d["ref4"][:, 0])):
            o, loc, aw = m(d["query"][:, 0], ref, d["src"][:, 0], d["shapes"], d["lsi"], d["mask"][:, 0])
            out[f"idol_{tag}_out"], out[f"idol_{tag}_loc"], out[f"idol_{tag}_attn"] = o.numpy(), loc.numpy(), aw.numpy()
        me = seq.MSDeformAttn(256, 4, 8, 4, "encode").eval()
        me.load_state_dict(fill_state_dict({k: v.shape for k, v in me.state_dict().items()}, 12))
        d = inputs(2, 1, 3, 40)
        qe = torch.randn(1, 3, d["S"], 256, generator=torch.Generator().manual_seed(5))
        refe = torch.rand(1, d["S"], 4, 2, generator=torch.Generator().manual_seed(6))
        out["seq_enc_out"] = me(qe, None, refe, d["src"], d["shapes"], d["lsi"], d["mask"]).numpy()
        md = seq.MSDeformAttn(256, 4, 8, 4, "decode").eval()
        md.load_state_dict(fill_state_dict({k: v.shape for k, v in md.state_dict().items()}, 13))
        # first layer (query_box [N, Lq, C], 2-d refs) and later layer (query_box [N, nf, Lq, C], 4-d refs)
        o, ob, loc, aw = md(d["query"][:, 0], d["query"][:, 0], d["ref2"], d["src"], d["shapes"], d["lsi"], d["mask"])
        out["seq_dec0_out"], out["seq_dec0_box"], out["seq_dec0_loc"], out["seq_dec0_attn"] = (o.numpy(), ob.numpy(),
                                                                                                 loc.numpy(), aw.numpy())
        o, ob, loc, aw = md(d["query"][:, 0], d["query"], d["ref4"], d["src"], d["shapes"], d["lsi"], d["mask"])
        out["seq_dec1_out"], out["seq_dec1_box"], out["seq_dec1_loc"], out["seq_dec1_attn"] = (o.numpy(), ob.numpy(),
                                                                                                 loc.numpy(), aw.numpy())
    path = os.path.join(HERE, "module_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
