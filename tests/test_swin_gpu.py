"""GPU parity of the Swin kernels (fused window attention, row LayerNorm, GELU epilogue) and of the Swin-backed engines."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import idol_oracle as IO, seqformer_oracle as SO, swin_oracle as WO
from vnext_b200.swin_spec import swin_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = dict(embed_dim=96, depths=(2, 2, 2, 2), num_heads=(3, 6, 12, 24), window_size=12)


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-6)).item()


@pytest.mark.parametrize("B,H,W,heads,ws,shift", [(2, 40, 56, 3, 12, 0), (2, 40, 56, 3, 12, 6), (1, 24, 33, 2, 7, 3),
                                                   (5, 23, 40, 48, 12, 6), (1, 12, 12, 6, 12, 0), (3, 5, 7, 4, 4, 2)])
def test_window_attention_kernel(B, H, W, heads, ws, shift):
    from vnext_b200 import fused
    from vnext_b200.swin.backbone import _window_attention_torch
    g = torch.Generator().manual_seed(B * 100 + H + shift)
    C = heads * 32
    qkv = torch.randn(B, H, W, 3 * C, generator=g).cuda()
    bias = torch.randn(3 * C, generator=g).cuda()
    table = (torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5).cuda()
    want = _window_attention_torch(qkv.double(), bias.double(), table.double(), heads, ws, shift)
    got = fused.swin_window_attention(qkv, bias, table, heads, ws, shift)
    assert (got.double() - want).abs().max().item() < 2e-5
    if ws >= 4:           # tensor-core (TF32) variant: same contract, TF32 tolerance
        tc = fused.swin_window_attention(qkv, bias, table, heads, ws, shift, tensor_cores=True)
        assert (tc.double() - want).abs().max().item() < 5e-3 * max(1.0, want.abs().max().item())
        # bf16 storage of qkv / out (the Swin bf16 mode): same kernel, operands widened in the gather
        q16 = qkv.bfloat16()
        want16 = _window_attention_torch(q16.double(), bias.double(), table.double(), heads, ws, shift)
        tc16 = fused.swin_window_attention(q16, bias, table, heads, ws, shift, tensor_cores=True)
        assert tc16.dtype == torch.bfloat16
        assert (tc16.double() - want16).abs().max().item() < 1.2e-2 * max(1.0, want16.abs().max().item())


def test_layernorm_and_gelu_epilogue():
    from vnext_b200 import fused
    from vnext_b200.linear import linear
    g = torch.Generator().manual_seed(0)
    for C in (192, 768, 3072, 6144, 96):
        x = (torch.randn(777, C, generator=g) * 2 + 0.5).cuda()
        w, b = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
        want = F.layer_norm(x.double(), (C,), w.double(), b.double(), 1e-5)
        assert (fused.layernorm(x, w, b).double() - want).abs().max().item() < 1e-4
        o16 = fused.layernorm(x, w, b, out_dtype=torch.bfloat16)                       # fp32 stream -> bf16 GEMM operand
        assert o16.dtype == torch.bfloat16 and (o16.double() - want).abs().max().item() < 8e-3 * want.abs().max().item()
        x16 = x.bfloat16()
        want16 = F.layer_norm(x16.double(), (C,), w.double(), b.double(), 1e-5)
        assert (fused.layernorm(x16, w, b, out_dtype=torch.float32).double() - want16).abs().max().item() < 1e-4
    x = torch.randn(1000, 192, generator=g).cuda()
    w = (torch.randn(768, 192, generator=g) / 14).cuda()
    b = torch.randn(768, generator=g).cuda()
    want = F.gelu(F.linear(x.double(), w.double(), b.double()))
    assert (linear(x, w, b, gelu=True).double() - want).abs().max().item() < 5e-3     # TF32 operands


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("tf32", 3e-2), ("bf16", 5e-2)])
def test_swin_backbone_and_engines(precision, tol):
    from vnext_b200.idol.engine import IDOLEngine
    from vnext_b200.idol.state_spec import idol_r50_state_spec
    from vnext_b200.seqformer.engine import SeqFormerEngine
    from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
    golden = np.load(os.path.join(ROOT, "tests", "golden", "swin_golden.npz"))
    # backbone alone vs the oracle and vs the reference module's goldens
    sd = fill_state_dict(swin_state_spec("detr.detr.backbone.0.backbone.", **SMALL), 0)
    x = torch.randn(2, 3, 160, 224, generator=torch.Generator().manual_seed(50))
    seq_sd = fill_state_dict(seqformer_r50_state_spec(swin=SMALL), 0)
    eng = SeqFormerEngine(seq_sd, device="cuda", precision=precision, swin=SMALL)
    feats = eng.swin.forward(x.cuda())
    with torch.no_grad():
        want = WO.swin(seq_sd, x, "detr.detr.backbone.0.backbone.", **SMALL)
    for k in want:
        assert _rel(feats[k].permute(0, 3, 1, 2).cpu(), want[k]) < tol, k
    if precision == "fp32":     # and straight against the reference module's own outputs (weights keyed without prefix)
        from vnext_b200.swin import SwinBackbone

        class Host:             # native LayerNorm / window-attention kernels, cuBLAS fp32 GEMMs
            native = True

            @staticmethod
            def lin(x, wb, relu=False, residual=None, ln=None, gelu=False):
                y = F.linear(x, wb[0], wb[1])
                y = y if residual is None else y + residual
                return F.gelu(y) if gelu else y
        torch.backends.cuda.matmul.allow_tf32 = False
        sd0 = fill_state_dict(swin_state_spec("", **SMALL), 0)
        f0 = SwinBackbone(sd0, "", SMALL, torch.device("cuda"), Host).forward(x.cuda())
        for k in f0:
            assert _rel(f0[k].permute(0, 3, 1, 2).cpu()[:, ::4], torch.from_numpy(golden[f"w12_{k}"])) < 1e-3, k
    # whole models with the Swin backbone
    frames = synthetic_clip(2, 96, 128, 0)
    out = eng.inference(torch.stack(frames).cuda())
    with torch.no_grad():
        ref = SO.inference(seq_sd, frames, swin=SMALL)
    errs = {k: _rel(out[k].cpu(), ref[k]) for k in ("pred_logits", "pred_boxes", "pred_masks")}
    print("seqformer+swin", precision, errs)
    assert max(errs.values()) < tol
    idol_sd = fill_state_dict(idol_r50_state_spec(swin=SMALL), 1)
    out = IDOLEngine(idol_sd, device="cuda", precision=precision, swin=SMALL).inference_forward(torch.stack(frames).cuda())
    with torch.no_grad():
        ref = IO.inference_forward(idol_sd, frames, swin=SMALL)
    errs = {k: _rel(out[k].cpu(), ref[k]) for k in ("pred_logits", "pred_boxes", "pred_inst_embed", "pred_masks")}
    print("idol+swin", precision, errs)
    assert max(errs.values()) < tol
