"""Swin backbone engine: same mathematics as the reference's `SwinTransformer.forward`
(projects/IDOL/idol/backbone/swin.py:640-662) from reference-layout weights, organised for the GPU:

  * tokens stay [T, H, W, C] (channels-last) for the whole backbone — no window_partition / window_reverse /
    torch.roll / F.pad copies: the fused window-attention kernel does the cyclic shift, the window gather, the
    zero-padded positions, the relative-position bias and the shift mask by index arithmetic (csrc/swin_attention.cu);
  * qkv / proj / fc1 / fc2 / patch-merging reductions are tcgen05 GEMMs with fused bias, exact-GELU and residual
    epilogues (csrc/gemm_tcgen05.cu); norm1 / norm2 / stage norms are the row-LayerNorm kernel;
  * the 4x4 stride-4 patch embedding is a GEMM over K = 48 on a channels-last patch view.
`lin` / `native` come from the owning engine (SeqFormerEngine), so precision modes and the CPU host-logic test path
(plain torch formulation, no kernels) are shared with the rest of the model.
"""
import torch
import torch.nn.functional as F


class SwinBackbone:
    def __init__(self, sd, prefix, cfg, device, engine):
        self.cfg, self.eng, self.device = cfg, engine, device
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        lin = lambda p: (f32(sd[p + ".weight"]), f32(sd[p + ".bias"]) if (p + ".bias") in sd else None)
        E, ps = cfg["embed_dim"], cfg.get("patch_size", 4)
        self.ps = ps
        w = sd[prefix + "patch_embed.proj.weight"]                     # [E, 3, ps, ps] -> [E, (kh, kw, c)]
        self.patch = (f32(w.permute(0, 2, 3, 1).reshape(E, -1)), f32(sd[prefix + "patch_embed.proj.bias"]))
        self.patch_norm = lin(prefix + "patch_embed.norm")
        self.stages = []
        for i, (depth, heads) in enumerate(zip(cfg["depths"], cfg["num_heads"])):
            blocks = []
            for j in range(depth):
                p = f"{prefix}layers.{i}.blocks.{j}."
                blocks.append({"norm1": lin(p + "norm1"), "qkv": lin(p + "attn.qkv"), "proj": lin(p + "attn.proj"),
                               "table": f32(sd[p + "attn.relative_position_bias_table"]), "norm2": lin(p + "norm2"),
                               "fc1": lin(p + "mlp.fc1"), "fc2": lin(p + "mlp.fc2"),
                               "shift": 0 if j % 2 == 0 else cfg["window_size"] // 2, "heads": heads})
            st = {"blocks": blocks, "norm": lin(f"{prefix}norm{i}"), "merge": None}
            if i < len(cfg["depths"]) - 1:
                p = f"{prefix}layers.{i}.downsample."
                st["merge"] = (lin(p + "norm"), (f32(sd[p + "reduction.weight"]), None))
            self.stages.append(st)

    # ------------------------------------------------------------------ pieces
    def _ln(self, x, wb, out_dtype=None):
        if self.eng.native:
            from .. import fused
            return fused.layernorm(x.contiguous(), wb[0], wb[1], out_dtype=out_dtype)
        return F.layer_norm(x, (x.shape[-1],), wb[0], wb[1], 1e-5)

    def _attention(self, qkv, blk):
        ws = self.cfg["window_size"]
        if self.eng.native:
            from .. import fused
            return fused.swin_window_attention(qkv.contiguous(), blk["qkv"][1], blk["table"], blk["heads"], ws, blk["shift"],
                                               tensor_cores=getattr(self.eng, "native_gemm", False))
        return _window_attention_torch(qkv, blk["qkv"][1], blk["table"], blk["heads"], ws, blk["shift"])

    def forward(self, x):
        """x: [T, 3, H, W] normalised frames (H, W multiples of 32) -> {"res2".."res5": [T, h, w, C] channels-last}."""
        T, _, H, W = x.shape
        ps, lin = self.ps, self.eng.lin
        # bf16 mode (engine precision "bf16" on the CUDA kernels): the residual stream `tok` stays fp32; LayerNorm hands bf16
        # rows to the bf16 GEMMs (qkv, proj, fc1, fc2, patch merging), qkv and the attention output are bf16 (the attention
        # kernel's MMAs are TF32 on operands widened in its gather), the GELU hidden is bf16
        act = getattr(self.eng, "bb_dtype", torch.float32) if self.eng.native else torch.float32
        mixed = act == torch.bfloat16
        f32 = {"out_dtype": torch.float32} if mixed else {}
        xh = x.permute(0, 2, 3, 1)
        h, w = H // ps, W // ps
        patches = xh.reshape(T, h, ps, w, ps, 3).permute(0, 1, 3, 2, 4, 5).reshape(T, h, w, ps * ps * 3)
        tok = self._ln(lin(patches, self.patch), self.patch_norm, **f32)
        feats = {}
        for i, st in enumerate(self.stages):
            for blk in st["blocks"]:
                att = self._attention(lin(self._ln(tok, blk["norm1"], act if mixed else None), blk["qkv"]), blk)
                tok = lin(att, blk["proj"], residual=tok, **f32)
                tok = lin(lin(self._ln(tok, blk["norm2"], act if mixed else None), blk["fc1"], gelu=True), blk["fc2"],
                          residual=tok, **f32)
            feats[f"res{i + 2}"] = self._ln(tok, st["norm"], act if mixed else None)
            if st["merge"] is not None:
                if h % 2 or w % 2:
                    tok = F.pad(tok, (0, 0, 0, w % 2, 0, h % 2))
                tok = torch.cat([tok[:, 0::2, 0::2], tok[:, 1::2, 0::2], tok[:, 0::2, 1::2], tok[:, 1::2, 1::2]], -1)
                h, w = (h + 1) // 2, (w + 1) // 2
                tok = lin(self._ln(tok, st["merge"][0], act if mixed else None), st["merge"][1], **f32)
        return feats


def _window_attention_torch(qkv, qkv_bias, table, heads, ws, shift):
    """Plain-torch formulation of the fused kernel's contract (CPU host-logic tests only):
    qkv [B,H,W,3C] of the un-padded, un-rolled tokens -> attention output [B,H,W,C]."""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    full = qkv_bias.view(1, 1, 1, C3).expand(B, Hp, Wp, C3).clone()     # padded positions carry the projection bias
    full[:, :H, :W] = qkv
    if shift:
        full = torch.roll(full, shifts=(-shift, -shift), dims=(1, 2))
    win = full.view(B, Hp // ws, ws, Wp // ws, ws, 3, heads, 32).permute(5, 0, 1, 3, 6, 2, 4, 7)
    win = win.reshape(3, B, Hp // ws, Wp // ws, heads, ws * ws, 32)
    q, k, v = win[0] * (32 ** -0.5), win[1], win[2]
    attn = q @ k.transpose(-1, -2)                                       # [B, nWy, nWx, heads, N, N]
    ar = torch.arange(ws, device=qkv.device)
    cy, cx = torch.meshgrid(ar, ar, indexing="ij")
    cy, cx = cy.reshape(-1), cx.reshape(-1)
    idx = (cy[:, None] - cy[None, :] + ws - 1) * (2 * ws - 1) + (cx[:, None] - cx[None, :] + ws - 1)
    attn = attn + table[idx.view(-1)].view(ws * ws, ws * ws, heads).permute(2, 0, 1)
    if shift:
        ys = torch.arange(Hp, device=qkv.device)
        xs = torch.arange(Wp, device=qkv.device)
        hr = (ys >= Hp - ws).long() + (ys >= Hp - shift).long()
        wr = (xs >= Wp - ws).long() + (xs >= Wp - shift).long()
        reg = (hr[:, None] * 3 + wr[None, :]).view(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(
            Hp // ws, Wp // ws, ws * ws)
        mask = (reg[..., :, None] != reg[..., None, :]).float() * -100.0
        attn = attn + mask[None, :, :, None]
    out = torch.softmax(attn, -1) @ v                                    # [B, nWy, nWx, heads, N, 32]
    out = out.view(B, Hp // ws, Wp // ws, heads, ws, ws, 32).permute(0, 1, 4, 2, 5, 3, 6).reshape(B, Hp, Wp, C)
    if shift:
        out = torch.roll(out, shifts=(shift, shift), dims=(1, 2))
    return out[:, :H, :W].contiguous()
