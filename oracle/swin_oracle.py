"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (functional PyTorch) of the reference's Swin Transformer backbone
(projects/IDOL/idol/backbone/swin.py, identical under projects/SeqFormer): PatchEmbed :462-505, BasicLayer incl. the
shifted-window mask :404-451, SwinTransformerBlock :233-293 (pad AFTER norm1, cyclic roll, window partition/reverse),
WindowAttention :129-169 (scaled QK^T + relative-position bias gather + shift mask, softmax, PV, proj),
Mlp :33-39 (exact GELU), PatchMerging :309-335, per-stage output norms + NCHW outputs :640-662.
Eval mode: DropPath / dropout are identities.  Pinned by tests/test_swin_golden.py against the reference module's own
outputs (tests/golden/make_swin_golden.py)."""
import numpy as np
import torch
import torch.nn.functional as F


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_position_index(ws):
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij"))
    flat = torch.flatten(coords, 1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_mask(H, W, ws, shift):
    Hp, Wp = int(np.ceil(H / ws)) * ws, int(np.ceil(W / ws)) * ws
    img = torch.zeros((1, Hp, Wp, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


def window_attention(sd, p, x, heads, ws, mask):
    B_, N, C = x.shape
    qkv = _lin(sd, p + ".qkv", x).reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * ((C // heads) ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    idx = sd.get(p + ".relative_position_index")
    if idx is None:
        idx = relative_position_index(ws)
    bias = sd[p + ".relative_position_bias_table"][idx.view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    attn = torch.softmax(attn, -1)
    return _lin(sd, p + ".proj", (attn @ v).transpose(1, 2).reshape(B_, N, C))


def block(sd, p, x, H, W, heads, ws, shift, mask):
    B, L, C = x.shape
    shortcut = x
    x = _ln(sd, p + ".norm1", x).view(B, H, W, C)
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = x.shape[1], x.shape[2]
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = window_partition(x, ws).view(-1, ws * ws, C)
    aw = window_attention(sd, p + ".attn", xw, heads, ws, mask if shift > 0 else None)
    x = window_reverse(aw.view(-1, ws, ws, C), ws, Hp, Wp)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    x = x[:, :H, :W, :].contiguous().view(B, H * W, C)
    x = shortcut + x
    return x + _lin(sd, p + ".mlp.fc2", F.gelu(_lin(sd, p + ".mlp.fc1", _ln(sd, p + ".norm2", x))))


def patch_merging(sd, p, x, H, W):
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    if H % 2 == 1 or W % 2 == 1:
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    x = x.view(B, -1, 4 * C)
    return _lin(sd, p + ".reduction", _ln(sd, p + ".norm", x))


def swin(sd, x, prefix, embed_dim, depths, num_heads, window_size, patch_size=4):
    """x [B,3,H,W] normalised -> {"res2".."res5": [B,C_i,H_i,W_i]} (swin.py:640-662)."""
    _, _, H, W = x.shape
    if W % patch_size:
        x = F.pad(x, (0, patch_size - W % patch_size))
    if H % patch_size:
        x = F.pad(x, (0, 0, 0, patch_size - H % patch_size))
    x = F.conv2d(x, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], stride=patch_size)
    Wh, Ww = x.shape[2], x.shape[3]
    x = _ln(sd, prefix + "patch_embed.norm", x.flatten(2).transpose(1, 2))
    outs = {}
    for i, (depth, heads) in enumerate(zip(depths, num_heads)):
        C = embed_dim * 2 ** i
        shift = window_size // 2
        mask = shift_mask(Wh, Ww, window_size, shift)
        for j in range(depth):
            x = block(sd, f"{prefix}layers.{i}.blocks.{j}", x, Wh, Ww, heads, window_size, 0 if j % 2 == 0 else shift, mask)
        out = _ln(sd, f"{prefix}norm{i}", x)
        outs[f"res{i + 2}"] = out.view(-1, Wh, Ww, C).permute(0, 3, 1, 2).contiguous()
        if i < len(depths) - 1:
            x = patch_merging(sd, f"{prefix}layers.{i}.downsample", x, Wh, Ww)
            Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
    return outs
