"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement, in plain functional PyTorch, of the reference's SeqFormer inference path
(SURVEY §3.2).  It is written directly against a reference-layout ``state_dict`` (same keys as the
reference module tree, SURVEY §8b "Checkpoint/state-dict boundary") and follows the reference
control flow literally — per-frame Python loops included — so that it is an independent statement
from the restructured CUDA engine in vnext_b200/.

What each function follows (paths under /root/reference/projects/SeqFormer/seqformer):
  pad_frames              util/misc.py:288-316 (nested_tensor_from_tensor_list, size_divisibility=32)
  resnet50                detectron2/modeling/backbone/resnet.py:194-210, 355-359, 435-458 and
                          detectron2/layers/batch_norm.py:44-65 (FrozenBN), STRIDE_IN_1X1=False
  sine_pos                models/position_encoding.py:36-56
  level masks             seqformer.py:46-55 (MaskedBackbone.forward), models/backbone.py:118-129
  input_proj / levels     models/segmentation_condInst.py:241-298, models/deformable_detr.py:63-84
  msda_module_*           models/ops/modules/ms_deform_attn.py:89-124 (encode), :128-217 (decode)
  encoder                 models/deformable_transformer.py:78-110, 132-198
  decoder                 models/deformable_transformer.py:112-129, 201-387
  heads + mask head       models/segmentation_condInst.py:300-351, 354-401, 404-493, 539-575, 614-678
  postprocess             seqformer.py:363-410 (whole_video_inference)
The MSDA core is oracle.msda_oracle.forward_torch (grid_sample; the reference's own CPU path).

Pinned by tests/test_seqformer_oracle_golden.py against outputs of the reference modules themselves
(tests/golden/make_seqformer_golden.py).  The ResNet part is pinned against torchvision's
resnet50 topology with the D2 key mapping (tools/convert-torchvision-to-d2.py:38-49) because
detectron2.modeling is not importable offline (needs fvcore) — SURVEY §8c.
"""
import math

import torch
import torch.nn.functional as F

from .msda_oracle import forward_torch as msda_core

BB = "detr.detr.backbone.0.backbone."
DETR = "detr.detr."
TR = "detr.detr.transformer."


# ----------------------------------------------------------------------------- input handling
def normalize_and_pad(frames, pixel_mean, pixel_std, div=32):
    """frames: list of [3,H,W] float tensors (0..255).  -> tensor [T,3,Hp,Wp], mask [T,Hp,Wp], (H,W)."""
    mean = torch.tensor(pixel_mean).view(3, 1, 1)
    std = torch.tensor(pixel_std).view(3, 1, 1)
    imgs = [(f - mean) / std for f in frames]
    H = max(i.shape[1] for i in imgs)
    W = max(i.shape[2] for i in imgs)
    Hp, Wp = (H + div - 1) // div * div, (W + div - 1) // div * div
    tensor = torch.zeros(len(imgs), 3, Hp, Wp)
    mask = torch.ones(len(imgs), Hp, Wp, dtype=torch.bool)
    for i, im in enumerate(imgs):
        tensor[i, :, :im.shape[1], :im.shape[2]] = im
        mask[i, :im.shape[1], :im.shape[2]] = False
    return tensor, mask, (imgs[0].shape[1], imgs[0].shape[2])


# ----------------------------------------------------------------------------- ResNet-50 (FrozenBN)
def _conv_bn(sd, prefix, x, stride=1, padding=0, relu=True, eps=1e-5):
    x = F.conv2d(x, sd[prefix + ".weight"], None, stride=stride, padding=padding)
    n = prefix + ".norm."
    x = F.batch_norm(x, sd[n + "running_mean"], sd[n + "running_var"], sd[n + "weight"], sd[n + "bias"],
                     training=False, eps=eps)
    return F.relu(x) if relu else x


def resnet50(sd, x, prefix=BB):
    out = {}
    x = _conv_bn(sd, prefix + "stem.conv1", x, stride=2, padding=3)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for stage in ("res2", "res3", "res4", "res5"):            # block count per stage read off the keys (R50 3,4,6,3; R101 3,4,23,3)
        nblocks = 0
        while f"{prefix}{stage}.{nblocks}.conv1.weight" in sd:
            nblocks += 1
        for b in range(nblocks):
            p = f"{prefix}{stage}.{b}."
            stride = 2 if (b == 0 and stage != "res2") else 1
            y = _conv_bn(sd, p + "conv1", x)
            y = _conv_bn(sd, p + "conv2", y, stride=stride, padding=1)
            y = _conv_bn(sd, p + "conv3", y, relu=False)
            sc = _conv_bn(sd, p + "shortcut", x, stride=stride, relu=False) if (p + "shortcut.weight") in sd else x
            x = F.relu(y + sc)
        out[stage] = x
    return out


# ----------------------------------------------------------------------------- position encoding
def sine_pos(mask, num_pos_feats=128, temperature=10000.0):
    not_mask = ~mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = (y_embed - 0.5) / (y_embed[:, -1:, :] + eps) * scale
    x_embed = (x_embed - 0.5) / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def _level_mask(mask, size):
    return F.interpolate(mask[None].float(), size=size).to(torch.bool)[0]


def _linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _mlp(sd, p, x, n):
    for i in range(n):
        x = _linear(sd, f"{p}.layers.{i}", x)
        if i < n - 1:
            x = F.relu(x)
    return x


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _mha(sd, p, q_in, k_in, v_in, heads=8):
    """nn.MultiheadAttention (batch-first here), eval mode.  q_in/k_in/v_in: [B, L, C]."""
    C = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q_in, w[:C], b[:C])
    k = F.linear(k_in, w[C:2 * C], b[C:2 * C])
    v = F.linear(v_in, w[2 * C:], b[2 * C:])
    B, Lq, _ = q.shape
    d = C // heads
    q = q.view(B, Lq, heads, d).transpose(1, 2) * (d ** -0.5)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, Lq, C)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


# ----------------------------------------------------------------------------- MSDeformAttn module (SeqFormer)
def _sampling_locations(ref, offsets, shapes, n_points):
    # ref [N, Lq, L, 2|4]; offsets [N, Lq, M, L, P, 2]
    if ref.shape[-1] == 2:
        normalizer = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
        return ref[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
    return ref[:, :, None, :, None, :2] + offsets / n_points * ref[:, :, None, :, None, 2:] * 0.5


def _value(sd, p, src, padding_mask, M):
    v = _linear(sd, p + ".value_proj", src)
    if padding_mask is not None:
        v = v.masked_fill(padding_mask[..., None], 0.0)
    return v.view(*v.shape[:-1], M, v.shape[-1] // M)


def msda_encode(sd, p, query, ref, src, shapes, lsi, padding_mask, M=8, L=4, P=4):
    N, nf, Lq, C = query.shape
    value = _value(sd, p, src, padding_mask, M)                      # [N, nf, S, M, D]
    off = _linear(sd, p + ".sampling_offsets", query).view(N, nf, Lq, M, L, P, 2)
    aw = torch.softmax(_linear(sd, p + ".attention_weights", query).view(N, nf, Lq, M, L * P), -1)
    aw = aw.view(N, nf, Lq, M, L, P)
    outs = []
    for f in range(nf):
        loc = _sampling_locations(ref, off[:, f], shapes, P)
        outs.append(msda_core(value[:, f].contiguous(), shapes, loc, aw[:, f].contiguous()).unsqueeze(1))
    return _linear(sd, p + ".output_proj", torch.cat(outs, 1))


def msda_decode(sd, p, query_box, ref, src, shapes, lsi, padding_mask, M=8, L=4, P=4):
    N, nf, S, C = src.shape
    value = _value(sd, p, src, padding_mask, M)
    first = query_box.dim() == 3
    Lq = query_box.shape[-2]
    if first:
        off = _linear(sd, p + ".sampling_offsets", query_box).view(N, Lq, M, L, P, 2)
        aw = torch.softmax(_linear(sd, p + ".attention_weights", query_box).view(N, Lq, M, L * P), -1)
        aw = aw.view(N, Lq, M, L, P)
    else:
        off = _linear(sd, p + ".sampling_offsets", query_box).view(N, nf, Lq, M, L, P, 2)
        aw = torch.softmax(_linear(sd, p + ".attention_weights", query_box).view(N, nf, Lq, M, L * P), -1)
        aw = aw.view(N, nf, Lq, M, L, P)
    outs = []
    for f in range(nf):
        o_f = off if first else off[:, f]
        a_f = aw if first else aw[:, f]
        loc = _sampling_locations(ref[:, f], o_f, shapes, P)
        outs.append(msda_core(value[:, f].contiguous(), shapes, loc.contiguous(), a_f.contiguous()).unsqueeze(1))
    res = torch.cat(outs, 1)
    return _linear(sd, p + ".output_proj", res), _linear(sd, p + ".output_proj_box", res)


# ----------------------------------------------------------------------------- encoder / decoder
def encoder_reference_points(shapes, valid_ratios):
    refs = []
    for lvl, (H, W) in enumerate(shapes.tolist()):
        ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
        ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H)
        ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W)
        refs.append(torch.stack((ref_x, ref_y), -1))
    ref = torch.cat(refs, 1)
    return ref[:, :, None] * valid_ratios[:, None]


def _ffn(sd, p, x, suffix=""):
    return _linear(sd, f"{p}.linear2{suffix}", F.relu(_linear(sd, f"{p}.linear1{suffix}", x)))


def encoder(sd, src, shapes, lsi, valid_ratios, pos, padding_mask, num_layers=6):
    out = src
    ref = encoder_reference_points(shapes, valid_ratios)
    for i in range(num_layers):
        p = f"{TR}encoder.layers.{i}"
        a = msda_encode(sd, p + ".self_attn", out + pos, ref, out, shapes, lsi, padding_mask)
        out = _ln(sd, p + ".norm1", out + a)
        out = _ln(sd, p + ".norm2", out + _ffn(sd, p, out))
    return out


def decoder(sd, tgt, reference_points, memory, shapes, lsi, valid_ratios, query_pos, padding_mask, num_layers=6):
    output, output_box = tgt, tgt
    inter, inter_box, inter_ref = [], [], []
    for lid in range(num_layers):
        p = f"{TR}decoder.layers.{lid}"
        if reference_points.shape[-1] == 4:
            ref_in = reference_points[:, :, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[:, None, None]
        else:
            ref_in = reference_points[:, :, :, None] * valid_ratios[:, None, None]
        # mask/class query self-attention
        q = output + query_pos
        output = _ln(sd, p + ".norm2", output + _mha(sd, p + ".self_attn", q, q, output))
        # box query self-attention (per frame after the first layer)
        if output_box.dim() == 3:
            qb = output_box + query_pos
            output_box = _ln(sd, p + ".norm2_box", output_box + _mha(sd, p + ".self_attn_box", qb, qb, output_box))
            t2, t2_box = msda_decode(sd, p + ".cross_attn", output_box + query_pos, ref_in, memory, shapes, lsi,
                                     padding_mask)
            output_box = output_box.unsqueeze(1) + t2_box
        else:
            frames = []
            for f in range(output_box.shape[1]):
                tb = output_box[:, f]
                qb = tb + query_pos
                frames.append(_ln(sd, p + ".norm2_box", tb + _mha(sd, p + ".self_attn_box", qb, qb, tb)).unsqueeze(1))
            output_box = torch.cat(frames, 1)
            t2, t2_box = msda_decode(sd, p + ".cross_attn", output_box + query_pos.unsqueeze(1), ref_in, memory,
                                     shapes, lsi, padding_mask)
            output_box = output_box + t2_box
        output_box = _ln(sd, p + ".norm1_box", output_box)
        output_box = _ln(sd, p + ".norm3_box", output_box + _ffn(sd, p, output_box, "_box"))
        tw = torch.softmax(_linear(sd, p + ".time_attention_weights", output_box), 1)
        output = _ln(sd, p + ".norm1", output + (t2 * tw).sum(1))
        output = _ln(sd, p + ".norm3", output + _ffn(sd, p, output))
        # iterative box refinement (with_box_refine=True)
        tmp = _mlp(sd, f"{DETR}bbox_embed.{lid}", output_box, 3)
        if reference_points.shape[-1] == 4:
            new_ref = (tmp + inverse_sigmoid(reference_points)).sigmoid()
        else:
            tmp = tmp.clone()
            tmp[..., :2] = tmp[..., :2] + inverse_sigmoid(reference_points)
            new_ref = tmp.sigmoid()
        reference_points = new_ref
        inter.append(output)
        inter_box.append(output_box)
        inter_ref.append(reference_points)
    return torch.stack(inter), torch.stack(inter_box), torch.stack(inter_ref)


# ----------------------------------------------------------------------------- mask head
def _conv(sd, p, x, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=padding)


def mask_branch(sd, feats):
    """MaskHeadSmallConv.forward with fpns=None.  feats: [level0, level1, level2] each [B, C, h, w]."""
    p = "detr.mask_head."
    x = F.relu(_conv(sd, p + "lay3", feats[-1]))
    x = feats[-2] + F.interpolate(x, size=feats[-2].shape[-2:], mode="nearest")
    x = F.relu(_conv(sd, p + "lay4", x))
    x = feats[-3] + F.interpolate(x, size=feats[-3].shape[-2:], mode="nearest")
    x = F.relu(_conv(sd, p + "dcn", x))
    x = F.relu(_conv(sd, p + "lay1", x))
    return F.relu(_conv(sd, p + "lay2", x))


def aligned_bilinear(t, factor):
    if factor == 1:
        return t
    h, w = t.shape[2:]
    t = F.pad(t, pad=(0, 1, 0, 1), mode="replicate")
    oh, ow = factor * h + 1, factor * w + 1
    t = F.interpolate(t, size=(oh, ow), mode="bilinear", align_corners=True)
    t = F.pad(t, pad=(factor // 2, 0, factor // 2, 0), mode="replicate")
    return t[:, :, :oh - 1, :ow - 1]


def dynamic_mask_head(mask_feats, ref_points, params, stride=8, out_stride=4, ch=8):
    """mask_feats [1, 8, H, W]; ref_points [1, n, 2] (x, y in pixels); params [1, n, 169] -> [1, n, 2H, 2W]."""
    _, cin, H, W = mask_feats.shape
    n = ref_points.shape[1]
    xs = torch.arange(0, W * stride, stride, dtype=torch.float32) + stride // 2
    ys = torch.arange(0, H * stride, stride, dtype=torch.float32) + stride // 2
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    locations = torch.stack((xx.reshape(-1), yy.reshape(-1)), 1)            # [H*W, 2]
    rel = ref_points.reshape(1, n, 1, 1, 2) - locations.reshape(1, 1, H, W, 2)
    rel = rel.permute(0, 1, 4, 2, 3).flatten(-2, -1)                        # [1, n, 2, H*W]
    feats = mask_feats[0].reshape(1, cin, H * W).unsqueeze(1).repeat(1, n, 1, 1)
    x = torch.cat([rel, feats], dim=2).reshape(1, -1, H, W)                 # [1, n*10, H, W]
    prm = params.flatten(0, 1)
    wn, bn = [(cin + 2) * ch, ch * ch, ch], [ch, ch, 1]
    parts = list(torch.split_with_sizes(prm, wn + bn, dim=1))
    ws, bs = parts[:3], parts[3:]
    ws = [ws[0].reshape(n * ch, -1, 1, 1), ws[1].reshape(n * ch, -1, 1, 1), ws[2].reshape(n, -1, 1, 1)]
    bs = [bs[0].reshape(n * ch), bs[1].reshape(n * ch), bs[2].reshape(n)]
    for i in range(3):
        x = F.conv2d(x, ws[i], bias=bs[i], groups=n)
        if i < 2:
            x = F.relu(x)
    x = aligned_bilinear(x.reshape(-1, 1, H, W), stride // out_stride)
    return x.reshape(1, -1, x.shape[-2], x.shape[-1])


# ----------------------------------------------------------------------------- full inference
def inference(sd, frames, pixel_mean=(123.675, 116.280, 103.530), pixel_std=(58.395, 57.120, 57.375),
              num_queries=None, enc_layers=6, dec_layers=6, return_intermediates=False, swin=None):
    """CondInst_segm.inference for ONE video (bs = 1) of T = len(frames) frames."""
    tensor, mask, image_size = normalize_and_pad(frames, pixel_mean, pixel_std)
    T = tensor.shape[0]
    if swin is not None:          # D2SwinTransformer backbone instead of ResNet-50 (cfg dict, see vnext_b200.swin_spec)
        from . import swin_oracle
        feats = swin_oracle.swin(sd, tensor, BB, **swin)
    else:
        feats = resnet50(sd, tensor)
    srcs, masks, poses = [], [], []
    for l, name in enumerate(("res3", "res4", "res5")):
        x = feats[name]
        m = _level_mask(mask, x.shape[-2:])
        src = F.conv2d(x, sd[f"{DETR}input_proj.{l}.0.weight"], sd[f"{DETR}input_proj.{l}.0.bias"])
        src = F.group_norm(src, 32, sd[f"{DETR}input_proj.{l}.1.weight"], sd[f"{DETR}input_proj.{l}.1.bias"], 1e-5)
        srcs.append(src[None])                        # [1, T, C, h, w]
        masks.append(m[None])
        poses.append(sine_pos(m)[None])
    src = F.conv2d(feats["res5"], sd[f"{DETR}input_proj.3.0.weight"], sd[f"{DETR}input_proj.3.0.bias"], stride=2,
                   padding=1)
    src = F.group_norm(src, 32, sd[f"{DETR}input_proj.3.1.weight"], sd[f"{DETR}input_proj.3.1.bias"], 1e-5)
    m = _level_mask(mask, src.shape[-2:])
    srcs.append(src[None]); masks.append(m[None]); poses.append(sine_pos(m)[None])

    # DeformableTransformer.forward: flatten levels
    shapes = torch.as_tensor([s.shape[-2:] for s in srcs], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    src_flat = torch.cat([s.flatten(3).transpose(2, 3) for s in srcs], 2)           # [1, T, S, C]
    mask_flat = torch.cat([mm.flatten(2) for mm in masks], 2)                        # [1, T, S]
    pos_flat = torch.cat([ps.flatten(3).transpose(2, 3) + sd[TR + "level_embed"][l].view(1, 1, 1, -1)
                          for l, ps in enumerate(poses)], 2)

    def valid_ratio(mk):                                                             # mk [1, h, w] (frame 0)
        _, H, W = mk.shape
        vh = torch.sum(~mk[:, :, 0], 1).float() / H
        vw = torch.sum(~mk[:, 0, :], 1).float() / W
        return torch.stack([vw, vh], -1)
    valid_ratios = torch.stack([valid_ratio(mm[:, 0]) for mm in masks], 1)           # [1, L, 2]

    memory = encoder(sd, src_flat, shapes, lsi, valid_ratios, pos_flat, mask_flat, enc_layers)

    qe = sd[DETR + "query_embed.weight"]
    if num_queries is not None:
        qe = qe[:num_queries]
    C = memory.shape[-1]
    query_pos, tgt = torch.split(qe, C, dim=1)
    query_pos, tgt = query_pos.unsqueeze(0), tgt.unsqueeze(0)
    init_ref = _linear(sd, TR + "reference_points", query_pos).sigmoid().unsqueeze(1).repeat(1, T, 1, 1)
    hs, hs_box, inter_ref = decoder(sd, tgt, init_ref, memory, shapes, lsi, valid_ratios, query_pos, mask_flat,
                                    dec_layers)

    lvl = dec_layers - 1           # only the last level is returned (segmentation_condInst.py:345-349)
    reference = inverse_sigmoid(init_ref if lvl == 0 else inter_ref[lvl - 1])
    logits = _linear(sd, f"{DETR}class_embed.{lvl}", hs[lvl])
    tmp = _mlp(sd, f"{DETR}bbox_embed.{lvl}", hs_box[lvl], 3)
    if reference.shape[-1] == 4:
        tmp = tmp + reference
    else:
        tmp = tmp.clone()
        tmp[..., :2] += reference
    boxes = tmp.sigmoid()
    params = _mlp(sd, "detr.controller", hs[lvl], 3)                                # [1, nq, 169]
    scale_f = torch.tensor([float(image_size[1]), float(image_size[0])])
    ref_px = reference[..., :2].sigmoid() * scale_f[None, None, None, :]            # [1, T, nq, 2]

    # forward_mask_head_train: mask features from encoder memory levels 0..2, per frame
    per_frame = []
    for f in range(T):
        lv = []
        for l in range(3):
            h, w = shapes[l].tolist()
            s0 = int(lsi[l])
            lv.append(memory[:, f, s0:s0 + h * w, :].reshape(1, h, w, C).permute(0, 3, 1, 2))
        feat = mask_branch(sd, lv)
        per_frame.append(dynamic_mask_head(feat, ref_px[:, f], params).unsqueeze(2))
    pred_masks = torch.cat(per_frame, dim=2)                                        # [1, nq, T, H/4, W/4]
    out = {"pred_logits": logits, "pred_boxes": boxes, "pred_masks": pred_masks, "image_size": image_size}
    if return_intermediates:
        out.update(memory=memory, hs=hs, hs_box=hs_box, inter_ref=inter_ref, feats=feats,
                   src_flat=src_flat, pos_flat=pos_flat, valid_ratios=valid_ratios)
    return out


def whole_video_postprocess(out, ori_size, mask_stride=4, apply_cls_thres=0.05, multi_cls=True, topk=10):
    """seqformer.py:363-410."""
    logits = out["pred_logits"][0]
    output_mask = out["pred_masks"][0]
    image_size = out["image_size"]
    oh, ow = output_mask.shape[-2:]
    k = min(topk, logits.shape[0])
    _, idx = torch.topk(logits.sigmoid().max(1)[0], k=k)
    idx = idx.tolist()
    valid_logits = logits[idx].sigmoid()
    output_mask = output_mask[idx]
    pm = F.interpolate(output_mask, size=(oh * mask_stride, ow * mask_stride), mode="bilinear",
                       align_corners=False).sigmoid()
    if multi_cls:
        above = torch.where(valid_logits > apply_cls_thres)
        scores, labels = valid_logits[above], above[1]
        pm = pm[above[0]]
    else:
        scores, labels = valid_logits.max(-1)
    pm = pm[:, :, :image_size[0], :image_size[1]]
    pm = F.interpolate(pm, size=(ori_size[0], ori_size[1]), mode="nearest")
    masks = pm > 0.5
    return {"image_size": ori_size, "pred_scores": scores.tolist(), "pred_labels": labels.tolist(),
            "pred_masks": [m for m in masks]}
