"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (functional PyTorch) of IDOL's per-frame dense path
`CondInst_segm.inference_forward` (projects/IDOL/idol/models/segmentation_condInst.py:233-321) on a reference-layout
state dict.  What each piece follows (paths under /root/reference/projects/IDOL/idol):
  msda_module   models/ops/modules/ms_deform_attn.py:78-116
  encoder       models/deformable_transformer.py:138-170, 200-269
  decoder       models/deformable_transformer.py:172-197, 272-384 (single query stream; iterative box refinement)
  heads         models/segmentation_condInst.py:279-320 (class / box / reid_embed_head / controller, mask head :324-374)
Backbone, position encoding, mask branch, dynamic mask head and helpers are shared with oracle.seqformer_oracle (the
reference files are the same code in both projects).  The online tracker (models/tracker.py) is out of scope.
Pinned by tests/test_idol_golden.py against the reference modules' own outputs (tests/golden/make_idol_golden.py).
"""
import torch
import torch.nn.functional as F

from . import seqformer_oracle as SO
from .msda_oracle import forward_torch as msda_core

DETR = "detr.detr."
TR = "detr.detr.transformer."


def msda_module(sd, p, query, ref, src, shapes, padding_mask, M=8, L=4, P=4):
    N, Lq, C = query.shape
    value = SO._value(sd, p, src, padding_mask, M)
    off = SO._linear(sd, p + ".sampling_offsets", query).view(N, Lq, M, L, P, 2)
    aw = torch.softmax(SO._linear(sd, p + ".attention_weights", query).view(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    loc = SO._sampling_locations(ref, off, shapes, P)
    return SO._linear(sd, p + ".output_proj", msda_core(value, shapes, loc, aw))


def encoder(sd, src, shapes, valid_ratios, pos, padding_mask, num_layers=6):
    out = src
    ref = SO.encoder_reference_points(shapes, valid_ratios)
    for i in range(num_layers):
        p = f"{TR}encoder.layers.{i}"
        out = SO._ln(sd, p + ".norm1", out + msda_module(sd, p + ".self_attn", out + pos, ref, out, shapes, padding_mask))
        out = SO._ln(sd, p + ".norm2", out + SO._ffn(sd, p, out))
    return out


def decoder(sd, tgt, reference_points, memory, shapes, valid_ratios, query_pos, padding_mask, num_layers=6):
    output = tgt
    inter, inter_ref = [], []
    for lid in range(num_layers):
        p = f"{TR}decoder.layers.{lid}"
        if reference_points.shape[-1] == 4:
            ref_in = reference_points[:, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[:, None]
        else:
            ref_in = reference_points[:, :, None] * valid_ratios[:, None]
        q = output + query_pos
        output = SO._ln(sd, p + ".norm2", output + SO._mha(sd, p + ".self_attn", q, q, output))
        t2 = msda_module(sd, p + ".cross_attn", output + query_pos, ref_in, memory, shapes, padding_mask)
        output = SO._ln(sd, p + ".norm1", output + t2)
        output = SO._ln(sd, p + ".norm3", output + SO._ffn(sd, p, output))
        tmp = SO._mlp(sd, f"{DETR}bbox_embed.{lid}", output, 3)
        if reference_points.shape[-1] == 4:
            reference_points = (tmp + SO.inverse_sigmoid(reference_points)).sigmoid()
        else:
            tmp = tmp.clone()
            tmp[..., :2] = tmp[..., :2] + SO.inverse_sigmoid(reference_points)
            reference_points = tmp.sigmoid()
        inter.append(output)
        inter_ref.append(reference_points)
    return torch.stack(inter), torch.stack(inter_ref)


def inference_forward(sd, frames, pixel_mean=(123.675, 116.280, 103.530), pixel_std=(58.395, 57.120, 57.375),
                      enc_layers=6, dec_layers=6, return_intermediates=False, swin=None):
    """frames: list of N [3,H,W] float tensors (0..255) = the frames of one video, processed as a batch."""
    tensor, mask, image_size = SO.normalize_and_pad(frames, pixel_mean, pixel_std)
    N = tensor.shape[0]
    if swin is not None:
        from . import swin_oracle
        feats = swin_oracle.swin(sd, tensor, SO.BB, **swin)
    else:
        feats = SO.resnet50(sd, tensor)
    srcs, masks, poses = [], [], []
    for l, name in enumerate(("res3", "res4", "res5")):
        x = feats[name]
        m = SO._level_mask(mask, x.shape[-2:])
        src = F.conv2d(x, sd[f"{DETR}input_proj.{l}.0.weight"], sd[f"{DETR}input_proj.{l}.0.bias"])
        srcs.append(F.group_norm(src, 32, sd[f"{DETR}input_proj.{l}.1.weight"], sd[f"{DETR}input_proj.{l}.1.bias"], 1e-5))
        masks.append(m)
        poses.append(SO.sine_pos(m))
    src = F.conv2d(feats["res5"], sd[f"{DETR}input_proj.3.0.weight"], sd[f"{DETR}input_proj.3.0.bias"], stride=2, padding=1)
    src = F.group_norm(src, 32, sd[f"{DETR}input_proj.3.1.weight"], sd[f"{DETR}input_proj.3.1.bias"], 1e-5)
    m = F.interpolate(masks[0][None].float(), size=src.shape[-2:]).to(torch.bool)[0]      # from masks[0] (segm:262)
    srcs.append(src); masks.append(m); poses.append(SO.sine_pos(m))

    shapes = torch.as_tensor([s.shape[-2:] for s in srcs], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    src_flat = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)                 # [N, S, C]
    mask_flat = torch.cat([mm.flatten(1) for mm in masks], 1)
    pos_flat = torch.cat([ps.flatten(2).transpose(1, 2) + sd[TR + "level_embed"][l].view(1, 1, -1)
                          for l, ps in enumerate(poses)], 1)

    def valid_ratio(mk):
        _, H, W = mk.shape
        return torch.stack([torch.sum(~mk[:, 0, :], 1).float() / W, torch.sum(~mk[:, :, 0], 1).float() / H], -1)
    valid_ratios = torch.stack([valid_ratio(mm) for mm in masks], 1)                      # [N, L, 2]
    memory = encoder(sd, src_flat, shapes, valid_ratios, pos_flat, mask_flat, enc_layers)

    qe = sd[DETR + "query_embed.weight"]
    C = memory.shape[-1]
    query_pos, tgt = torch.split(qe, C, dim=1)
    query_pos, tgt = query_pos.unsqueeze(0).expand(N, -1, -1), tgt.unsqueeze(0).expand(N, -1, -1)
    init_ref = SO._linear(sd, TR + "reference_points", query_pos).sigmoid()
    hs, inter_ref = decoder(sd, tgt, init_ref, memory, shapes, valid_ratios, query_pos, mask_flat, dec_layers)

    lvl = dec_layers - 1
    reference = SO.inverse_sigmoid(inter_ref[-2])
    logits = SO._linear(sd, f"{DETR}class_embed.{lvl}", hs[-1])
    tmp = SO._mlp(sd, f"{DETR}bbox_embed.{lvl}", hs[-1], 3)
    tmp = tmp + reference if reference.shape[-1] == 4 else torch.cat([tmp[..., :2] + reference, tmp[..., 2:]], -1)
    boxes = tmp.sigmoid()
    inst_embed = SO._mlp(sd, "detr.reid_embed_head", hs[-1], 3)
    params = SO._mlp(sd, "detr.controller", hs[-1], 3)                                    # [N, nq, 169]
    scale_f = torch.tensor([float(image_size[1]), float(image_size[0])])
    ref_px = inter_ref[-2][..., :2] * scale_f[None, None, :]                              # [N, nq, 2]
    masks_out = []
    for i in range(N):
        lv = []
        for l in range(3):
            h, w = shapes[l].tolist()
            s0 = int(lsi[l])
            lv.append(memory[i:i + 1, s0:s0 + h * w, :].reshape(1, h, w, C).permute(0, 3, 1, 2))
        feat = SO.mask_branch(sd, lv)
        masks_out.append(SO.dynamic_mask_head(feat, ref_px[i:i + 1], params[i:i + 1]).unsqueeze(2))
    out = {"pred_logits": logits, "pred_boxes": boxes, "pred_inst_embed": inst_embed,
           "pred_masks": torch.cat(masks_out, 0), "reference_points": inter_ref[-2][..., :2], "image_size": image_size}
    if return_intermediates:
        out.update(memory=memory, hs=hs)
    return out
