"""Load-time weight packing for the SeqFormer engine: reference-layout state dict -> GPU-friendly tensors.

Everything here is exact algebra on the checkpoint (no numerics change beyond fp32 rounding of the
folded products), done once when a checkpoint is loaded, invisible to the checkpoint format
(SURVEY §7 "Keeping state-dict and config API identical while changing layouts").
  * FrozenBN folded into the preceding conv:  y = conv(x, w) * s + b,  s = gamma / sqrt(var + eps),
    b = beta - mean * s   (detectron2/layers/batch_norm.py:48-53, eps = 1e-5)
  * pixel normalisation folded into the stem:  conv((x - mean)/std, w) = conv(x, w/std) - sum(w*mean/std)
    — only valid without zero padding of the *normalised* image, so it is NOT applied (the reference pads
    the normalised frame with zeros, util/misc.py:306-312); normalisation stays a separate fused kernel.
  * sampling_offsets and attention_weights Linears concatenated into one [384, 256] projection.
"""
import torch

BB = "detr.detr.backbone.0.backbone."
STAGE_NAMES = ("res2", "res3", "res4", "res5")


def resnet_blocks(sd, prefix=BB):
    """(stage, number of bottleneck blocks) read off the checkpoint's keys: ResNet-50 (3,4,6,3), -101 (3,4,23,3), ..."""
    out = []
    for stage in STAGE_NAMES:
        n = 0
        while f"{prefix}{stage}.{n}.conv1.weight" in sd:
            n += 1
        if n == 0:
            raise KeyError(f"checkpoint has no {prefix}{stage}.0.conv1.weight")
        out.append((stage, n))
    return tuple(out)


def fold_conv_bn(sd, prefix, eps=1e-5):
    w = sd[prefix + ".weight"].float()
    n = prefix + ".norm."
    scale = sd[n + "weight"].float() * (sd[n + "running_var"].float() + eps).rsqrt()
    bias = sd[n + "bias"].float() - sd[n + "running_mean"].float() * scale
    return w * scale.view(-1, 1, 1, 1), bias


def pack_resnet50(sd, prefix=BB):
    """-> list of stages; each block = dict(conv1=(w,b), conv2=(w,b,stride), conv3=(w,b), shortcut=(w,b,stride)|None)."""
    packed = {"stem": fold_conv_bn(sd, prefix + "stem.conv1"), "stages": []}
    for stage, nblocks in resnet_blocks(sd, prefix):
        blocks = []
        for b in range(nblocks):
            p = f"{prefix}{stage}.{b}."
            stride = 2 if (b == 0 and stage != "res2") else 1
            blk = {"conv1": fold_conv_bn(sd, p + "conv1"), "conv2": fold_conv_bn(sd, p + "conv2"),
                   "conv3": fold_conv_bn(sd, p + "conv3"), "stride": stride, "shortcut": None}
            if (p + "shortcut.weight") in sd:
                blk["shortcut"] = fold_conv_bn(sd, p + "shortcut")
            blocks.append(blk)
        packed["stages"].append((stage, blocks))
    return packed


def pack_msda(sd, p, box):
    so_w, so_b = sd[p + ".sampling_offsets.weight"], sd[p + ".sampling_offsets.bias"]
    aw_w, aw_b = sd[p + ".attention_weights.weight"], sd[p + ".attention_weights.bias"]
    out = {
        "oa_w": torch.cat([so_w, aw_w], 0).float().contiguous(),      # [M*L*P*2 + M*L*P, C]
        "oa_b": torch.cat([so_b, aw_b], 0).float().contiguous(),
        "n_off": so_w.shape[0],
        "v_w": sd[p + ".value_proj.weight"].float(), "v_b": sd[p + ".value_proj.bias"].float(),
        "o_w": sd[p + ".output_proj.weight"].float(), "o_b": sd[p + ".output_proj.bias"].float(),
    }
    if box:
        # output_proj and output_proj_box consume the same sampled tensor -> one [2C, C] projection
        out["ob_w"] = torch.cat([sd[p + ".output_proj.weight"], sd[p + ".output_proj_box.weight"]], 0).float()
        out["ob_b"] = torch.cat([sd[p + ".output_proj.bias"], sd[p + ".output_proj_box.bias"]], 0).float()
    return out


def to_device(obj, device, dtype=None):
    if isinstance(obj, torch.Tensor):
        if obj.dtype.is_floating_point and dtype is not None:
            return obj.to(device=device, dtype=dtype)
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: to_device(v, device, dtype) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_device(v, device, dtype) for v in obj)
    return obj
