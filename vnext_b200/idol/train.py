"""IDOL training step (SURVEY §8d cfg 3): the dense forward of `CondInst_segm.forward(..., train=True)`
(projects/IDOL/idol/models/segmentation_condInst.py:77-231) on a reference-layout state dict, differentiable, with
every multi-scale deformable attention (6 encoder + 6 decoder layers, key and reference frames) going through this
repo's sm_100a forward/backward kernels (`vnext_b200.msda.MSDeformAttnFunction`, C ABI vnext_msda_forward/backward).

What runs where this round:
  * MSDA forward + backward .................. hand-written kernels (csrc/msda_forward.cu, csrc/msda_backward.cu)
  * AdamW over the flat parameter buffer ..... hand-written kernel (csrc/optimizer.cu, C ABI vnext_adamw)
  * convolutions / linears and their grads ... cuDNN / cuBLAS through torch autograd (the reference's own training
    path; replacing their dgrad/wgrad with tcgen05 kernels is next-round work)
  * losses, simOTA, contrastive selection .... host Python/PyTorch (criterion.py), as the north star prescribes
  * gradient exchange ........................ ONE all-reduce of the flat gradient buffer (NCCL over NVLink)

B200-first structure rather than the reference's: key and reference frames go through backbone + transformer as ONE
batch (the reference runs the transformer twice, segm:140-141; samples are independent so values are identical);
FrozenBN is folded into the conv weights on the fly (no per-conv affine pass over the activations); the mask branch
is evaluated once per step, not once per decoder level (segm:190, identical values, gradients accumulate); all
trainable parameters and their gradients live in two flat fp32 buffers (parameters are views), so gradient averaging
is one collective and the optimiser is one launch per learning-rate group.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import _cabi
from ..linear import linear as tc_linear
from ..msda import MSDeformAttnFunction
from . import criterion as crit
from . import criterion_batched as critb
from .state_spec import idol_r50_state_spec

BB = "detr.detr.backbone.0.backbone."
DETR = "detr.detr."
TR = "detr.detr.transformer."
ALIAS = TR + "decoder.bbox_embed."            # same Parameters as detr.detr.bbox_embed.* (deformable_detr.py:101-104)


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _is_frozen(key, freeze_at):
    if not key.startswith(BB):
        return False
    if ".norm." in key:                         # FrozenBatchNorm2d: all four tensors are buffers (batch_norm.py:36-42)
        return True
    stage = key[len(BB):].split(".")[0]                     # FREEZE_AT=k freezes stem and res2..res(k) (resnet.py:613-650)
    return stage in (["stem"] + [f"res{i}" for i in range(2, 6)])[:freeze_at]


class FlatParams:
    """Trainable tensors as views into one fp32 buffer per role (weights, grads, Adam moments), grouped by learning
    rate: group 0 = backbone (BASE_LR x BACKBONE_MULTIPLIER, train_net.py:104-108), group 1 = everything else."""

    ALIGN = 64

    def __init__(self, tensors, device):
        keys = sorted(tensors, key=lambda k: (0 if "backbone" in k else 1, k))
        self.offsets, n = {}, 0
        self.group_end = [0, 0]
        for k in keys:
            self.offsets[k] = (n, tensors[k].numel(), tuple(tensors[k].shape))
            n += (tensors[k].numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            self.group_end[0 if "backbone" in k else 1] = n
        if self.group_end[1] < self.group_end[0]:
            self.group_end[1] = self.group_end[0]
        self.numel = n
        self.w = torch.zeros(n, dtype=torch.float32, device=device)
        self.g = torch.zeros_like(self.w)
        self.m = torch.zeros_like(self.w)
        self.v = torch.zeros_like(self.w)
        self.params = {}
        for k in keys:
            o, cnt, shape = self.offsets[k]
            self.w[o:o + cnt].copy_(tensors[k].reshape(-1).to(device=device, dtype=torch.float32))
            p = nn.Parameter(self.w[o:o + cnt].view(shape))
            p.grad = self.g[o:o + cnt].view(shape)          # autograd accumulates in place into the flat buffer
            self.params[k] = p

    def groups(self):
        return [(0, self.group_end[0]), (self.group_end[0], self.group_end[1])]


def adamw_(w, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale):
    """One fused AdamW update of a flat fp32 segment (C ABI vnext_adamw; torch.optim.AdamW semantics).  `grad_scale`
    is a 1-element device tensor (clip coefficient / world size) so no host sync sits between backward and update."""
    if not w.is_cuda:
        raise RuntimeError("vnext_adamw: Not implemented on the CPU")
    if w.numel() == 0:
        return
    rc = _cabi.lib().vnext_adamw(w.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), w.numel(), float(lr),
                                 float(beta1), float(beta2), float(eps), float(weight_decay),
                                 float(1.0 - beta1 ** step), float(1.0 - beta2 ** step), grad_scale.data_ptr(),
                                 torch.cuda.current_stream(w.device).cuda_stream)
    _cabi.check(rc, "vnext_adamw")


class TcLinear(torch.autograd.Function):
    """y = relu?(x W^T + b) with the forward GEMM and the data-gradient GEMM (dX = dY W) on this repo's tcgen05 TF32
    kernel (csrc/gemm_tcgen05.cu, bias / ReLU fused in the epilogue); the weight gradient dW = dY^T X contracts over the
    row dimension, which that kernel would need both operands transposed for, and stays on cuBLAS this round."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x2 = x.reshape(-1, x.shape[-1])
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        out = tc_linear(x2, weight.contiguous(), bias, relu=relu)
        ctx.relu = relu
        ctx.save_for_backward(x2, weight, out if relu else None)
        ctx.has_bias = bias is not None
        return out.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, grad):
        x2, weight, out = ctx.saved_tensors
        g = grad.reshape(-1, grad.shape[-1])
        if ctx.relu:
            g = g * (out > 0)
        g = g if g.is_contiguous() else g.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if weight.shape[0] % 4 == 0:
                gx = tc_linear(g, weight.t().contiguous())
            else:
                gx = g @ weight
            gx = gx.view(*grad.shape[:-1], weight.shape[1])
        if ctx.needs_input_grad[1]:
            gw = g.t() @ x2
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g.sum(0)
        return gx, gw, gb, None


class IDOLTrainModel(nn.Module):
    def __init__(self, state_dict, device="cuda", num_classes=40, num_queries=300, dropout=0.1, freeze_at=2,
                 pixel_mean=(123.675, 116.280, 103.530), pixel_std=(58.395, 57.120, 57.375), mask_stride=4,
                 weight_dict=None, heads=8, levels=4, points=4, precision="fp32", batched=True,
                 channels_last=True):
        super().__init__()
        if precision not in ("fp32", "tf32"):
            raise ValueError("precision must be 'fp32' (cuBLAS fp32 linears, strict parity) or 'tf32' (tcgen05 linears)")
        self.tc = precision == "tf32"
        self.cl = channels_last               # NHWC activations / weights for cuDNN; tokens become free views of them
        if self.tc:                                   # weight-gradient GEMMs (cuBLAS) in TF32 as well; process-global flag
            torch.backends.cuda.matmul.allow_tf32 = True
        self.device = torch.device(device)
        from ..seqformer.state_spec import resnet_depth_of
        spec = self._state_spec(num_queries, num_classes, resnet_depth_of(state_dict))
        missing = [k for k in spec if k not in state_dict]
        if missing:
            raise KeyError(f"checkpoint is missing {len(missing)} keys, e.g. {missing[:3]}")
        train, self.frozen = {}, {}
        for k in spec:
            if k.startswith(ALIAS):
                continue
            t = state_dict[k].detach()
            if _is_frozen(k, freeze_at):
                self.frozen[k] = t.to(self.device, torch.float32)
            else:
                train[k] = t
        self.flat = FlatParams(train, self.device)
        self._plist = nn.ParameterList(list(self.flat.params.values()))     # module bookkeeping (.parameters())
        self.P = dict(self.frozen)
        self.P.update(self.flat.params)
        self.dropout, self.M, self.L, self.K = dropout, heads, levels, points
        self.num_classes, self.mask_stride = num_classes, mask_stride
        self.mean = torch.tensor(pixel_mean, device=self.device).view(1, 3, 1, 1)
        self.std = torch.tensor(pixel_std, device=self.device).view(1, 3, 1, 1)
        self.weight_dict = weight_dict or crit.default_weight_dict()
        self.criterion = crit.SetCriterion(num_classes, self.weight_dict, mask_out_stride=mask_stride)
        self.batched = batched                   # device-batched assignment + losses (criterion_batched.py)
        self.criterion_b = critb.BatchedCriterion(num_classes, mask_out_stride=mask_stride)
        self.enc_layers = 1 + max(int(k.split(".")[5]) for k in spec if k.startswith(TR + "encoder.layers."))
        self.dec_layers = 1 + max(int(k.split(".")[5]) for k in spec if k.startswith(TR + "decoder.layers."))

    @staticmethod
    def _state_spec(num_queries, num_classes, depth=50):
        return idol_r50_state_spec(num_queries, num_classes, depth=depth)

    # ------------------------------------------------------------------ state dict in the reference's layout
    def reference_state_dict(self):
        out = {k: v.detach().clone() for k, v in self.P.items()}
        for k in list(out):
            if k.startswith(DETR + "bbox_embed."):
                out[ALIAS + k[len(DETR + "bbox_embed."):]] = out[k]
        return out

    # ------------------------------------------------------------------ small helpers
    def _drop(self, x):
        return F.dropout(x, self.dropout, self.training) if self.dropout > 0 else x

    def _linear(self, x, w, b, relu=False):
        if self.tc:
            return TcLinear.apply(x, w, b, relu)
        y = F.linear(x, w, b)
        return F.relu(y) if relu else y

    def _lin(self, p, x, relu=False):
        return self._linear(x, self.P[p + ".weight"], self.P[p + ".bias"], relu)

    def _ln(self, p, x):
        return F.layer_norm(x, (x.shape[-1],), self.P[p + ".weight"], self.P[p + ".bias"], 1e-5)

    def _mlp(self, p, x, n=3):
        for i in range(n):
            x = self._lin(f"{p}.layers.{i}", x, relu=i < n - 1)
        return x

    def _conv2d(self, x, w, b, stride=1, padding=0):
        if self.cl:
            w = w.contiguous(memory_format=torch.channels_last)
        return F.conv2d(x, w, b, stride=stride, padding=padding)

    def _conv_bn(self, p, x, stride=1, padding=0, relu=True):
        n = p + ".norm."
        scale = self.P[n + "weight"] * torch.rsqrt(self.P[n + "running_var"] + 1e-5)      # batch_norm.py:48-53
        shift = self.P[n + "bias"] - self.P[n + "running_mean"] * scale
        x = self._conv2d(x, self.P[p + ".weight"] * scale.view(-1, 1, 1, 1), shift, stride=stride, padding=padding)
        return F.relu(x, inplace=True) if relu else x

    # ------------------------------------------------------------------ backbone (resnet.py:194-210, 355-359, 435-458)
    def backbone(self, x):
        out = {}
        with torch.no_grad():
            x = self._conv_bn(BB + "stem.conv1", x, stride=2, padding=3)
            x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        for stage in ("res2", "res3", "res4", "res5"):       # block counts from the checkpoint's keys (R50 / R101 / R152)
            nblocks = 0
            while f"{BB}{stage}.{nblocks}.conv1.weight" in self.P:
                nblocks += 1
            for b in range(nblocks):
                p = f"{BB}{stage}.{b}."
                stride = 2 if (b == 0 and stage != "res2") else 1
                y = self._conv_bn(p + "conv1", x)
                y = self._conv_bn(p + "conv2", y, stride=stride, padding=1)
                y = self._conv_bn(p + "conv3", y, relu=False)
                sc = self._conv_bn(p + "shortcut", x, stride=stride, relu=False) if (p + "shortcut.weight") in self.P \
                    else x
                x = F.relu(y + sc)
            out[stage] = x
        return out

    @staticmethod
    def _sine_pos(mask, num_pos_feats=128, temperature=10000.0):
        """position_encoding.py:36-56 (normalize=True)."""
        nm = ~mask
        y, x = nm.cumsum(1, dtype=torch.float32), nm.cumsum(2, dtype=torch.float32)
        y = (y - 0.5) / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
        x = (x - 0.5) / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
        d = torch.arange(num_pos_feats, dtype=torch.float32, device=mask.device)
        d = temperature ** (2 * (d // 2) / num_pos_feats)
        px, py = x[..., None] / d, y[..., None] / d
        px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), 4).flatten(3)
        py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), 4).flatten(3)
        return torch.cat((py, px), 3)                                           # [N, h, w, 256] (token-major)

    # ------------------------------------------------------------------ MSDeformAttn module (ops/modules/ms_deform_attn.py:78-116)
    def _msda(self, p, query, ref, src, shapes, lsi, pad_mask):
        N, Lq, C = query.shape
        M, L, K = self.M, self.L, self.K
        v = self._lin(p + ".value_proj", src)
        if pad_mask is not None:
            v = v.masked_fill(pad_mask[..., None], 0.0)
        n_off = M * L * K * 2
        w = torch.cat([self.P[p + ".sampling_offsets.weight"], self.P[p + ".attention_weights.weight"]], 0)
        b = torch.cat([self.P[p + ".sampling_offsets.bias"], self.P[p + ".attention_weights.bias"]], 0)
        oa = self._linear(query, w, b)                                          # one pass over the queries for both
        off = oa[..., :n_off].reshape(N, Lq, M, L, K, 2)
        aw = torch.softmax(oa[..., n_off:].reshape(N, Lq, M, L * K), -1).view(N, Lq, M, L, K)
        if ref.shape[-1] == 2:
            norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1).to(off.dtype)
            loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        else:
            loc = ref[:, :, None, :, None, :2] + off / K * ref[:, :, None, :, None, 2:] * 0.5
        out = MSDeformAttnFunction.apply(v.view(N, -1, M, C // M), shapes, lsi, loc.contiguous(), aw.contiguous(), 64)
        return self._lin(p + ".output_proj", out)

    def _mha(self, p, q_in, v_in):
        """nn.MultiheadAttention(q=k=q_in, v=v_in) of the decoder self-attention (deformable_transformer.py:306-313)."""
        C = q_in.shape[-1]
        w, b = self.P[p + ".in_proj_weight"], self.P[p + ".in_proj_bias"]
        qk = self._linear(q_in, w[:2 * C], b[:2 * C])
        v = self._linear(v_in, w[2 * C:], b[2 * C:])
        B, Lq, _ = qk.shape
        d = C // self.M
        q = qk[..., :C].reshape(B, Lq, self.M, d).transpose(1, 2)
        k = qk[..., C:].reshape(B, Lq, self.M, d).transpose(1, 2)
        v = v.reshape(B, Lq, self.M, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=self.dropout if self.training else 0.0)
        return self._linear(o.transpose(1, 2).reshape(B, Lq, C), self.P[p + ".out_proj.weight"], self.P[p + ".out_proj.bias"])

    # ------------------------------------------------------------------ dense forward over all frames
    def front(self, frames, level3_from_full_mask=False, frames_per_video=1):
        """Normalise + pad + backbone + input_proj + token flattening + encoder geometry.  frames: [N,3,H,W] tensor or
        list of N [3,h_i,w_i] tensors, float/uint8 0..255; ragged sizes are zero-padded after normalisation like
        ImageList.from_tensors(…, 32) (idol.py:487-497).  `frames_per_video` > 1: valid ratios come from each video's
        first frame (SeqFormer, deformable_transformer.py:107); `level3_from_full_mask`: the stride-64 level's padding
        mask is resized from the full-resolution mask (SeqFormer segm:117) instead of from level 0's (IDOL segm:117-118).
        -> dict(src, pos, pad [N,S,(C)], shapes, lsi, vr [N,L,2], enc_ref [N,S,L,2], sizes)"""
        if torch.is_tensor(frames):
            frames = list(frames)
        N = len(frames)
        sizes = [tuple(f.shape[-2:]) for f in frames]
        Hp = (max(h for h, _ in sizes) + 31) // 32 * 32
        Wp = (max(w for _, w in sizes) + 31) // 32 * 32
        x = torch.zeros(N, 3, Hp, Wp, device=self.device)
        mask = torch.ones(N, Hp, Wp, dtype=torch.bool, device=self.device)
        for i, f in enumerate(frames):
            h, w = sizes[i]
            x[i, :, :h, :w] = (f.to(self.device, torch.float32) - self.mean[0]) / self.std[0]
            mask[i, :h, :w] = False
        if self.cl:
            x = x.contiguous(memory_format=torch.channels_last)
        feats = self.backbone(x)
        srcs, masks, poses = [], [], []
        for l, name in enumerate(("res3", "res4", "res5", "res5")):
            k = 3 if l == 3 else 1
            s = self._conv2d(feats[name], self.P[f"{DETR}input_proj.{l}.0.weight"], self.P[f"{DETR}input_proj.{l}.0.bias"],
                         stride=2 if l == 3 else 1, padding=k // 2)
            s = F.group_norm(s, 32, self.P[f"{DETR}input_proj.{l}.1.weight"], self.P[f"{DETR}input_proj.{l}.1.bias"], 1e-5)
            with torch.no_grad():
                # level masks: nearest-resized padding mask; level 3 resizes masks[0] (segm:117-118)
                base = mask if (l < 3 or level3_from_full_mask) else masks[0]
                m = F.interpolate(base[None].float(), size=s.shape[-2:]).to(torch.bool)[0]
                pos = self._sine_pos(m).flatten(1, 2)
            srcs.append(s.permute(0, 2, 3, 1).flatten(1, 2))                       # a view when s is channels-last
            masks.append(m)
            poses.append(pos + self.P[TR + "level_embed"][l].view(1, 1, -1))
        level_hw = [tuple(m.shape[-2:]) for m in masks]
        shapes = torch.as_tensor(level_hw, dtype=torch.long, device=self.device)
        lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
        from ..msda import attach_host_geometry
        attach_host_geometry(shapes, lsi, level_hw)          # the self-attention MSDA kernel plans its tiles on the host
        src = torch.cat(srcs, 1)
        pos = torch.cat(poses, 1)
        pad = torch.cat([m.flatten(1) for m in masks], 1)
        with torch.no_grad():
            vr = torch.stack([torch.stack([(~m[:, 0, :]).sum(1).float() / m.shape[2],
                                           (~m[:, :, 0]).sum(1).float() / m.shape[1]], -1) for m in masks], 1)   # [N,L,2]
            if frames_per_video > 1:
                vr = vr[0::frames_per_video].repeat_interleave(frames_per_video, 0)
            refs = []
            for l, (h, w) in enumerate(level_hw):
                ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=self.device),
                                        torch.linspace(0.5, w - 0.5, w, device=self.device), indexing="ij")
                refs.append(torch.stack((rx.reshape(-1)[None] / (vr[:, None, l, 0] * w),
                                         ry.reshape(-1)[None] / (vr[:, None, l, 1] * h)), -1))
            enc_ref = torch.cat(refs, 1)[:, :, None] * vr[:, None]                               # [N,S,L,2]
        return {"src": src, "pos": pos, "pad": pad, "shapes": shapes, "lsi": lsi, "vr": vr, "enc_ref": enc_ref,
                "sizes": sizes}

    def encode(self, fr):
        out, pos, shapes, lsi, pad = fr["src"], fr["pos"], fr["shapes"], fr["lsi"], fr["pad"]
        for i in range(self.enc_layers):
            p = f"{TR}encoder.layers.{i}"
            a = self._msda(p + ".self_attn", out + pos, fr["enc_ref"], out, shapes, lsi, pad)
            out = self._ln(p + ".norm1", out + self._drop(a))
            f = self._lin(p + ".linear2", self._drop(self._lin(p + ".linear1", out, relu=True)))
            out = self._ln(p + ".norm2", out + self._drop(f))
        return out

    def dense(self, frames):
        """frames (N = 2 x videos: key, ref, key, ref ...) -> dict(hs, coords, level_ref, last_ref, memory [N,S,C],
        shapes, lsi, sizes [(h,w)]*N)"""
        fr = self.front(frames)
        memory = self.encode(fr)
        N = memory.shape[0]
        shapes, lsi, pad, vr, sizes = fr["shapes"], fr["lsi"], fr["pad"], fr["vr"], fr["sizes"]

        qe = self.P[DETR + "query_embed.weight"]
        C = memory.shape[-1]
        qpos, tgt = qe[:, :C].unsqueeze(0).expand(N, -1, -1), qe[:, C:].unsqueeze(0).expand(N, -1, -1)
        ref = self._lin(TR + "reference_points", qpos).sigmoid()
        init_ref = ref
        hs, coords, inter = [], [], []
        o = tgt
        for lid in range(self.dec_layers):
            p = f"{TR}decoder.layers.{lid}"
            ref_in = ref[:, :, None] * (torch.cat([vr, vr], -1) if ref.shape[-1] == 4 else vr)[:, None]
            q = o + qpos
            o = self._ln(p + ".norm2", o + self._drop(self._mha(p + ".self_attn", q, o)))
            a = self._msda(p + ".cross_attn", o + qpos, ref_in, memory, shapes, lsi, pad)
            o = self._ln(p + ".norm1", o + self._drop(a))
            f = self._lin(p + ".linear2", self._drop(self._lin(p + ".linear1", o, relu=True)))
            o = self._ln(p + ".norm3", o + self._drop(f))
            # box refinement (deformable_transformer.py:363-374); the same expression is the level's predicted box
            # (segm:154-165), so it is computed once: with grad for the head, detached for the next layer.
            tmp = self._mlp(f"{DETR}bbox_embed.{lid}", o)
            rl = inverse_sigmoid(ref)
            new = (tmp + rl if ref.shape[-1] == 4 else torch.cat([tmp[..., :2] + rl, tmp[..., 2:]], -1)).sigmoid()
            hs.append(o)
            coords.append(new)
            inter.append(ref)                 # the reference this level's head/mask branch uses (init or previous)
            ref = new.detach()
        return {"hs": hs, "coords": coords, "level_ref": inter, "last_ref": ref, "memory": memory, "shapes": shapes,
                "lsi": lsi, "sizes": sizes, "init_ref": init_ref}

    # ------------------------------------------------------------------ mask branch + dynamic head (segm:324-374, 376-470, 539-575)
    def _conv(self, p, x):
        return self._conv2d(x, self.P[p + ".weight"], self.P[p + ".bias"], padding=1)

    def mask_features(self, memory, shapes, lsi):
        lv = []
        for l in range(self.L - 1):
            h, w = shapes[l].tolist()
            s0 = int(lsi[l])
            lv.append(memory[:, s0:s0 + h * w].reshape(memory.shape[0], h, w, -1).permute(0, 3, 1, 2))
        p = "detr.mask_head."
        x = F.relu(self._conv(p + "lay3", lv[-1]))
        x = lv[-2] + F.interpolate(x, size=lv[-2].shape[-2:], mode="nearest")
        x = F.relu(self._conv(p + "lay4", x))
        x = lv[-3] + F.interpolate(x, size=lv[-3].shape[-2:], mode="nearest")
        x = F.relu(self._conv(p + "dcn", x))
        x = F.relu(self._conv(p + "lay1", x))
        return F.relu(self._conv(p + "lay2", x))                                   # [B, 8, H/8, W/8]

    @staticmethod
    def _aligned_bilinear(t, factor):
        if factor == 1:
            return t
        h, w = t.shape[2:]
        t = F.pad(t, pad=(0, 1, 0, 1), mode="replicate")
        oh, ow = factor * h + 1, factor * w + 1
        t = F.interpolate(t, size=(oh, ow), mode="bilinear", align_corners=True)
        t = F.pad(t, pad=(factor // 2, 0, factor // 2, 0), mode="replicate")
        return t[:, :, :oh - 1, :ow - 1]

    def instance_masks(self, feat, ref_px, params, stride=8, ch=8):
        """feat [8,h,w] of one image; ref_px [n,2]; params [n,169] -> logits [n, 2h, 2w]."""
        cin, h, w = feat.shape
        n = ref_px.shape[0]
        if n == 0:
            f = stride // self.mask_stride
            return (feat.sum() * 0 + params.sum() * 0).expand(0, h * f, w * f)
        xs = torch.arange(0, w * stride, stride, dtype=torch.float32, device=feat.device) + stride // 2
        ys = torch.arange(0, h * stride, stride, dtype=torch.float32, device=feat.device) + stride // 2
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        loc = torch.stack((xx.reshape(-1), yy.reshape(-1)), 0)                     # [2, hw]
        rel = ref_px[:, :, None] - loc[None]                                       # [n, 2, hw]
        x = torch.cat([rel, feat.reshape(1, cin, h * w).expand(n, -1, -1)], 1)     # [n, 10, hw]
        sizes = [(cin + 2) * ch, ch * ch, ch, ch, ch, 1]
        w0, w1, w2, b0, b1, b2 = torch.split_with_sizes(params, sizes, dim=1)
        x = F.relu(torch.baddbmm(b0[:, :, None], w0.reshape(n, ch, cin + 2), x))
        x = F.relu(torch.baddbmm(b1[:, :, None], w1.reshape(n, ch, ch), x))
        x = torch.baddbmm(b2[:, :, None], w2.reshape(n, 1, ch), x)
        x = self._aligned_bilinear(x.reshape(n, 1, h, w), stride // self.mask_stride)
        return x[:, 0]

    def instance_masks_batched(self, feat, image_of, ref_px, params, stride=8, ch=8):
        """feat [B,8,h,w]; image_of [n] long; ref_px [n,2]; params [n,169] -> logits [n, 2h, 2w] (all levels at once)."""
        B, cin, h, w = feat.shape
        n = ref_px.shape[0]
        xs = torch.arange(0, w * stride, stride, dtype=torch.float32, device=feat.device) + stride // 2
        ys = torch.arange(0, h * stride, stride, dtype=torch.float32, device=feat.device) + stride // 2
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        loc = torch.stack((xx.reshape(-1), yy.reshape(-1)), 0)
        x = torch.cat([ref_px[:, :, None] - loc[None], feat.flatten(2)[image_of]], 1)             # [n, 10, hw]
        sizes = [(cin + 2) * ch, ch * ch, ch, ch, ch, 1]
        w0, w1, w2, b0, b1, b2 = torch.split_with_sizes(params, sizes, dim=1)
        x = F.relu(torch.baddbmm(b0[:, :, None], w0.reshape(n, ch, cin + 2), x))
        x = F.relu(torch.baddbmm(b1[:, :, None], w1.reshape(n, ch, ch), x))
        x = torch.baddbmm(b2[:, :, None], w2.reshape(n, 1, ch), x)
        return self._aligned_bilinear(x.reshape(n, 1, h, w), stride // self.mask_stride)[:, 0]

    # ------------------------------------------------------------------ the training forward
    def forward(self, frames, det_targets, ref_targets):
        """-> (loss_dict unweighted, outputs).  frames [2B,3,H,W] or a list of 2B [3,h,w]; targets as
        IDOL.prepare_targets builds them."""
        if self.batched:
            return self.forward_batched(frames, det_targets, ref_targets)
        return self.forward_sequential(frames, det_targets, ref_targets)

    def forward_batched(self, frames, det_targets, ref_targets):
        """All decoder levels' heads, assignments and losses as stacked tensors (criterion_batched.py)."""
        d = self.dense(frames)
        key, rf = slice(0, None, 2), slice(1, None, 2)
        L = self.dec_layers
        hs = torch.stack(d["hs"])                                                   # [L, N, nq, C]
        hk = hs[:, key]
        B, nq = hk.shape[1:3]
        wc = torch.stack([self.P[f"{DETR}class_embed.{i}.weight"] for i in range(L)])
        bc = torch.stack([self.P[f"{DETR}class_embed.{i}.bias"] for i in range(L)])
        logits = torch.baddbmm(bc[:, None], hk.flatten(1, 2), wc.transpose(1, 2)).view(L, B, nq, -1)
        boxes = torch.stack([c[key] for c in d["coords"]])
        losses, assign = self.criterion_b.detection_losses(logits, boxes, det_targets)
        # dynamic mask head for every selected (level, image, query) at once
        sel = assign["sel"]
        idx = torch.nonzero(sel)                                                    # the one sync listing the instances
        l_i, b_i, q_i = idx.unbind(1)
        scale = torch.tensor([[float(w), float(h)] for h, w in d["sizes"][key]], device=self.device)[:, None]
        ref_px = torch.stack([inverse_sigmoid(r[key]).sigmoid()[..., :2] for r in d["level_ref"]]) * scale   # [L,B,nq,2]
        feat = self.mask_features(d["memory"][key], d["shapes"], d["lsi"])
        tm = self.criterion_b.mask_targets(det_targets, ref_targets, self.device)
        if idx.shape[0] > 0:
            params = self._mlp("detr.controller", hk[l_i, b_i, q_i])
            ml = self.instance_masks_batched(feat, b_i, ref_px[l_i, b_i, q_i], params)
            losses.update(self.criterion_b.mask_losses(ml, tm[b_i, assign["gt_of"][l_i, b_i, q_i]], l_i, L))
        else:
            z = (feat.sum() * 0).expand(L)
            losses.update({"loss_mask": z, "loss_dice": z})
        # contrastive reID terms (last level)
        hs_key, hs_ref = hs[-1, key], hs[-1, rf]
        ref_prob = self._lin(f"{DETR}class_embed.{L - 1}", hs_ref).sigmoid()
        ke, re = self._mlp("detr.reid_embed_head", hs_key), self._mlp("detr.reid_embed_head", hs_ref)
        reid = self.criterion_b.reid_losses(d["last_ref"][rf], ref_prob.detach(), ke, re, assign["best_q"][-1], ref_targets)
        if reid is None:                                                            # repair-loop cases: sequential code
            matched = [assign["best_q"][-1][b, :len(t["labels"])] for b, t in enumerate(det_targets)]
            items = crit.select_pos_neg(d["last_ref"][rf], matched, ref_targets, det_targets, ke, re, ref_prob)
            r = self.criterion.loss_reid({"pred_qd": items, "pred_logits": logits})
            reid = (r["loss_reid"], r["loss_reid_aux"], len(items))
        out = {}
        for k, v in losses.items():
            out[k] = v[L - 1]
            for i in range(L - 1):
                out[f"{k}_{i}"] = v[i]
        out["loss_reid"], out["loss_reid_aux"] = reid[0], reid[1]
        return out, {"pred_logits": logits[-1], "pred_boxes": boxes[-1], "assign": assign, "num_contrast_items": reid[2]}

    def forward_sequential(self, frames, det_targets, ref_targets):
        """The reference's control flow level by level (criterion.py) — the parity anchor for forward_batched."""
        d = self.dense(frames)
        key, rf = slice(0, None, 2), slice(1, None, 2)
        scale = torch.tensor([[float(w), float(h)] for h, w in d["sizes"][key]], device=self.device)[:, None]  # segm:177-181
        feat = self.mask_features(d["memory"][key], d["shapes"], d["lsi"])
        levels, indices_list, matched = [], [], None
        for lvl in range(self.dec_layers):
            hk = d["hs"][lvl][key]
            out = {"pred_logits": self._lin(f"{DETR}class_embed.{lvl}", hk), "pred_boxes": d["coords"][lvl][key]}
            params = self._mlp("detr.controller", hk)
            indices, matched = self.criterion.matcher(out, det_targets)
            ref_px = inverse_sigmoid(d["level_ref"][lvl][key]).sigmoid()[..., :2] * scale
            out["pred_masks"] = [self.instance_masks(feat[b], ref_px[b][sel], params[b][sel])[None, :, None]
                                 for b, (sel, _) in enumerate(indices)]
            levels.append(out)
            indices_list.append(indices)
        outputs = dict(levels[-1])
        hs_key, hs_ref = d["hs"][-1][key], d["hs"][-1][rf]
        ref_cls = self._lin(f"{DETR}class_embed.{self.dec_layers - 1}", hs_ref).sigmoid()
        outputs["pred_qd"] = crit.select_pos_neg(d["last_ref"][rf], matched, ref_targets, det_targets,
                                                 self._mlp("detr.reid_embed_head", hs_key),
                                                 self._mlp("detr.reid_embed_head", hs_ref), ref_cls)
        outputs["aux_outputs"] = levels[:-1]
        return self.criterion(outputs, det_targets, ref_targets, indices_list), outputs

    def weighted(self, loss_dict):
        return {k: v * self.weight_dict[k] for k, v in loss_dict.items() if k in self.weight_dict}


class IDOLTrainer:
    """One optimisation step = zero flat grads, forward, backward, ONE all-reduce of the flat gradient buffer,
    full-model gradient-norm clipping (train_net.py:110-125, CLIP_VALUE 0.01) and AdamW (train_net.py:132-135:
    BASE_LR 1e-4, backbone x0.1, WEIGHT_DECAY 1e-4 — configs/ytvis19_r50.yaml:19-30)."""

    def __init__(self, model, base_lr=1e-4, backbone_multiplier=0.1, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 clip_value=0.01, process_group=None):
        self.model, self.flat = model, model.flat
        self.lrs = (base_lr * backbone_multiplier, base_lr)
        self.wd, self.betas, self.eps, self.clip = weight_decay, betas, eps, clip_value
        self.step_count = 0
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
            model.criterion.all_reduce_fn = lambda t: torch.distributed.all_reduce(t, group=process_group)
        self.scale = torch.ones(1, device=self.flat.w.device)

    def step(self, frames, det_targets, ref_targets):
        self.flat.g.zero_()
        loss_dict, _ = self.model(frames, det_targets, ref_targets)
        weighted = self.model.weighted(loss_dict)
        total = sum(weighted.values())
        total.backward()
        return self.apply_gradients(), total.detach(), weighted

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Everything a resume needs: weights in the REFERENCE key layout (what DetectionCheckpointer writes under
        "model"), the flat Adam moments with the layout they refer to, and the step count."""
        return {"model": self.model.reference_state_dict(), "adam_m": self.flat.m.detach().cpu().clone(),
                "adam_v": self.flat.v.detach().cpu().clone(), "step": self.step_count,
                "layout": {k: v[:2] for k, v in self.flat.offsets.items()}}

    def load_state_dict(self, state):
        layout = {k: v[:2] for k, v in self.flat.offsets.items()}
        if state["layout"] != layout:
            raise ValueError("optimizer state was saved for a different parameter layout")
        sd = state["model"]
        with torch.no_grad():
            for k, p in self.flat.params.items():
                p.copy_(sd[k].to(p.device, p.dtype))
            for k, t in self.model.frozen.items():
                t.copy_(sd[k].to(t.device, t.dtype))
            self.flat.m.copy_(state["adam_m"])
            self.flat.v.copy_(state["adam_v"])
        self.step_count = int(state["step"])

    def reduce_gradients(self):
        """Sum the flat gradient buffer over ranks (one collective) and set the device-side scale that turns it into
        the clipped mean gradient.  -> L2 norm of the mean gradient (a device scalar; nothing is synchronised)."""
        g = self.flat.g
        if self.world > 1:
            torch.distributed.all_reduce(g, group=self.pg)                       # sum; the 1/world is folded below
        norm = torch.linalg.vector_norm(g) / self.world
        if self.clip and self.clip > 0:
            self.scale = (self.clip / (norm + 1e-6)).clamp(max=1.0).reshape(1) / self.world   # clip_grad_norm_
        else:
            self.scale = torch.full((1,), 1.0 / self.world, device=g.device)
        return norm

    def apply_gradients(self):
        norm = self.reduce_gradients()
        self.step_count += 1
        f = self.flat
        for (lo, hi), lr in zip(f.groups(), self.lrs):
            adamw_(f.w[lo:hi], f.g[lo:hi], f.m[lo:hi], f.v[lo:hi], lr, self.betas[0], self.betas[1], self.eps, self.wd,
                   self.step_count, self.scale)
        return norm
