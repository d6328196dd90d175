"""`MSDeformAttn` nn.Module drop-ins on the B200 kernels.

Same constructor arguments, parameter names (state-dict keys), initialisation and return values as
  * IDOL / Deformable-DETR:  projects/IDOL/idol/models/ops/modules/ms_deform_attn.py:30-116
  * SeqFormer (per-frame, dual output projection):  projects/SeqFormer/seqformer/models/ops/modules/ms_deform_attn.py:32-217
but: the T frames of SeqFormer's variant are folded into the batch of ONE op call (the reference loops over frames),
and under `torch.no_grad()` the softmax + sampling-location arithmetic + sampling run as one fused kernel
(vnext_msda_forward_fused).  With gradients enabled the classic autograd op (MSDeformAttnFunction) is used, so the
modules train exactly like the reference's.
"""
import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from .. import msda


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class _MSDeformAttnBase(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head a "
                          "power of 2 which is more efficient in our CUDA implementation.")
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)

    def _reset_parameters(self):
        constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2)
        grid_init = grid_init.repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    # ---- shared pieces
    def _value(self, input_flatten, input_padding_mask):
        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        return value.view(*value.shape[:-1], self.n_heads, self.d_model // self.n_heads)

    def _locations(self, reference_points, sampling_offsets, input_spatial_shapes):
        # reference_points [N, Lq, L, 2|4]; sampling_offsets [N, Lq, M, L, P, 2]
        if reference_points.shape[-1] == 2:
            normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            return reference_points[:, :, None, :, None, :] + sampling_offsets / normalizer[None, None, None, :, None, :]
        if reference_points.shape[-1] == 4:
            return reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead."
                         .format(reference_points.shape[-1]))

    def _core(self, query, reference_points, value, shapes, lsi, want_locations):
        """query [N, Lq, C]; reference_points [N, Lq, L, 2|4]; value [N, S, M, D] -> sampled [N, Lq, C]
        (+ sampling_locations, attention_weights when asked for or when autograd is on)."""
        N, Lq, _ = query.shape
        M, L, P = self.n_heads, self.n_levels, self.n_points
        fused_ok = (not torch.is_grad_enabled()) and not want_locations and value.is_cuda \
            and value.shape[-1] == 32 and (L * P) in (4, 8, 16, 32) and value.dtype in (torch.float32, torch.bfloat16)
        if fused_ok:
            w = torch.cat([self.sampling_offsets.weight, self.attention_weights.weight], 0)
            b = torch.cat([self.sampling_offsets.bias, self.attention_weights.bias], 0)
            oa = F.linear(query, w, b).float()
            out = msda.ms_deform_attn_fused(value.contiguous(), shapes, lsi, oa, reference_points.float().contiguous(),
                                            L, P)
            return out, None, None
        offsets = self.sampling_offsets(query).view(N, Lq, M, L, P, 2)
        attn = F.softmax(self.attention_weights(query).view(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
        loc = self._locations(reference_points, offsets, shapes)
        out = msda.MSDeformAttnFunction.apply(value.contiguous(), shapes, lsi, loc.contiguous(), attn.contiguous(),
                                              self.im2col_step)
        return out, loc, attn


class MSDeformAttn(_MSDeformAttnBase):
    """IDOL / Deformable-DETR variant: returns (output, sampling_locations, attention_weights)."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__(d_model, n_levels, n_heads, n_points)
        self._reset_parameters()

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        N, Len_in, _ = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
        value = self._value(input_flatten, input_padding_mask)
        out, loc, attn = self._core(query, reference_points, value, input_spatial_shapes, input_level_start_index,
                                    want_locations=True)
        return self.output_proj(out), loc, attn


class MSDeformAttnSeqFormer(_MSDeformAttnBase):
    """SeqFormer variant (`mode` = 'encode' | 'decode'); frames are folded into the batch of one op call."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4, mode="encode"):
        super().__init__(d_model, n_levels, n_heads, n_points)
        self.mode = mode
        self.output_proj_box = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def forward(self, query, query_box, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        if self.mode == "encode":
            return self.encode_forward(query, reference_points, input_flatten, input_spatial_shapes,
                                       input_level_start_index, input_padding_mask)
        return self.decode_forward(query, query_box, reference_points, input_flatten, input_spatial_shapes,
                                   input_level_start_index, input_padding_mask)

    def encode_forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                       input_padding_mask=None):
        N, nf, Len_q, C = query.shape
        _, _, Len_in, _ = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
        if reference_points.shape[-1] != 2:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead."
                             .format(reference_points.shape[-1]))
        value = self._value(input_flatten, input_padding_mask).flatten(0, 1)          # [N*nf, S, M, D]
        ref = reference_points[:, None].expand(N, nf, *reference_points.shape[1:]).flatten(0, 1)
        out, _, _ = self._core(query.flatten(0, 1), ref, value, input_spatial_shapes, input_level_start_index, False)
        return self.output_proj(out.view(N, nf, Len_q, C))

    def decode_forward(self, query, query_box, reference_points, input_flatten, input_spatial_shapes,
                       input_level_start_index, input_padding_mask=None):
        N, nf, Len_in, C = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
        value = self._value(input_flatten, input_padding_mask).flatten(0, 1)
        if query_box.dim() == 3:        # first decoder layer: one set of offsets / weights shared by all frames
            qb = query_box[:, None].expand(N, nf, *query_box.shape[1:])
        else:
            assert query_box.dim() == 4
            qb = query_box
        Len_q = qb.shape[2]
        out, loc, attn = self._core(qb.flatten(0, 1), reference_points.flatten(0, 1), value, input_spatial_shapes,
                                    input_level_start_index, want_locations=True)
        out = out.view(N, nf, Len_q, C)
        # the reference returns the sampling locations of the LAST frame of its loop, and the attention weights as it
        # computed them: shared [N, Lq, M, L, P] in the first layer, per frame [N, nf, Lq, M, L, P] afterwards
        loc = loc.view(N, nf, *loc.shape[1:])[:, -1]
        attn = attn.view(N, nf, *attn.shape[1:])
        if query_box.dim() == 3:
            attn = attn[:, 0]
        return self.output_proj(out), self.output_proj_box(out), loc, attn
