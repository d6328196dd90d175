"""IDOL per-frame inference engine: `CondInst_segm.inference_forward` (projects/IDOL/idol/models/
segmentation_condInst.py:233-321) on the B200 kernels.  The frames of a video (or of a BATCH_INFER_LEN chunk,
idol.py:252-274) are the batch; backbone, encoder, mask branch and all kernels are shared with the SeqFormer engine,
the decoder is the single-stream Deformable-DETR one (models/deformable_transformer.py:272-384) and the heads add
IDOL's reID embedding MLP (segmentation_condInst.py:75, :295).  The online tracker that consumes these outputs is
host-side post-processing and out of scope."""
import torch
import torch.nn.functional as F

from ..seqformer import packing
from ..seqformer.engine import DETR, TR, SeqFormerEngine, inverse_sigmoid


class IDOLEngine(SeqFormerEngine):
    def _pack_decoder(self, sd, lin, f32, dec_layers):
        dev = self.device
        self.dec = []
        for i in range(dec_layers):
            p = f"{TR}decoder.layers.{i}"
            d = {"attn": packing.to_device(packing.pack_msda(sd, p + ".cross_attn", box=False), dev)}
            for k in ("norm1", "norm2", "norm3", "linear1", "linear2"):
                d[k] = lin(f"{p}.{k}")
            d["self_attn"] = (f32(sd[f"{p}.self_attn.in_proj_weight"]), f32(sd[f"{p}.self_attn.in_proj_bias"]),
                              f32(sd[f"{p}.self_attn.out_proj.weight"]), f32(sd[f"{p}.self_attn.out_proj.bias"]))
            d["bbox"] = [lin(f"{DETR}bbox_embed.{i}.layers.{j}") for j in range(3)]
            d["cls"] = lin(f"{DETR}class_embed.{i}")
            self.dec.append(d)
        self.reid = [lin(f"detr.reid_embed_head.layers.{j}") for j in range(3)]

    def _mlp(self, x, layers):
        for j, wb in enumerate(layers):
            x = self.lin(x, wb, relu=(j < len(layers) - 1))
        return x

    def decoder(self, memory, plan, collect_all=False):
        N = plan.T
        qpos = self.query_pos.expand(N, -1, -1)
        out = self.tgt.expand(N, -1, -1)
        ref = self._lin(self.query_pos, self.ref_point).sigmoid().expand(N, -1, -1)           # [N, nq, 2]
        vr = plan.valid_ratios
        prev_ref = ref
        for lid, w in enumerate(self.dec):
            ref_in = ref[:, :, None] * (torch.cat([vr, vr], -1) if ref.shape[-1] == 4 else vr)[:, None]
            a = w["attn"]
            if self.native_gemm and out.dtype == torch.float32:
                # merged q | k | v projection with the query-position term as a 2-D bias + one attention kernel (see
                # SeqFormerEngine._mha_self); same for the offsets | logits projection
                out = self._mha_self(out.contiguous(), w["self_attn"], (lid, "sa"), w["norm2"])
                oa = self.lin(out, (a["oa_w"], None), bias2d=self._qpos_bias((lid, "oa"), a["oa_w"], a["oa_b"]))
            else:
                q = out + qpos
                out = self._mha(q, out, w["self_attn"], residual=out, ln=w["norm2"])
                oa = self.lin(out + qpos, (a["oa_w"], a["oa_b"]))
            s = self.sample(self._value(memory, a, plan), oa, ref_in, plan, a["n_off"]).float()
            out = self.lin(s, (a["o_w"], a["o_b"]), residual=out, ln=w["norm1"])
            out = self.lin(self.lin(out, w["linear1"], relu=True), w["linear2"], residual=out, ln=w["norm3"])
            tmp = self._mlp(out, w["bbox"])
            prev_ref = ref
            if self.native_gemm and tmp.dtype == torch.float32:
                from .. import fused
                ref = fused.box_refine(tmp, ref)
            elif ref.shape[-1] == 4:
                ref = (tmp + inverse_sigmoid(ref)).sigmoid()
            else:
                ref = torch.cat([tmp[..., :2] + inverse_sigmoid(ref), tmp[..., 2:]], -1).sigmoid()
        return out, prev_ref                     # hs[-1], inter_references[-2]

    @torch.no_grad()
    def inference_forward(self, frames):
        """frames [N,3,H,W] uint8/float (0..255) on the device -> dict like the reference's inference_forward:
        pred_logits [N,nq,K], pred_boxes [N,nq,4], pred_inst_embed [N,nq,C], pred_masks [N,nq,1,H/4,W/4]."""
        self._flags()
        N, _, H, W = frames.shape
        plan = self.plan(N, H, W)
        memory = self.encoder(self.tokens(self.backbone(self.preprocess(frames, plan)), plan), plan)
        hs, ref_prev = self.decoder(memory, plan)
        w = self.dec[-1]
        reference = inverse_sigmoid(ref_prev)
        logits = self.lin(hs, w["cls"])
        tmp = self._mlp(hs, w["bbox"])
        tmp = tmp + reference if reference.shape[-1] == 4 else torch.cat([tmp[..., :2] + reference, tmp[..., 2:]], -1)
        boxes = tmp.sigmoid()
        embed = self._mlp(hs, self.reid)
        params = self._mlp(hs, self.controller)                                     # [N, nq, 169]
        ref_px = ref_prev[..., :2] * plan.extra["scale_wh"]
        feats = self.mask_features(memory, plan)
        masks = self.dynamic_masks(feats, params, ref_px)                           # [N, nq, H/4, W/4]
        return {"pred_logits": logits, "pred_boxes": boxes, "pred_inst_embed": embed, "pred_masks": masks.unsqueeze(2),
                "reference_points": ref_prev[..., :2], "memory": memory}
