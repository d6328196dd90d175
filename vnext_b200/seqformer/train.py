"""SeqFormer training forward (the clip-level counterpart of vnext_b200/idol/train.py): `CondInst_segm.forward(samples,
gt_targets, criterion, train=True)` (projects/SeqFormer/seqformer/models/segmentation_condInst.py:69-204) on a
reference-layout state dict, differentiable, every multi-scale deformable attention — 6 encoder layers x nf frames and 6
decoder layers x nf frames, folded into the batch — on the sm_100a forward/backward kernels.  Backbone, encoder, flat
parameter / gradient buffers, AdamW kernel, TcLinear and the trainer are shared with the IDOL path (same module tree
below the transformer decoder); what is SeqFormer's own is the dual-stream decoder (class/mask query shared by the
clip, box query per frame, temporal soft pooling: models/deformable_transformer.py:201-387 and
models/ops/modules/ms_deform_attn.py:128-217), the clip-level Hungarian matching and the per-frame mask head.

Structure choices as in the IDOL path: frames folded into the batch (the reference loops over frames in Python); the box
query's per-frame self-attention (deformable_transformer.py:288-297) is one batched attention; the mask branch is
evaluated once per step, not once per decoder level.
"""
import torch
import torch.nn.functional as F

from ..idol import train as IT
from ..idol.train import DETR, TR, IDOLTrainModel, IDOLTrainer, inverse_sigmoid  # noqa: F401  (re-exported)
from . import criterion as crit
from .state_spec import seqformer_r50_state_spec


class SeqFormerTrainModel(IDOLTrainModel):
    def __init__(self, state_dict, num_frames, weight_dict=None, **kw):
        kw.setdefault("batched", False)
        super().__init__(state_dict, weight_dict=weight_dict or crit.default_weight_dict(), **kw)
        self.nf = num_frames
        self.criterion = crit.SetCriterion(self.num_classes, self.weight_dict, num_frames, mask_out_stride=self.mask_stride)

    @staticmethod
    def _state_spec(num_queries, num_classes, depth=50):
        return seqformer_r50_state_spec(num_queries, num_classes, depth=depth)

    # ------------------------------------------------------------------ decode-mode MSDeformAttn (ms_deform_attn.py:128-217)
    def _sample(self, p, query_box, ref, value, shapes, lsi):
        """query_box [N,Lq,C] (drives offsets + weights), ref [N,Lq,L,2|4], value [N,S,M,D] -> sampled [N,Lq,C]."""
        N, Lq, C = query_box.shape
        M, L, K = self.M, self.L, self.K
        n_off = M * L * K * 2
        w = torch.cat([self.P[p + ".sampling_offsets.weight"], self.P[p + ".attention_weights.weight"]], 0)
        b = torch.cat([self.P[p + ".sampling_offsets.bias"], self.P[p + ".attention_weights.bias"]], 0)
        oa = self._linear(query_box, w, b)
        off = oa[..., :n_off].reshape(N, Lq, M, L, K, 2)
        aw = torch.softmax(oa[..., n_off:].reshape(N, Lq, M, L * K), -1).view(N, Lq, M, L, K)
        if ref.shape[-1] == 2:
            norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1).to(off.dtype)
            loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        else:
            loc = ref[:, :, None, :, None, :2] + off / K * ref[:, :, None, :, None, 2:] * 0.5
        return IT.MSDeformAttnFunction.apply(value, shapes, lsi, loc.contiguous(), aw.contiguous(), 64)

    # ------------------------------------------------------------------ dense forward over B clips of nf frames
    def dense(self, frames):
        """frames: B*nf frames, video-major (seqformer.py:413-429).  -> dict(hs [dec][B,nq,C], coords [dec][B,nf,nq,4],
        level_ref, memory [B,nf,S,C], shapes, lsi, sizes, vr [B,L,2])"""
        nf = self.nf
        fr = self.front(frames, level3_from_full_mask=True, frames_per_video=nf)
        memory = self.encode(fr)                                                    # [B*nf, S, C]
        N, S, C = memory.shape
        B = N // nf
        shapes, lsi, pad = fr["shapes"], fr["lsi"], fr["pad"]
        vr = fr["vr"][0::nf]                                                        # [B, L, 2] (first frame's)
        qe = self.P[DETR + "query_embed.weight"]
        nq = qe.shape[0]
        qpos, tgt = qe[:, :C].unsqueeze(0).expand(B, -1, -1), qe[:, C:].unsqueeze(0).expand(B, -1, -1)
        ref = self._lin(TR + "reference_points", qpos).sigmoid().unsqueeze(1).repeat(1, nf, 1, 1)    # [B,nf,nq,2]
        qpos_f = qpos.unsqueeze(1).expand(B, nf, nq, C).reshape(N, nq, C)
        out, out_box = tgt, tgt                                                     # box query [B,nq,C] -> [B,nf,nq,C]
        hs, hs_box, coords, level_ref = [], [], [], []
        for lid in range(self.dec_layers):
            p = f"{TR}decoder.layers.{lid}"
            scale = torch.cat([vr, vr], -1) if ref.shape[-1] == 4 else vr
            ref_in = (ref[:, :, :, None] * scale[:, None, None]).reshape(N, nq, self.L, ref.shape[-1])
            out = self._ln(p + ".norm2", out + self._drop(self._mha(p + ".self_attn", out + qpos, out)))
            if out_box.dim() == 3:
                ob = self._ln(p + ".norm2_box", out_box + self._drop(self._mha(p + ".self_attn_box", out_box + qpos, out_box)))
                query_box = (ob + qpos).unsqueeze(1).expand(B, nf, nq, C).reshape(N, nq, C)
                ob = ob.unsqueeze(1)                                                # broadcasts over the frames below
            else:
                obf = out_box.reshape(N, nq, C)                                     # the per-frame loop as one batch
                obf = self._ln(p + ".norm2_box", obf + self._drop(self._mha(p + ".self_attn_box", obf + qpos_f, obf)))
                query_box = obf + qpos_f
                ob = obf.view(B, nf, nq, C)
            a = p + ".cross_attn"
            v = self._lin(a + ".value_proj", memory).masked_fill(pad[..., None], 0.0).view(N, S, self.M, C // self.M)
            res = self._sample(a, query_box, ref_in, v, shapes, lsi).view(B, nf, nq, C)
            t2 = self._lin(a + ".output_proj", res)
            ob = self._ln(p + ".norm1_box", ob + self._drop(self._lin(a + ".output_proj_box", res)))
            f = self._lin(p + ".linear2_box", self._drop(self._lin(p + ".linear1_box", ob, relu=True)))
            out_box = self._ln(p + ".norm3_box", ob + self._drop(f))
            tw = torch.softmax(self._lin(p + ".time_attention_weights", out_box), 1)
            out = self._ln(p + ".norm1", out + self._drop((t2 * tw).sum(1)))
            f = self._lin(p + ".linear2", self._drop(self._lin(p + ".linear1", out, relu=True)))
            out = self._ln(p + ".norm3", out + self._drop(f))
            tmp = self._mlp(f"{DETR}bbox_embed.{lid}", out_box)
            rl = inverse_sigmoid(ref)
            new = (tmp + rl if ref.shape[-1] == 4 else torch.cat([tmp[..., :2] + rl, tmp[..., 2:]], -1)).sigmoid()
            hs.append(out)
            hs_box.append(out_box)
            coords.append(new)
            level_ref.append(ref)
            ref = new.detach()
        return {"hs": hs, "hs_box": hs_box, "coords": coords, "level_ref": level_ref, "memory": memory.view(B, nf, S, C),
                "shapes": shapes, "lsi": lsi, "sizes": fr["sizes"], "vr": vr}

    # ------------------------------------------------------------------ the training forward
    def forward(self, frames, targets):
        """-> (loss_dict unweighted, outputs).  frames: B*nf frames video-major ([B*nf,3,H,W] or a list); targets:
        per video {"labels","boxes" [n,nf,4],"masks" [n,nf,H,W],"size"} as SeqFormer.prepare_targets builds them."""
        if self.batched:
            return self.forward_batched(frames, targets)
        return self.forward_sequential(frames, targets)

    def forward_batched(self, frames, targets):
        """All six decoder levels at once: one cost tensor and ONE device→host copy for the 6 x B Hungarian problems,
        every loss a dense / indexed reduction over [levels, clips, queries], every matched (level, clip, query) through
        the dynamic mask head in one batch per frame.  Same goldens as forward_sequential (tests/test_seqformer_train_
        host.py); opt-in (`batched=True`) until it has been timed on the hardware."""
        from scipy.optimize import linear_sum_assignment
        from ..idol.criterion import cxcywh_to_xyxy, paired_giou_loss, pairwise_giou, pad_masks
        from ..idol.criterion_batched import focal_sum
        d = self.dense(frames)
        nf, L, dev = self.nf, self.dec_layers, self.device
        hs = torch.stack(d["hs"])                                                       # [L,B,nq,C]
        B, nq = hs.shape[1:3]
        wc = torch.stack([self.P[f"{DETR}class_embed.{i}.weight"] for i in range(L)])
        bc = torch.stack([self.P[f"{DETR}class_embed.{i}.bias"] for i in range(L)])
        logits = torch.baddbmm(bc[:, None], hs.flatten(1, 2), wc.transpose(1, 2)).view(L, B, nq, -1)
        boxes = torch.stack(d["coords"])                                                # [L,B,nf,nq,4]
        K = logits.shape[-1]
        sizes = [len(t["labels"]) for t in targets]
        n_tot = sum(sizes)
        tgt_ids = torch.cat([t["labels"] for t in targets])
        tgt_box = torch.cat([t["boxes"] for t in targets]).reshape(n_tot, nf, 4)
        with torch.no_grad():                                                           # models/matcher.py:53-91, all levels
            prob = logits.flatten(0, 2).sigmoid()
            ob = boxes.permute(0, 1, 3, 2, 4).flatten(0, 2)                             # [L*B*nq, nf, 4]
            c_bbox = torch.cdist(ob.flatten(1, 2), tgt_box.flatten(1, 2))
            tb = torch.clip(tgt_box, min=1e-7, max=1)
            c_giou = 0
            for i in range(nf):
                c_giou = c_giou - pairwise_giou(cxcywh_to_xyxy(ob[:, i]), cxcywh_to_xyxy(tb[:, i]))
            c_giou = c_giou / nf
            neg = 0.75 * (prob ** 2.0) * (-(1 - prob + 1e-8).log())
            pos = 0.25 * ((1 - prob) ** 2.0) * (-(prob + 1e-8).log())
            cc, cb, cg = self.criterion.costs
            C = (cb * c_bbox + cc * (pos[:, tgt_ids] - neg[:, tgt_ids]) + cg * c_giou).view(L, B, nq, n_tot).cpu()
        lv, bi, qi, ti, gi = [], [], [], [], []
        off = 0
        indices_list = [[None] * B for _ in range(L)]
        for b, n in enumerate(sizes):
            for l in range(L):
                r, c = linear_sum_assignment(C[l, b][:, off:off + n])
                indices_list[l][b] = (torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(c, dtype=torch.int64))
                lv += [l] * len(r); bi += [b] * len(r); qi += r.tolist(); ti += c.tolist(); gi += (c + off).tolist()
            off += n
        lv, bi, qi, ti, gi = (torch.tensor(x, dtype=torch.long, device=dev) for x in (lv, bi, qi, ti, gi))
        nb = torch.as_tensor([float(n_tot)], device=dev)
        if self.criterion.all_reduce_fn is not None:
            self.criterion.all_reduce_fn(nb)
        num_boxes = torch.clamp(nb / self.criterion.world_size_fn(), min=1)[0]
        zero = torch.zeros(L, device=dev)
        # labels
        classes = torch.full((L, B, nq), K, dtype=torch.int64, device=dev)
        classes[lv, bi, qi] = tgt_ids[gi]
        onehot = F.one_hot(classes, K + 1)[..., :K].to(logits.dtype)
        losses = {"loss_ce": focal_sum(logits, onehot, self.criterion.focal_alpha).flatten(1).sum(1) / num_boxes}
        last = lv == L - 1
        sel = logits[L - 1][bi[last], qi[last]]
        want = tgt_ids[gi[last]]
        class_error = 100 - ((sel.argmax(-1) == want).float().sum() * (100.0 / max(int(last.sum()), 1)))
        # boxes
        src = boxes.permute(0, 1, 3, 2, 4)[lv, bi, qi]                                   # [n, nf, 4]
        tgt = tgt_box[gi]
        l1 = (src - tgt).abs().flatten(1).sum(1) / nf
        gl = paired_giou_loss(cxcywh_to_xyxy(src.flatten(0, 1)), cxcywh_to_xyxy(tgt.flatten(0, 1))).view(-1, nf).sum(1) / nf
        losses["loss_bbox"] = zero.index_add(0, lv, l1) / num_boxes
        losses["loss_giou"] = zero.index_add(0, lv, gl) / num_boxes
        # masks: one dynamic-head batch per frame over all matched (level, clip, query) triples
        params = self._mlp("detr.controller", hs[lv, bi, qi])
        scale = torch.stack([torch.stack([t["size"][1], t["size"][0]]).to(dev, torch.float32) for t in targets])   # (w, h)
        ref_px = torch.stack([inverse_sigmoid(r).sigmoid()[..., :2] for r in d["level_ref"]])                    # [L,B,nf,nq,2]
        ml = []
        for f in range(nf):
            feat = self.mask_features(d["memory"][:, f], d["shapes"], d["lsi"])
            ml.append(self.instance_masks_batched(feat, bi, ref_px[lv, bi, f, qi] * scale[bi], params))
        ml = torch.stack(ml, 1)                                                          # [n, nf, h, w]
        tm = pad_masks([t["masks"].flatten(0, 1) for t in targets]).to(ml)
        s = self.criterion.mask_out_stride
        tm = tm[:, :, int(s // 2)::s, int(s // 2)::s]
        tm = tm.reshape(B, -1, nf, tm.shape[-2], tm.shape[-1])[bi, ti]
        srcm, tgtm = ml.flatten(1), tm.flatten(1)
        p = srcm.sigmoid()
        dice = 1 - (2 * (p * tgtm).sum(1) + 1) / (p.sum(1) + tgtm.sum(1) + 1)
        losses["loss_mask"] = zero.index_add(0, lv, focal_sum(srcm, tgtm).mean(1)) / num_boxes
        losses["loss_dice"] = zero.index_add(0, lv, dice) / num_boxes
        out = {"class_error": class_error}
        for k, v in losses.items():
            out[k] = v[L - 1]
            for i in range(L - 1):
                out[f"{k}_{i}"] = v[i]
        return out, {"pred_logits": logits[-1], "pred_boxes": boxes[-1], "indices_list": indices_list}

    def forward_sequential(self, frames, targets):
        """The reference's control flow level by level (criterion.py) — the parity anchor for forward_batched."""
        d = self.dense(frames)
        nf = self.nf
        B = d["memory"].shape[0]
        feats = [self.mask_features(d["memory"][:, f], d["shapes"], d["lsi"]) for f in range(nf)]   # nf x [B,8,h,w]
        levels, indices_list = [], []
        for lvl in range(self.dec_layers):
            out = {"pred_logits": self._lin(f"{DETR}class_embed.{lvl}", d["hs"][lvl]), "pred_boxes": d["coords"][lvl]}
            params = self._mlp("detr.controller", d["hs"][lvl])
            indices = self.criterion.matcher(out, targets)
            reference = inverse_sigmoid(d["level_ref"][lvl]).sigmoid()[..., :2]                      # [B,nf,nq,2]
            masks = []
            for b, (qi, _) in enumerate(indices):
                qi = qi.to(self.device)
                h, w = [float(x) for x in targets[b]["size"].tolist()]
                scale = torch.tensor([w, h], device=self.device)                                     # segm:160-166
                per_frame = [self.instance_masks(feats[f][b], reference[b, f][qi] * scale, params[b][qi]) for f in range(nf)]
                masks.append(torch.stack(per_frame, 1)[None])                                        # [1, n, nf, h, w]
            out["pred_masks"] = masks
            levels.append(out)
            indices_list.append(indices)
        outputs = dict(levels[-1])
        outputs["aux_outputs"] = levels[:-1]
        return self.criterion(outputs, targets, indices_list), outputs


class SeqFormerTrainer(IDOLTrainer):
    """IDOLTrainer with SeqFormer's step signature (frames, clip targets)."""

    def __init__(self, model, **kw):
        super().__init__(model, **kw)
        model.criterion.world_size_fn = lambda: self.world          # num_boxes is the cross-rank MEAN GT count here

    def step(self, frames, targets):
        self.flat.g.zero_()
        loss_dict, _ = self.model(frames, targets)
        weighted = self.model.weighted(loss_dict)
        total = sum(weighted.values())
        total.backward()
        return self.apply_gradients(), total.detach(), weighted
