"""Stand-alone restatement of the reference's loss / post-processing modules, used when the reference package is not on
sys.path (bench.py, tests).  Same constructor arguments, ``weight_dict`` attribute, input and output dict keys as
models/criterion.py:101-202 (CriterionVGMultiPhrase), models/reftr_segmentation.py:282-337 (PostProcessSegm,
CriterionVGOnePhraseSeg) and models/post_process.py:41-82 (PostProcessVGMultiPhrase).

Differences that do not change results: the paired GIoU is computed directly instead of taking the
diagonal of the N x N matrix (box_ops.py:52-77), the per-sample Python loop of masked_select is one boolean index, and
``num_boxes`` stays on the host when no process group is initialised (no ``.item()`` sync).
"""
import torch
import torch.nn.functional as F
from torch import nn


def box_cxcywh_to_xyxy(x):  # util/box_ops.py:9-13
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def paired_giou(a, b):
    """diag(generalized_box_iou(a, b)) for xyxy boxes (util/box_ops.py:40-77)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, 2:], b[:, 2:]) - torch.max(a[:, :2], b[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = area_a + area_b - inter
    iou = inter / union
    wh2 = (torch.max(a[:, 2:], b[:, 2:]) - torch.min(a[:, :2], b[:, :2])).clamp(min=0)
    area = wh2[:, 0] * wh2[:, 1]
    return iou - (area - union) / area


def dice_loss(inputs, targets, num_boxes):  # models/modeling/segmentation.py:178-194
    inputs = inputs.sigmoid().flatten(1)
    numerator = 2 * (inputs * targets).sum(1)
    denominator = inputs.sum(-1) + targets.sum(-1)
    return (1 - (numerator + 1) / (denominator + 1)).sum() / num_boxes


def sigmoid_focal_loss(inputs, targets, num_boxes, alpha=0.25, gamma=2):  # models/modeling/segmentation.py:197-221
    prob = inputs.sigmoid()
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = prob * targets + (1 - prob) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss.mean(1).sum() / num_boxes


class _FusedBoxLoss(torch.autograd.Function):
    """L1 + GIoU for every decoder layer in ONE kernel (rb_box_loss); the kernel also emits d(loss)/d(boxes), so the backward is
    two broadcast multiplies.  Returns the 2 x n_layers losses as SEPARATE 0-dim outputs (views of one buffer): the training loop
    weights and sums them one by one (engine_vg.py:42-43), and as ``L[i, j]`` selections of one output every term's backward was
    two SelectBackward nodes = zeros + copy + accumulate, ~60 eager launches (0.25 ms of an idle GPU) before the backward graph."""

    @staticmethod
    def forward(ctx, boxes_all, tgt, valid, inv_norm, inv_norm_dev):
        from . import ops
        nl, N = boxes_all.shape[0], boxes_all.shape[1]
        losses = torch.empty(nl, 2, dtype=torch.float32, device=boxes_all.device)
        dl1, dgiou = torch.empty_like(boxes_all), torch.empty_like(boxes_all)
        ops.box_loss(boxes_all, tgt, valid, inv_norm, inv_norm_dev, losses, dl1, dgiou)
        ctx.save_for_backward(dl1, dgiou)
        ctx.set_materialize_grads(False)
        return tuple(losses.view(-1).unbind(0))  # (l1_0, giou_0, l1_1, giou_1, ...)

    @staticmethod
    def backward(ctx, *gs):
        dl1, dgiou = ctx.saved_tensors
        nl = dl1.shape[0]
        if any(g is None for g in gs):
            zero = torch.zeros((), dtype=torch.float32, device=dl1.device)
            gs = [zero if g is None else g for g in gs]
        g = torch.stack([x.reshape(()) for x in gs]).view(nl, 2)
        return torch.addcmul(dl1 * g[:, 0].view(nl, 1, 1), dgiou, g[:, 1].view(nl, 1, 1)), None, None, None, None


class CriterionVGMultiPhrase(nn.Module):
    def __init__(self, weight_dict, losses):
        super().__init__()
        self.weight_dict = weight_dict
        self.losses = losses

    @staticmethod
    def _all_layer_boxes(outputs):
        """[n_layers, B, n_ph, k, 4] (last layer last) when the model's per-layer ``pred_boxes`` are slices of ONE tensor (reftr_b200's
        modules: ``coord[-1]`` and ``coord[i]`` of the same sigmoid output), else None.  Found through ``Tensor._base`` so the output
        dict keeps exactly the reference's keys (reftr_transformer.py:293-304)."""
        pb = outputs["pred_boxes"]
        aux = outputs.get("aux_outputs") or []
        if not pb.is_cuda:
            return None
        if not aux:
            return pb.unsqueeze(0)
        base = pb._base
        nl = len(aux) + 1
        if base is None or base.dim() != pb.dim() + 1 or base.shape[0] != nl or tuple(base.shape[1:]) != tuple(pb.shape) or not base.is_contiguous():
            return None
        step = base[0].numel() * base.element_size()
        if pb.data_ptr() != base.data_ptr() + (nl - 1) * step:
            return None
        for i, a in enumerate(aux):
            ab = a["pred_boxes"]
            if ab._base is not base or ab.data_ptr() != base.data_ptr() + i * step:
                return None
        return base

    def _fused_boxes(self, allb, outputs, targets, num_boxes):
        """All layers' box losses in one kernel from the stacked boxes [n_layers, B, n_ph, k, 4] (last layer last) -- sync-free."""
        nl, b, n_ph, k, _ = allb.shape
        tgt = torch.cat([t["boxes"] for t in targets], dim=0).to(torch.float32)
        valid = None
        if tgt.shape[0] != b * n_ph:  # padded phrases: scatter the valid targets to their (sample, phrase) slots
            pm = outputs["phrase_mask"].view(b, n_ph, k)[:, :, 0].reshape(b * n_ph)
            full = torch.zeros(b * n_ph, 4, dtype=torch.float32, device=allb.device)
            full.masked_scatter_(pm.unsqueeze(-1).expand(-1, 4), tgt)
            tgt = full
            valid = pm.unsqueeze(-1).expand(-1, k).reshape(-1).to(torch.uint8).contiguous()
        tgt = tgt.unsqueeze(1).expand(-1, k, -1).reshape(b * n_ph * k, 4).contiguous()
        if torch.is_tensor(num_boxes):
            inv, inv_dev = 0.0, (1.0 / (num_boxes.to(torch.float32) * k)).reshape(1).contiguous()
        else:
            inv, inv_dev = 1.0 / (num_boxes * k), None
        L = _FusedBoxLoss.apply(allb.reshape(nl, b * n_ph * k, 4).contiguous(), tgt, valid, inv, inv_dev)
        losses = {"loss_bbox": L[2 * (nl - 1)], "loss_giou": L[2 * (nl - 1) + 1]}
        for i in range(nl - 1):
            losses[f"loss_bbox_{i}"] = L[2 * i]
            losses[f"loss_giou_{i}"] = L[2 * i + 1]
        return losses

    def loss_boxes(self, outputs, targets, num_boxes):  # criterion.py:113-153
        src = outputs["pred_boxes"]
        b, n_ph, k, _ = src.shape
        tgt = torch.cat([t["boxes"] for t in targets], dim=0)
        if tgt.shape[0] == b * n_ph:  # every phrase valid (all single-phrase configs): no data-dependent shapes
            pred = src.reshape(b * n_ph, k, 4)
        else:
            pred = src[outputs["phrase_mask"].view(b, n_ph, k)[:, :, 0]]
        assert pred.shape[0] == tgt.shape[0]
        tgt = tgt.unsqueeze(1).expand(-1, k, -1).reshape(-1, 4)
        pred = pred.reshape(-1, 4)
        losses = {"loss_bbox": F.l1_loss(pred, tgt, reduction="none").sum() / (num_boxes * k)}
        giou = paired_giou(box_cxcywh_to_xyxy(pred), box_cxcywh_to_xyxy(tgt))
        losses["loss_giou"] = (1 - giou).sum() / (num_boxes * k)
        return losses

    def loss_masks(self, outputs, targets, num_boxes):
        raise NotImplementedError

    def get_loss(self, loss, outputs, targets, num_boxes, **kw):
        return {"boxes": self.loss_boxes, "masks": self.loss_masks}[loss](outputs, targets, num_boxes, **kw)

    def forward(self, outputs, targets):  # criterion.py:166-202
        num_boxes = float(sum(len(t["labels"]) for t in targets))
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            # the reference calls .item() here (criterion.py:180): a host sync per step.  The count stays a device scalar instead.
            # torch.full = a fill kernel; torch.as_tensor([..], device=cuda) would be a pageable H2D copy, i.e. a host sync
            nb = torch.full((1,), num_boxes, dtype=torch.float, device=outputs["pred_boxes"].device)
            torch.distributed.all_reduce(nb)
            num_boxes = torch.clamp(nb / torch.distributed.get_world_size(), min=1)[0]
        else:
            num_boxes = max(num_boxes, 1.0)
        losses = {}
        allb = self._all_layer_boxes(outputs) if "boxes" in self.losses else None
        if allb is not None:
            losses.update(self._fused_boxes(allb, outputs, targets, num_boxes))
            for loss in self.losses:
                if loss != "boxes":
                    losses.update(self.get_loss(loss, outputs, targets, num_boxes))
            return losses
        for loss in self.losses:
            losses.update(self.get_loss(loss, outputs, targets, num_boxes))
        for i, aux in enumerate(outputs.get("aux_outputs", [])):
            for loss in self.losses:
                if loss == "masks":
                    continue
                losses.update({k + f"_{i}": v for k, v in self.get_loss(loss, aux, targets, num_boxes).items()})
        return losses


class CriterionVGOnePhraseSeg(CriterionVGMultiPhrase):
    def loss_masks(self, outputs, targets, num_boxes):  # reftr_segmentation.py:314-337
        src = outputs["pred_masks"]
        bs, num_q = src.shape[:2]
        masks = [t["masks"] for t in targets]
        hh, ww = max(m.shape[-2] for m in masks), max(m.shape[-1] for m in masks)
        tgt = torch.zeros((bs, masks[0].shape[0], hh, ww), dtype=src.dtype, device=src.device)  # nested_tensor_from_tensor_list
        for i, m in enumerate(masks):
            tgt[i, :, :m.shape[-2], :m.shape[-1]] = m.to(src.dtype)
        src = F.interpolate(src, size=(hh, ww), mode="bilinear", align_corners=False).view(bs * num_q, -1)
        tgt = tgt.view(bs * num_q, -1)
        losses = {"loss_mask": sigmoid_focal_loss(src, tgt, bs * num_q), "loss_dice": dice_loss(src, tgt, bs * num_q)}
        if "cem_loss" in outputs:
            losses["loss_cem"] = outputs["cem_loss"]
        return losses


class PostProcessVGMultiPhrase(nn.Module):
    @torch.no_grad()
    def forward(self, outputs, target_sizes, scale_to_original_shape=False):  # post_process.py:45-82
        out_bbox = outputs["pred_boxes"]
        bsz, n_ph, k, _ = out_bbox.shape
        mask = outputs["phrase_mask"].view(bsz, n_ph, k)
        assert bsz == len(target_sizes) and target_sizes.shape[1] == 2
        results = []
        for i in range(bsz):
            boxes = box_cxcywh_to_xyxy(out_bbox[i][mask[i][:, 0]][:, 0, :])
            if scale_to_original_shape:
                img_h, img_w = target_sizes[i:i + 1].unbind(1)
                boxes = boxes * torch.stack([img_w, img_h, img_w, img_h], dim=1)
            results.append({"boxes": boxes})
        return results


class PostProcessSegm(nn.Module):
    def __init__(self, threshold=0.5):
        super().__init__()
        self.threshold = threshold

    @torch.no_grad()
    def forward(self, results, outputs, orig_target_sizes, max_target_sizes):  # reftr_segmentation.py:287-302
        assert len(orig_target_sizes) == len(max_target_sizes)
        max_h, max_w = max_target_sizes.max(0)[0].tolist()
        m = F.interpolate(outputs["pred_masks"].squeeze(2), size=(max_h, max_w), mode="bilinear", align_corners=False)
        m = m.sigmoid() > self.threshold
        for i, (cur, t, tt) in enumerate(zip(m, max_target_sizes, orig_target_sizes)):
            results[i]["masks"] = cur[:, :t[0], :t[1]].unsqueeze(1)
            results[i]["masks_origin"] = F.interpolate(results[i]["masks"].float(), size=tuple(tt.tolist()), mode="nearest").byte()
        return results
