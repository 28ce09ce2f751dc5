"""SURVEY 8(f) N2 pinned to the REFERENCE: tests/golden/*.pt hold what the reference's own criterion (models/criterion.py:101-202,
reftr_segmentation.py:305-337) and post-processors (models/post_process.py:41-82, reftr_segmentation.py:282-302) returned for the
reference model's outputs (oracle/make_golden.py::reference_criterion_and_postprocess), plus the decisions engine_vg.evaluate
derives from them (engine_vg.py:131-140, :152).  reftr_b200.criterion must reproduce all of it: the plain-torch path on CPU, the
one-kernel path (rb_box_loss) on the GPU."""
import os

import pytest
import torch

from oracle.cases import CASES
from reftr_b200.criterion import (CriterionVGMultiPhrase, CriterionVGOnePhraseSeg, PostProcessSegm, PostProcessVGMultiPhrase,
                                  box_cxcywh_to_xyxy)
from reftr_b200.synthetic import criterion_targets, synthetic_mask_targets, synthetic_targets

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _setup(name, device):
    case = CASES[name]
    gold = torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)
    inp = case["inputs"]
    B, H, W = inp["B"], inp["H"], inp["W"]
    n_ph = max(inp.get("n_ph", 0), 1)
    layers = ([b for b in gold["aux_boxes"]] if "aux_boxes" in gold else []) + [gold["pred_boxes"]]
    boxes_all = torch.stack(layers).to(device).requires_grad_(True)
    pm = gold["phrase_mask"].to(device)
    nl = boxes_all.shape[0]
    out = {"pred_boxes": boxes_all[-1], "phrase_mask": pm}
    if nl > 1:
        out["aux_outputs"] = [{"pred_boxes": boxes_all[i], "phrase_mask": pm} for i in range(nl - 1)]
    masks = None
    if "pred_masks" in gold:
        masks = synthetic_mask_targets(B, H, W, device=device)
        out["pred_masks"] = gold["pred_masks"].to(device).requires_grad_(True)
        out["mask_att"] = gold["mask_att"].to(device)
    tgt = synthetic_targets(B, n_ph, device=device)
    targets = criterion_targets(tgt, pm if n_ph > 1 else None, masks, sizes=(H, W))
    wd = gold["crit_weight_dict"]
    crit = (CriterionVGOnePhraseSeg(wd, ["masks", "boxes"]) if case["seg"] else CriterionVGMultiPhrase(wd, ["boxes"])).to(device)
    return case, gold, boxes_all, out, targets, crit


def _check(name, device, tol):
    case, gold, boxes_all, out, targets, crit = _setup(name, device)
    ld = crit(out, targets)
    assert set(ld) == set(gold["crit_losses"]), (sorted(ld), sorted(gold["crit_losses"]))
    for k, v in gold["crit_losses"].items():
        assert abs(float(ld[k]) - float(v)) <= tol * max(1.0, abs(float(v))), (k, float(ld[k]), float(v))
    wd = crit.weight_dict
    total = sum(ld[k] * wd[k] for k in ld if k in wd)   # engine_vg.py:42-43
    assert abs(float(total) - float(gold["crit_total"])) <= tol * max(1.0, abs(float(gold["crit_total"])))
    total.backward()
    g = boxes_all.grad.cpu()
    assert (g - gold["crit_grad_boxes"]).abs().max().item() <= tol * max(1.0, gold["crit_grad_boxes"].abs().max().item())
    if "crit_grad_masks" in gold:
        gm = out["pred_masks"].grad.cpu()
        assert (gm - gold["crit_grad_masks"]).abs().max().item() <= tol * gold["crit_grad_masks"].abs().max().item() + 1e-9
    # ---- post-processing + the evaluation decisions ---------------------------------------------------------------------------
    with torch.no_grad():
        sizes = torch.stack([t["orig_size"] for t in targets])
        post = PostProcessVGMultiPhrase()
        res = post(out, sizes)
        res_scaled = post(out, sizes, scale_to_original_shape=True)
        for r, rs, gb, gbs in zip(res, res_scaled, gold["post_boxes"], gold["post_boxes_scaled"]):
            assert torch.equal(r["boxes"].cpu(), gb) and torch.equal(rs["boxes"].cpu(), gbs)
        iou = torch.cat([_diag_iou(box_cxcywh_to_xyxy(t["boxes"]), r["boxes"]) for t, r in zip(targets, res)]).cpu()
        assert (iou - gold["iou"]).abs().max().item() < 1e-6
        assert torch.equal(iou > 0.5, gold["iou_gt_half"])
        if "post_masks" in gold:
            res = PostProcessSegm()(res, out, sizes, sizes)
            got = torch.stack([r["masks"][0, 0] for r in res]).cpu()
            # the threshold is taken on the bilinear upsampling of the SAME logits: identical arithmetic on CPU -> bit-exact there
            flips = (got != gold["post_masks"]).float().mean().item()
            assert flips <= (0.0 if device == "cpu" else 1e-4), flips


def _diag_iou(a, b):  # util/box_ops.py:22-37 on matched pairs
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, 2:], b[:, 2:]) - torch.max(a[:, :2], b[:, :2])).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    return inter / (area_a + area_b - inter)


@pytest.mark.parametrize("name", list(CASES))
def test_torch_criterion_reproduces_the_reference_fixture(name):
    _check(name, "cpu", 2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_fused_criterion_kernel_reproduces_the_reference_fixture(name):
    """Same fixture through rb_box_loss (all layers' L1 + GIoU and their analytic gradient in one launch)."""
    case, gold, boxes_all, out, targets, crit = _setup(name, "cuda")
    assert crit._all_layer_boxes(out) is not None   # the one-kernel path is the one that runs
    _check(name, "cuda", 1e-5)


@pytest.mark.gpu
def test_box_loss_kernel_with_no_boxes_returns_zeros():
    from reftr_b200 import ops
    boxes = torch.zeros(3, 0, 4, device="cuda")
    losses = torch.full((3, 2), 7.0, device="cuda")
    dl1, dg = torch.empty_like(boxes), torch.empty_like(boxes)
    ops.box_loss(boxes, torch.zeros(0, 4, device="cuda"), None, 1.0, None, losses, dl1, dg)
    assert torch.equal(losses.cpu(), torch.zeros(3, 2))
