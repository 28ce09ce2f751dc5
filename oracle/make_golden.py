"""Generate tests/golden/*.pt by running the REAL reference (imported from /root/reference) on CPU.

Run in the build container only (``python oracle/make_golden.py``); the GPU box has no /root/reference.
Recipe: put the reference on sys.path, patch ``is_main_process`` so torchvision
does not try to download ResNet weights, patch ``BertModel.from_pretrained`` to build a random-init BERT
from a config (no network), build with the reference's own argparse parser, load by-name synthetic
weights, run forward + criterion-equivalent loss + backward, and store inputs' recipe + outputs.

The fixtures hold only outputs (a few KB each): weights and inputs are regenerated from seeds by
``reftr_b200.synthetic`` on the consumer side.
"""
import argparse
import os
import sys

import torch

REF = os.environ.get("REFTR_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from reftr_b200.synthetic import (criterion_targets, synthetic_mask_targets, synthetic_samples, synthetic_targets,  # noqa: E402
                                  synthetic_weights)
from oracle.cases import CASES, bert_config  # noqa: E402

TRAIN_CASES = ("cfg1_box", "multi_phrase")


def load_reference():
    import numpy as np  # noqa: F401  (main_vg's parser uses np.pi)
    import models.modeling.backbone as rb
    rb.is_main_process = lambda: False
    import models.reftr_transformer as rt
    import models.reftr_segmentation as rs
    from transformers import BertModel

    class _FakeBert:
        cfg = None

        @staticmethod
        def from_pretrained(name):
            torch.manual_seed(1234)
            return BertModel(_FakeBert.cfg)

    rt.BertModel = rs.BertModel = _FakeBert
    rt.RobertaModel = rs.RobertaModel = _FakeBert
    src = open(os.path.join(REF, "main_vg.py")).read()
    a, b = src.index("def get_args_parser"), src.index("def main(args)")
    ns = {"argparse": argparse, "np": np}
    exec(src[a:b], ns)
    import models
    from util.misc import NestedTensor
    return models, ns["get_args_parser"], NestedTensor, _FakeBert


TRAIN_SEED = 11  # seed of oracle/det_dropout.py for the train-mode fixtures


def run_case(name, case, models, get_args_parser, NestedTensor, fake_bert, train=False):
    """train=True: the reference runs in train mode (--dropout 0.1) under oracle/det_dropout.py, and the fixture also stores the
    sequence of dropout calls (index, shape, p) -- it pins WHERE the oracle applies dropout, not just the eval-mode arithmetic."""
    import contextlib
    from oracle.det_dropout import deterministic_dropout
    from oracle.reftr_oracle import total_box_loss
    parser = argparse.ArgumentParser(parents=[get_args_parser()])
    flags = list(case["flags"])
    if train:
        flags[flags.index("--dropout") + 1] = "0.1"
    args = parser.parse_args(flags + ["--device", "cpu"])
    fake_bert.cfg = bert_config(case)
    if train:
        fake_bert.cfg._attn_implementation = "eager"  # sdpa draws its dropout inside a fused kernel; the eager path calls F.dropout
    torch.manual_seed(0)
    model, criterion, post = models.build_reftr(args)
    synthetic_weights(model, seed=case["wseed"])
    model.train() if train else model.eval()  # eval: dropout inactive
    s = synthetic_samples(**case["inputs"])
    samples = dict(s)
    samples["img"] = NestedTensor(s["img"].tensors, s["img"].mask)
    ctx = deterministic_dropout(TRAIN_SEED) if train else contextlib.nullcontext()
    with ctx as dstate:
        out = model(samples)
    n_ph = max(case["inputs"].get("n_ph", 0), 1)
    tgt = synthetic_targets(case["inputs"]["B"], n_ph)
    loss = total_box_loss(out, tgt)
    if "pred_masks" in out:
        loss = loss + out["pred_masks"].sigmoid().mean()
    loss.backward()
    gold = {"pred_boxes": out["pred_boxes"].detach(), "phrase_mask": out["phrase_mask"], "loss": loss.detach()}
    if out.get("aux_outputs"):
        gold["aux_boxes"] = torch.stack([a["pred_boxes"].detach() for a in out["aux_outputs"]])
    if "pred_masks" in out:
        gold["pred_masks"] = out["pred_masks"].detach()
        gold["mask_att"] = out["mask_att"].detach()
    if not train:
        gold.update(reference_criterion_and_postprocess(case, out, tgt, criterion, post))
    grads = {}
    for pname, p in model.named_parameters():
        if p.grad is not None and case["grad_filter"](pname):
            grads[pname] = (p.grad.norm().item(), p.grad.flatten()[:8].clone())
    gold["grads"] = grads
    gold["n_params_with_grad"] = sum(1 for _, p in model.named_parameters() if p.grad is not None)
    gold["state_dict_keys"] = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    if train:
        gold["dropout_calls"] = dstate["log"]
        name = name + "_train"
    path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
    torch.save(gold, path)
    print(name, "loss", float(loss), "boxes", gold["pred_boxes"].flatten()[:4].tolist(), "->", path,
          os.path.getsize(path), "bytes")


def reference_criterion_and_postprocess(case, out, tgt, criterion, post):
    """Pins SURVEY 8(f) N2 and the discrete outputs of 3.4 to the REFERENCE: runs the reference's own criterion
    (models/criterion.py:101-202; CriterionVGOnePhraseSeg reftr_segmentation.py:305-337 for --masks) on the reference model's
    outputs, storing every entry of its loss dict, the weighted total (engine_vg.py:42-43) and its gradient w.r.t. every layer's boxes
    (and the mask logits); then the reference's post-processors and the decisions engine_vg.evaluate takes from them:
    ``iou > 0.5`` per sample (engine_vg.py:131-140) and ``sigmoid > 0.5`` masks (reftr_segmentation.py:288-302) + mask IoU (:152)."""
    from util.box_ops import box_cxcywh_to_xyxy, box_iou, mask_iou
    inp = case["inputs"]
    B, H, W = inp["B"], inp["H"], inp["W"]
    aux = out.get("aux_outputs") or []
    boxes_all = torch.stack([a["pred_boxes"].detach() for a in aux] + [out["pred_boxes"].detach()]).requires_grad_(True)
    pm = out["phrase_mask"]
    o2 = {"pred_boxes": boxes_all[-1], "phrase_mask": pm}
    if aux:
        o2["aux_outputs"] = [{"pred_boxes": boxes_all[i], "phrase_mask": pm} for i in range(len(aux))]
    masks = None
    if "pred_masks" in out:
        masks = synthetic_mask_targets(B, H, W)
        o2["pred_masks"] = out["pred_masks"].detach().requires_grad_(True)
        o2["mask_att"] = out["mask_att"].detach()
    targets = criterion_targets(tgt, pm if tgt.shape[1] > 1 else None, masks, sizes=(H, W))
    ld = criterion(o2, targets)
    wd = criterion.weight_dict
    total = sum(ld[k] * wd[k] for k in ld if k in wd)
    total.backward()
    res = {"crit_losses": {k: v.detach().clone() for k, v in ld.items()}, "crit_weight_dict": dict(wd), "crit_total": total.detach(),
           "crit_grad_boxes": boxes_all.grad.clone()}
    if masks is not None:
        res["crit_grad_masks"] = o2["pred_masks"].grad.clone()
    with torch.no_grad():
        sizes = torch.stack([t["orig_size"] for t in targets])
        results = post["bbox"](o2, sizes)
        ious = []
        for i, r in enumerate(results):
            iou, _ = box_iou(box_cxcywh_to_xyxy(targets[i]["boxes"]), r["boxes"])
            ious.append(torch.diag(iou))
        res["post_boxes"] = [r["boxes"].clone() for r in results]
        res["post_boxes_scaled"] = [r["boxes"].clone() for r in post["bbox"](o2, sizes, scale_to_original_shape=True)]
        res["iou"] = torch.cat(ious)
        res["iou_gt_half"] = res["iou"] > 0.5
        if "segm" in post:
            results = post["segm"](results, o2, sizes, sizes)
            res["post_masks"] = torch.stack([r["masks"][0, 0] for r in results])
            res["seg_iou"] = torch.stack([mask_iou(r["masks"][0][0], targets[i]["masks"]) for i, r in enumerate(results)])
    return res


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    ref = load_reference()
    only = sys.argv[1:] or list(CASES)
    for name in only:
        run_case(name, CASES[name], *ref)
    for name in TRAIN_CASES:
        if name in only:
            run_case(name, CASES[name], *ref, train=True)
