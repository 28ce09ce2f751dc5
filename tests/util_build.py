"""Builds the B200 module for a parity case (oracle/cases.py) with the case's by-name synthetic weights."""
import torch
from transformers import BertModel

from oracle.cases import bert_config
from reftr_b200.modules import BackboneParams, Joiner, PositionEmbeddingSine, RefTR, RefTRSeg, VLTransformerParams
from reftr_b200.synthetic import synthetic_weights


def build_candidate(case, device="cpu", backbone=None, bert=None):
    kw = case["oracle_kw"]
    backbone = backbone or kw.get("backbone", "resnet50")
    torch.manual_seed(1234)
    bert = bert if bert is not None else BertModel(bert_config(case))
    seg = case["seg"]
    bb = Joiner(BackboneParams(backbone, True, seg), PositionEmbeddingSine(128))
    vt = VLTransformerParams(256, 8, kw["enc"], kw["dec"], 2048, kw.get("dropout", 0.0), 1, 128)
    if seg:
        model = RefTRSeg(bb, bert, vt)
    else:
        model = RefTR(bb, bert, vt, aux_loss=kw.get("aux_loss", True))
    synthetic_weights(model, seed=case["wseed"])
    return model.eval().to(device)


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def compare_grads(cand, oracle, skip=()):
    """{name: rel-L2 error} over every parameter that has a gradient in the oracle."""
    og = {n: p.grad for n, p in oracle.named_parameters() if p.grad is not None}
    cg = {n: p.grad for n, p in cand.named_parameters()}
    out = {}
    for n, g in og.items():
        if any(s in n for s in skip):
            continue
        assert cg.get(n) is not None, f"candidate has no gradient for {n}"
        out[n] = rel_l2(cg[n], g)
    return out
