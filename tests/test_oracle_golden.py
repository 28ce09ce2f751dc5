"""Pins oracle/reftr_oracle.py (the CPU restatement) against fixtures produced by the REAL reference
(oracle/make_golden.py, run in the build container).  CPU only."""
import os

import pytest
import torch

from oracle.cases import CASES, build_oracle
from oracle.reftr_oracle import total_box_loss
from reftr_b200.synthetic import synthetic_samples, synthetic_targets

# fp32 on the same CPU: the restatement reorders a few sums (e.g. vectorised masks), so allow ulp-level noise.
ATOL = 2e-5


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_fixture(name, golden_dir):
    case = CASES[name]
    gold = torch.load(os.path.join(golden_dir, f"{name}.pt"), weights_only=False)
    torch.set_num_threads(os.cpu_count())
    model = build_oracle(case)
    # state_dict layout is the drop-in contract
    assert [(k, tuple(v.shape)) for k, v in model.state_dict().items()] == gold["state_dict_keys"]
    out = model(synthetic_samples(**case["inputs"]))
    n_ph = max(case["inputs"].get("n_ph", 0), 1)
    loss = total_box_loss(out, synthetic_targets(case["inputs"]["B"], n_ph))
    if "pred_masks" in out:
        loss = loss + out["pred_masks"].sigmoid().mean()
    loss.backward()
    assert torch.equal(out["phrase_mask"], gold["phrase_mask"])  # discrete output: bit-exact
    assert torch.allclose(out["pred_boxes"], gold["pred_boxes"], atol=ATOL, rtol=0)
    if "aux_boxes" in gold:
        aux = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]])
        assert torch.allclose(aux, gold["aux_boxes"], atol=ATOL, rtol=0)
    if "pred_masks" in gold:
        assert torch.allclose(out["pred_masks"], gold["pred_masks"], atol=2e-4, rtol=1e-4)
        assert torch.allclose(out["mask_att"], gold["mask_att"], atol=ATOL, rtol=1e-4)
    assert abs(float(loss) - float(gold["loss"])) < 1e-4
    params = dict(model.named_parameters())
    assert sum(1 for p in params.values() if p.grad is not None) == gold["n_params_with_grad"]
    assert len(gold["grads"]) >= 20
    for pname, (norm, head) in gold["grads"].items():
        g = params[pname].grad
        assert g is not None, pname
        assert abs(g.norm().item() - norm) <= 1e-3 * norm + 1e-7, (pname, g.norm().item(), norm)
        assert torch.allclose(g.flatten()[:8], head, atol=1e-3 * head.abs().max().item() + 1e-7), pname


@pytest.mark.parametrize("name", ["cfg1_box", "multi_phrase"])
def test_oracle_train_mode_matches_reference_fixture(name, golden_dir):
    """WHERE dropout is applied (transformer.py:151-160, :176-179, :211-223, :241-250; reftr_transformer.py:19; HF BERT): the real
    reference was run in train mode under oracle/det_dropout.py (the k-th dropout call draws mask k); the oracle reproduces its
    outputs and gradients only if it calls dropout the same number of times, in the same order, on the same shapes and layouts."""
    from oracle.det_dropout import deterministic_dropout
    from oracle.make_golden import TRAIN_SEED
    case = dict(CASES[name])
    case["oracle_kw"] = dict(case["oracle_kw"], dropout=0.1)
    gold = torch.load(os.path.join(golden_dir, f"{name}_train.pt"), weights_only=False)
    torch.set_num_threads(os.cpu_count())
    model = build_oracle(case).train()
    model.lang_backbone.config._attn_implementation = "eager"
    n_ph = max(case["inputs"].get("n_ph", 0), 1)
    with deterministic_dropout(TRAIN_SEED) as st:
        out = model(synthetic_samples(**case["inputs"]))
    loss = total_box_loss(out, synthetic_targets(case["inputs"]["B"], n_ph))
    loss.backward()
    assert st["log"] == gold["dropout_calls"] and len(st["log"]) > 10  # same sequence of (call index, shape, p)
    assert torch.allclose(out["pred_boxes"], gold["pred_boxes"], atol=ATOL, rtol=0)
    if "aux_boxes" in gold:
        aux = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]])
        assert torch.allclose(aux, gold["aux_boxes"], atol=ATOL, rtol=0)
    assert abs(float(loss) - float(gold["loss"])) < 1e-4
    eval_gold = torch.load(os.path.join(golden_dir, f"{name}.pt"), weights_only=False)
    assert (gold["pred_boxes"] - eval_gold["pred_boxes"]).abs().max() > 1e-3  # the train-mode fixture really differs from eval
    params = dict(model.named_parameters())
    for pname, (norm, head) in gold["grads"].items():
        g = params[pname].grad
        assert g is not None, pname
        assert abs(g.norm().item() - norm) <= 1e-3 * norm + 1e-7, (pname, g.norm().item(), norm)
