"""End-to-end GPU parity: the B200 module (CUDA kernels through the C ABI) against the fp32 oracle on the parity cases of
oracle/cases.py -- same by-name synthetic weights, same seeded inputs.  Every case is stepped three times so the eager
step, the CUDA-graph capture step and a graph replay are all compared."""
import os

import pytest
import torch

from oracle.cases import CASES, build_oracle
from oracle.reftr_oracle import total_box_loss
from reftr_b200.synthetic import synthetic_samples, synthetic_targets

from util_build import build_candidate, compare_grads, rel_l2

pytestmark = pytest.mark.gpu


def _have_seg():
    try:
        import reftr_b200.seg  # noqa: F401
        return True
    except ImportError:
        return False


def _loss(out, case, device):
    n_ph = max(case["inputs"].get("n_ph", 0), 1)
    loss = total_box_loss(out, synthetic_targets(case["inputs"]["B"], n_ph, device=device))
    if "pred_masks" in out:
        loss = loss + out["pred_masks"].sigmoid().mean()
    return loss


def _f32(o):
    if torch.is_tensor(o):
        return o.float() if o.is_floating_point() else o
    if isinstance(o, dict):
        return {k: _f32(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_f32(v) for v in o]
    return o


@pytest.mark.parametrize("name", list(CASES))
def test_e2e_matches_oracle(name):
    case = CASES[name]
    if case["seg"] and not _have_seg():
        pytest.skip("segmentation head (reftr_b200/seg.py) not built yet")
    torch.set_num_threads(os.cpu_count())
    oracle = build_oracle(case)
    s_cpu = synthetic_samples(**case["inputs"])
    out_o = oracle(s_cpu)
    _loss(out_o, case, "cpu").backward()
    # noise floor of the gradients: the ORACLE under bf16 autocast against itself in fp32.  The query encoder's attention has no
    # 1/sqrt(d) scaling (reftr_transformer.py:53), so its softmax is near one-hot and d(linear1/linear2) amplifies any operand
    # rounding: the oracle itself is 0.5 off there under autocast.
    oracle_bf = build_oracle(case)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out_b = oracle_bf(s_cpu)
        loss_b = _loss(_f32(out_b), case, "cpu")
    loss_b.backward()
    og = {n: p.grad for n, p in oracle.named_parameters() if p.grad is not None}
    floor = {n: ((p.grad.float() - og[n]).norm() / (og[n].norm() + 1e-12)).item() for n, p in oracle_bf.named_parameters()
             if p.grad is not None and n in og}
    cand = build_candidate(case, device="cuda")
    from reftr_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    s = synthetic_samples(**case["inputs"], device="cuda")
    norms = {n: p.grad.norm().item() for n, p in oracle.named_parameters() if p.grad is not None}
    big = max(norms.values())
    for step in range(3):
        cand.zero_grad(set_to_none=True)
        out_c = cand(s)
        _loss(out_c, case, "cuda").backward()
        torch.cuda.synchronize()
        assert torch.equal(out_c["phrase_mask"].cpu(), out_o["phrase_mask"])  # discrete output: bit-exact
        err = (out_c["pred_boxes"].cpu() - out_o["pred_boxes"]).abs().max().item()
        rel = rel_l2(out_c["pred_boxes"], out_o["pred_boxes"])
        print(name, "step", step, "pred_boxes max abs err", err, "rel-L2", rel)
        # tolerance = the north star's 1e-3 relative (BASELINE.json), as rel-L2 over the boxes; IEEE-half tensor-core operands with
        # fp32 accumulation / residual stream measure 4.6e-4..8.4e-4 on these cases (the oracle under bf16 autocast is 5e-3..7e-3 off
        # itself).  Boxes are in (0,1): the largest single-coordinate error is bounded as well.
        assert rel < 1e-3
        assert err < 2.5e-3
        if "aux_outputs" in out_o:
            for a, b in zip(out_c["aux_outputs"], out_o["aux_outputs"]):
                assert (a["pred_boxes"].cpu() - b["pred_boxes"]).abs().max().item() < 2.5e-3
        if "pred_masks" in out_o:
            print(name, "pred_masks rel-L2", rel_l2(out_c["pred_masks"], out_o["pred_masks"]), "mask_att rel-L2", rel_l2(out_c["mask_att"], out_o["mask_att"]))
            assert rel_l2(out_c["pred_masks"], out_o["pred_masks"]) < 5e-3
            assert rel_l2(out_c["mask_att"], out_o["mask_att"]) < 5e-3
        errs = compare_grads(cand, oracle)
        live = {n: e for n, e in errs.items() if norms[n] > 1e-6 * big}
        worst = sorted(live.items(), key=lambda kv: -kv[1])[:6]
        print(name, "step", step, "worst grads", worst)
        assert len(errs) > 150
        # (query_encoder.linear1/2 feed the UN-scaled, near one-hot softmax of reftr_transformer.py:53: any operand rounding is
        # amplified there -- the oracle under bf16 autocast is 0.5 off itself on them -- so they get the loose bound)
        # Where the noise floor itself exceeds 1 (pad_box: the oracle under 16-bit autocast is 8.8 .. 9.7 off ITSELF on these three
        # tensors) the fp32 gradient is not determined at 16-bit operand precision at all; only the order of magnitude is asserted.
        lim = lambda n: ((1.5 if floor.get(n, 0.0) > 1.0 else 0.75) if "query_encoder.linear" in n
                         else max(0.3, 3.0 * min(floor.get(n, 0.0), 0.6)))
        bad = {n: (e, floor.get(n)) for n, e in live.items() if e != e or e > lim(n)}
        assert not bad, bad
        med = sorted(live.values())[len(live) // 2]
        print(name, "step", step, "median grad rel-L2", med)
        assert med < 0.1
    eng = cand.engine()
    assert eng.launches > 0
    if eng.use_graphs:
        assert any(st["fwd"] is not None and st["bwd"] is not None for st in eng._states.values())


def test_missing_library_fails_loudly(monkeypatch):
    """No CPU / PyTorch fallback: CPU tensors are refused."""
    case = CASES["cfg1_box"]
    cand = build_candidate(case, device="cpu")
    with pytest.raises(RuntimeError):
        cand(synthetic_samples(**case["inputs"]))


@pytest.mark.parametrize("name", ["cfg1_box", "multi_phrase"])
def test_split_backward_graphs_match_single_graph(name, monkeypatch):
    """The multi-graph backward used under data parallelism (engine._run_backward_split: every part's gradient slice is exchanged while
    the following parts still compute) must produce the gradients of the single-graph backward."""
    case = CASES[name]
    s = synthetic_samples(**case["inputs"], device="cuda")

    def grads(split):
        monkeypatch.setenv("REFTR_B200_SPLIT_BWD", "1" if split else "0")
        cand = build_candidate(case, device="cuda")
        for _ in range(4):  # eager, capture, replay, replay
            cand.zero_grad(set_to_none=True)
            _loss(cand(s), case, "cuda").backward()
        torch.cuda.synchronize()
        eng = cand.engine()
        assert any((st.get("bwd3") is not None) == split for st in eng._states.values() if st["fwd"] is not None)
        return {n: p.grad.detach().clone() for n, p in cand.named_parameters() if p.grad is not None}

    a, b = grads(False), grads(True)
    assert a.keys() == b.keys() and len(a) > 150
    big = max(v.norm().item() for v in a.values())
    for n in a:
        if a[n].norm().item() > 1e-6 * big:
            assert rel_l2(b[n], a[n]) < 2e-3, n  # fp32 atomics order is the only difference
