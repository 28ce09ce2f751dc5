"""Host-logic check (CPU, no GPU): the engine's orchestration (reftr_b200/engine.py) run on tests/emu_ops.py -- a torch
emulation of the C-ABI kernels -- must reproduce the oracle's outputs and gradients up to bf16 operand rounding.
The GPU parity of the kernels themselves is in tests/test_*_gpu.py; the end-to-end GPU parity in tests/test_e2e_gpu.py."""
import os

import pytest
import torch

from oracle.cases import CASES, build_oracle
from oracle.reftr_oracle import total_box_loss
from reftr_b200.synthetic import synthetic_samples, synthetic_targets

import emu_ops
from util_build import build_candidate, compare_grads, rel_l2


class _TorchFp32Proxy:
    """``torch`` as seen by engine.py / pack.py / seg.py in EXACT mode: every "bf16" buffer becomes fp32."""

    def __getattr__(self, name):
        return torch.float32 if name == "bfloat16" else getattr(torch, name)


def _modules():
    import reftr_b200.engine as engine
    import reftr_b200.pack as pack
    import reftr_b200.bert as bert
    mods = [engine, pack, bert]
    try:
        import reftr_b200.seg as seg
        mods.append(seg)
    except ImportError:
        pass
    return mods


@pytest.fixture()
def emulated(monkeypatch):
    for m in _modules():
        monkeypatch.setattr(m, "ops", emu_ops)
    yield


@pytest.fixture()
def emulated_exact(monkeypatch):
    for m in _modules():
        monkeypatch.setattr(m, "ops", emu_ops)
        monkeypatch.setattr(m, "torch", _TorchFp32Proxy())
    emu_ops.EXACT[0] = True
    yield
    emu_ops.EXACT[0] = False


def _have_seg():
    try:
        import reftr_b200.seg  # noqa: F401
        return True
    except ImportError:
        return False


def _loss(out, case):
    n_ph = max(case["inputs"].get("n_ph", 0), 1)
    loss = total_box_loss(out, synthetic_targets(case["inputs"]["B"], n_ph))
    if "pred_masks" in out:
        loss = loss + out["pred_masks"].sigmoid().mean()
    return loss


@pytest.mark.parametrize("name", [n for n in CASES])
def test_engine_on_emulated_kernels_matches_oracle(name, emulated):
    case = CASES[name]
    if case["seg"] and not _have_seg():
        pytest.skip("segmentation head (reftr_b200/seg.py) not built yet")
    torch.set_num_threads(os.cpu_count())
    oracle = build_oracle(case)
    cand = build_candidate(case)
    assert [(k, tuple(v.shape)) for k, v in cand.state_dict().items()] == [(k, tuple(v.shape)) for k, v in oracle.state_dict().items()]
    s = synthetic_samples(**case["inputs"])
    out_o = oracle(s)
    _loss(out_o, case).backward()
    out_c = cand(s)
    _loss(out_c, case).backward()
    assert torch.equal(out_c["phrase_mask"], out_o["phrase_mask"])
    err = (out_c["pred_boxes"] - out_o["pred_boxes"]).abs().max().item()
    print(name, "pred_boxes max abs err", err)
    assert err < 2.5e-3  # IEEE-half operand rounding (measured 3e-4..9e-4); the oracle under bf16 autocast is 5e-3..7e-3 off itself
    if "aux_outputs" in out_o:
        for a, b in zip(out_c["aux_outputs"], out_o["aux_outputs"]):
            assert (a["pred_boxes"] - b["pred_boxes"]).abs().max().item() < 2.5e-3
    if "pred_masks" in out_o:
        assert rel_l2(out_c["pred_masks"], out_o["pred_masks"]) < 5e-3
        assert rel_l2(out_c["mask_att"], out_o["mask_att"]) < 5e-3
    errs = compare_grads(cand, oracle)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
    print(name, "worst grads", worst)
    assert len(errs) > 150
    # Gradient tolerance in bf16 mode: a ReLU whose pre-activation is within the bf16 noise of zero flips its mask, and a
    # fraction p of flipped units costs sqrt(p) in rel-L2, so ~1% forward noise gives 10..40% on deep-layer gradients.
    # Measured floor on cfg1_box: the ORACLE under torch.autocast(bf16) vs itself in fp32 is 0.38 (layer2.0.conv2),
    # 0.18 (layer4.2.conv2), 0.20 (encoder linear1), 0.06 (bbox_embed.0).  test_engine_wiring_exact is the tight check.
    norms = {n: p.grad.norm().item() for n, p in oracle.named_parameters() if p.grad is not None}
    big = max(norms.values())
    live = {n: e for n, e in errs.items() if norms[n] > 1e-6 * big}
    bad = {n: e for n, e in live.items() if e > 0.75}
    assert not bad, bad
    assert sorted(live.values())[len(live) // 2] < 0.1


def _linear_loss(out):
    g = torch.Generator().manual_seed(5)
    boxes = [out["pred_boxes"]] + [a["pred_boxes"] for a in out.get("aux_outputs", [])]
    loss = sum((b * torch.randn(b.shape, generator=g)).sum() for b in boxes)
    if "pred_masks" in out:
        loss = loss + (out["pred_masks"] * torch.randn(out["pred_masks"].shape, generator=g)).mean()
        loss = loss + (out["mask_att"] * torch.randn(out["mask_att"].shape, generator=g)).sum()
    return loss


@pytest.mark.parametrize("name", [n for n in CASES])
def test_engine_wiring_exact(name, emulated_exact):
    """EXACT mode (no bf16 rounding anywhere): the engine's forward and backward wiring must agree with the oracle to
    fp32 accuracy -- this isolates host-logic errors from the (ReLU-mask-flip dominated) bf16 noise of gradients."""
    case = CASES[name]
    if case["seg"] and not _have_seg():
        pytest.skip("segmentation head (reftr_b200/seg.py) not built yet")
    torch.set_num_threads(os.cpu_count())
    oracle = build_oracle(case)
    cand = build_candidate(case)
    s = synthetic_samples(**case["inputs"])
    out_o = oracle(s)
    _linear_loss(out_o).backward()
    out_c = cand(s)
    _linear_loss(out_c).backward()
    assert (out_c["pred_boxes"] - out_o["pred_boxes"]).abs().max().item() < 2e-5
    if "pred_masks" in out_o:
        assert rel_l2(out_c["pred_masks"], out_o["pred_masks"]) < 1e-4
        assert rel_l2(out_c["mask_att"], out_o["mask_att"]) < 1e-4
    errs = compare_grads(cand, oracle)
    norms = {n: p.grad.norm().item() for n, p in oracle.named_parameters() if p.grad is not None}
    big = max(norms.values())
    bad = {n: e for n, e in errs.items() if e > 1e-2 and norms[n] > 1e-6 * big}  # exclude mathematically-zero gradients (key biases)
    print(name, sorted(errs.items(), key=lambda kv: -kv[1])[:6])
    assert not bad, bad


def _train_case(name):
    case = dict(CASES[name])
    case["oracle_kw"] = dict(case["oracle_kw"], dropout=0.1)
    return case


@pytest.mark.parametrize("name", ["cfg1_box", "multi_phrase", "pad_box"])
def test_engine_train_mode_dropout_exact(name, emulated_exact):
    """Train mode (transformer.py:151-160, :211-223 dropouts, mlp_mapping's nn.Dropout, HF BERT's dropouts): with the oracle
    drawing the SAME counter-based masks (tests/dropout_ref.py), outputs and every gradient must agree to fp32 accuracy --
    this pins where each mask is applied in the forward and how it re-enters the backward."""
    import oracle.reftr_oracle as orc
    import dropout_ref
    case = _train_case(name)
    torch.set_num_threads(os.cpu_count())
    oracle = build_oracle(case).train()
    cand = build_candidate(case).train()
    seed = 0x1234ABCD5678EF01 & 0x7FFFFFFFFFFFFFFF
    hook = dropout_ref.OracleDropoutHook(seed)
    undo = dropout_ref.hook_hf_bert(oracle.lang_backbone, hook, orc.DROP_CTX)
    orc.DROPOUT_HOOK = hook
    try:
        s = synthetic_samples(**case["inputs"])
        out_o = oracle(s)
        _linear_loss(out_o).backward()
    finally:
        orc.DROPOUT_HOOK = None
        undo()
    eng = cand.engine()
    eng.next_seed = seed
    out_c = cand(s)
    assert eng.last_seed == seed and eng.train_mode
    _linear_loss(out_c).backward()
    # every site the oracle visited exists in the engine under the same name, and dropout really happened
    assert set(hook.seen) == set(eng._drops), (set(hook.seen) ^ set(eng._drops))
    assert len(hook.seen) > 10
    eval_out = build_oracle(case)(s)["pred_boxes"]
    assert (out_o["pred_boxes"] - eval_out).abs().max().item() > 1e-4  # train output differs from the eval output
    assert (out_c["pred_boxes"] - out_o["pred_boxes"]).abs().max().item() < 2e-5
    errs = compare_grads(cand, oracle)
    norms = {n: p.grad.norm().item() for n, p in oracle.named_parameters() if p.grad is not None}
    big = max(norms.values())
    bad = {n: e for n, e in errs.items() if e > 1e-2 and norms[n] > 1e-6 * big}
    print(name, sorted(errs.items(), key=lambda kv: -kv[1])[:6])
    assert not bad, bad


def test_dropout_generator_statistics():
    """The counter-based generator: keep rate = 1 - p to sampling accuracy, different sites / seeds decorrelate, and the
    two 16-bit lanes of a word are independent."""
    import dropout_ref
    p = 0.1
    m1 = dropout_ref.keep_mask(12345, dropout_ref.site_id("enc0.attn"), 2048, 420, p)
    m2 = dropout_ref.keep_mask(12345, dropout_ref.site_id("enc1.attn"), 2048, 420, p)
    m3 = dropout_ref.keep_mask(12346, dropout_ref.site_id("enc0.attn"), 2048, 420, p)
    n = m1.numel()
    for m in (m1, m2, m3):
        assert abs(m.float().mean().item() - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5 + 1e-5
    for a, b in ((m1, m2), (m1, m3), (m1[:, 0::2], m1[:, 1::2]), (m1[:-1], m1[1:])):
        both = (~a & ~b).float().mean().item()  # P(both dropped) = p^2 when independent
        assert abs(both - p * p) < 1.5e-3, both


def test_engine_wiring_exact_two_queries_per_phrase(emulated_exact):
    """--num_queries_per_phrase 2 with the multi-phrase input (T = n_ph * n_q = 6 decoder queries; the query encoder tiles every
    phrase feature over the learned query embeddings, reftr_transformer.py:62-66): not a shipped config, but part of the surface."""
    from transformers import BertModel
    from oracle.cases import bert_config
    from oracle.reftr_oracle import RefTROracle
    from reftr_b200.modules import BackboneParams, Joiner, PositionEmbeddingSine, RefTR, VLTransformerParams
    from reftr_b200.synthetic import synthetic_weights
    case = CASES["multi_phrase"]
    torch.set_num_threads(os.cpu_count())

    def build(kind):
        torch.manual_seed(1234)
        b = BertModel(bert_config(case))
        if kind == "oracle":
            m = RefTROracle(b, enc=1, dec=2, dropout=0.0, aux_loss=True, n_q=2)
        else:
            m = RefTR(Joiner(BackboneParams("resnet50", True, False), PositionEmbeddingSine(128)), b,
                      VLTransformerParams(256, 8, 1, 2, 2048, 0.0, 1, 128), num_queries_per_phrase=2, aux_loss=True)
        synthetic_weights(m, seed=5)
        return m.eval()

    oracle, cand = build("oracle"), build("cand")
    s = synthetic_samples(**case["inputs"])
    out_o, out_c = oracle(s), cand(s)
    assert out_c["pred_boxes"].shape == out_o["pred_boxes"].shape == (2, 3, 2, 4)
    assert torch.equal(out_c["phrase_mask"], out_o["phrase_mask"])
    assert (out_c["pred_boxes"] - out_o["pred_boxes"]).abs().max().item() < 2e-5
    _linear_loss(out_o).backward()
    _linear_loss(out_c).backward()
    errs = compare_grads(cand, oracle)
    norms = {n: p.grad.norm().item() for n, p in oracle.named_parameters() if p.grad is not None}
    big = max(norms.values())
    bad = {n: e for n, e in errs.items() if e > 1e-2 and norms[n] > 1e-6 * big}
    assert not bad, bad


def test_engine_wiring_exact_frozen_bert(emulated_exact):
    """A frozen language backbone (requires_grad False on every lang_backbone parameter, set by the training script; the reference's
    own --freeze_bert flag is stored but never applied, reftr_transformer.py:128, :152-157): BERT still runs on the kernels, receives
    no gradients, its backward is skipped, and every other gradient is unchanged."""
    case = CASES["cfg1_box"]
    torch.set_num_threads(os.cpu_count())
    oracle, cand = build_oracle(case), build_candidate(case)
    for m in (oracle, cand):
        for p in m.lang_backbone.parameters():
            p.requires_grad_(False)
    s = synthetic_samples(**case["inputs"])
    out_o, out_c = oracle(s), cand(s)
    _linear_loss(out_o).backward()
    _linear_loss(out_c).backward()
    assert cand.engine().bert is not None and not cand.engine().bert.trainable
    assert all(p.grad is None for p in cand.lang_backbone.parameters())
    errs = compare_grads(cand, oracle)
    norms = {n: p.grad.norm().item() for n, p in oracle.named_parameters() if p.grad is not None}
    big = max(norms.values())
    assert len(errs) > 100 and not {n: e for n, e in errs.items() if e > 1e-2 and norms[n] > 1e-6 * big}


def test_backward_of_a_stale_forward_raises(emulated):
    """The engine keeps the activations of the LAST forward only: a second forward (an evaluation inside the training loop, another
    micro-batch) before the first one's backward must fail loudly instead of pairing the wrong saved activations."""
    case = CASES["cfg1_box"]
    cand = build_candidate(case)
    s = synthetic_samples(**case["inputs"])
    out_a = cand(s)
    with torch.no_grad():
        cand(s)
    with pytest.raises(RuntimeError, match="stale forward"):
        _loss(out_a, case).backward()
    out_b = cand(s)                      # a fresh forward / backward pair still works afterwards
    _loss(out_b, case).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in cand.parameters() if p.requires_grad)


def test_nonfinite_gradient_step_is_zeroed_and_counted(emulated):
    """Overflow sentinel (engine._handover / _finish_guard): with an absurd loss scale the 16-bit activation gradients overflow to inf;
    the flat gradient of that step must come out as exact zeros (a skipped step), the counter must say 1, and the next, normal step
    must be unaffected."""
    case = CASES["cfg1_box"]
    cand = build_candidate(case)
    s = synthetic_samples(**case["inputs"])
    eng = cand.engine()
    eng.grad_scale = 1e30
    _loss(cand(s), case).backward()
    assert eng.overflow_steps() == 1
    for n, p in cand.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and not p.grad.any(), n
    eng.grad_scale = 1024.0
    cand.zero_grad(set_to_none=True)
    _loss(cand(s), case).backward()
    assert eng.overflow_steps() == 1
    g = cand.bbox_embed.layers[0].weight.grad
    assert torch.isfinite(g).all() and g.abs().sum() > 0


def test_changing_requires_grad_after_the_first_forward_raises(emulated):
    case = CASES["cfg1_box"]
    cand = build_candidate(case)
    s = synthetic_samples(**case["inputs"])
    cand(s)
    for p in cand.lang_backbone.parameters():
        p.requires_grad_(False)
    with pytest.raises(RuntimeError, match="requires_grad"):
        cand(s)
    cand.reset_engine()
    _loss(cand(s), case).backward()
    assert all(p.grad is None for p in cand.lang_backbone.parameters())
    assert cand.bbox_embed.layers[0].weight.grad is not None


@pytest.mark.parametrize("name", ["cfg1_box", "seg", "r101_box"])
def test_split_backward_plan_partitions_the_flat_gradient(name):
    """The data-parallel backward (engine._run_backward_split) exchanges the flat gradient buffer part by part; the plan's slices must
    tile [0, n_grad) exactly once, each slice holding only parameters that the part completes."""
    case = CASES[name]
    if case["seg"] and not _have_seg():
        pytest.skip("no seg head")
    eng = build_candidate(case).engine()
    plan = eng._split_plan()
    assert [p for p, _ in plan][0] == "heads" and sum(p.startswith("bert") for p, _ in plan) == 2
    slices = sorted(sl for _, sls in plan for sl in sls)
    assert slices[0][0] == 0 and slices[-1][1] == eng.n_grad
    assert all(a[1] == b[0] for a, b in zip(slices, slices[1:]))
    owner = {}
    for part, sls in plan:
        for lo, hi in sls:
            for n, _ in eng.named:
                if lo <= eng.slots[n][0] < hi:
                    owner[n] = part
    assert len(owner) == len(eng.named)
    for n, part in owner.items():
        if n.startswith("lang_backbone."):
            assert part.startswith("bert:"), n
            hi, lo = (int(v) for v in part.split(":")[1:])
            if ".encoder.layer." in n:
                assert lo <= int(n.split(".encoder.layer.")[1].split(".")[0]) < hi, (n, part)
            else:
                assert (".pooler." in n) == (lo != 0), (n, part)
        elif n.startswith("img_backbone."):
            assert part == "bb:" + n.split("layer")[1][0], (n, part)
        elif n.startswith("input_proj."):
            assert part.startswith("bb:"), n
        else:
            assert part == "heads", (n, part)


def test_loss_scale_backs_off_after_overflow_and_grows_when_clean(emulated):
    """Lazy dynamic loss scaling (engine._maybe_adjust_scale): the overflow counter is read every SCALE_CHECK_EVERY backward passes;
    new overflows shrink the scale (so training recovers without user action), a long clean stretch grows it back."""
    case = CASES["cfg1_box"]
    cand = build_candidate(case)
    s = synthetic_samples(**case["inputs"])
    eng = cand.engine()
    eng.SCALE_CHECK_EVERY, eng.SCALE_GROWTH_INTERVAL = 2, 4
    eng.grad_scale = 2.0 ** 100        # every step overflows until the scale has come down far enough
    scales = []
    for _ in range(12):
        cand.zero_grad(set_to_none=True)
        _loss(cand(s), case).backward()
        scales.append(eng.grad_scale)
    assert scales[-1] < scales[0] and eng.overflow_steps() >= 2
    eng.grad_scale, n0 = 1024.0, eng.overflow_steps()
    eng._scale_seen = n0
    for _ in range(6):
        cand.zero_grad(set_to_none=True)
        _loss(cand(s), case).backward()
    assert eng.overflow_steps() == n0 and eng.grad_scale == 2048.0
    assert torch.isfinite(cand.bbox_embed.layers[0].weight.grad).all()


@pytest.mark.parametrize("cut", [1, 2])
def test_bert_backward_by_layer_ranges_equals_the_whole(cut, emulated_exact):
    """Under data parallelism BERT's backward runs as TWO graphs (upper / lower encoder layers, engine._split_plan): calling
    ``bert.backward(layers=(hi, lo))`` range by range must leave exactly the gradients of one whole call -- the running gradient lives in
    two alternating buffers, so an odd and an even cut are both checked (3-layer BERT)."""
    case = dict(CASES["cfg1_box"], bert_layers=3)
    torch.set_num_threads(os.cpu_count())
    cand = build_candidate(case)
    s = synthetic_samples(**case["inputs"])
    _linear_loss(cand(s)).backward()          # runs the forward (saved activations) and allocates the flat gradient buffer
    eng = cand.engine()
    bert = eng.bert
    assert bert is not None and len(bert.layers) == 3
    B, L = case["inputs"]["B"], case["inputs"]["L"]
    g = torch.Generator().manual_seed(7)
    d_seq = torch.randn(B * L, bert.D, generator=g)
    d_pooled = torch.randn(B, bert.D, generator=g)
    b0, b1 = eng._bert_slice

    def run(ranges):
        eng.gflat.zero_()
        for r in ranges:
            bert.backward("s", d_seq, d_pooled, layers=r)
        return eng.gflat[b0:b1].clone()

    whole = run([None])
    parts = run([(3, cut), (cut, 0)])
    assert whole.abs().max().item() > 0
    assert torch.equal(parts, whole) or (parts - whole).abs().max().item() <= 1e-6 * whole.abs().max().item()
