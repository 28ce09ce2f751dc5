"""Fused clip + AdamW on flat buffers (reftr_b200/optim.py) against the reference's own calls:
torch.optim.AdamW(param_dicts) + torch.nn.utils.clip_grad_norm_ (main_vg.py:234-268, engine_vg.py:62-67).  CPU part: the host
logic (flattening, group segments, zero-copy gradient detection, deferred clipping, LR schedules, state_dict) on the torch
emulation of the two kernels; the kernels themselves are compared on the GPU in test_optim_gpu below."""
import copy

import pytest
import torch
from torch import nn

import emu_ops


def _toy():
    torch.manual_seed(3)
    return nn.Sequential(nn.Linear(20, 33), nn.LayerNorm(33), nn.Linear(33, 7), nn.Linear(7, 5))


def _groups(model, lrs=(1e-2, 3e-3, 1e-3)):
    ps = dict(model.named_parameters())
    names = list(ps)
    g0 = [ps[n] for n in names if n.startswith("0.") or n.startswith("3.")]  # interleaved with the others in named order
    g1 = [ps[n] for n in names if n.startswith("1.")]
    g2 = [ps[n] for n in names if n.startswith("2.")]
    return [{"params": g0, "lr": lrs[0]}, {"params": g1, "lr": lrs[1], "weight_decay": 0.0}, {"params": g2, "lr": lrs[2]}]


def _run(model, opt, clip_fn, steps, sched=None, max_norm=0.1, device="cpu"):
    norms = []
    for i in range(steps):
        g = torch.Generator().manual_seed(100 + i)
        x = torch.randn(16, 20, generator=g).to(device)
        opt.zero_grad()
        (model(x) ** 2).mean().mul(50.0).backward()
        norms.append(float(clip_fn(model.parameters(), max_norm)))
        opt.step()
        if sched is not None:
            sched.step()
    return norms


def _check(device, monkeypatch=None):
    import reftr_b200.optim as ro
    if monkeypatch is not None:
        monkeypatch.setattr(ro, "ops", emu_ops)
    ref = _toy().to(device)
    cand = copy.deepcopy(ref)
    o_ref = torch.optim.AdamW(_groups(ref), lr=1e-2, weight_decay=1e-2)
    o_cand = ro.FusedAdamW(_groups(cand), lr=1e-2, weight_decay=1e-2)
    assert len(o_cand.seg_end) >= 3  # the groups interleave in parameter order
    s_ref = torch.optim.lr_scheduler.StepLR(o_ref, 2, gamma=0.5)
    s_cand = torch.optim.lr_scheduler.StepLR(o_cand, 2, gamma=0.5)
    n_ref = _run(ref, o_ref, torch.nn.utils.clip_grad_norm_, 5, s_ref, device=device)
    n_cand = _run(cand, o_cand, ro.clip_grad_norm_, 5, s_cand, device=device)
    for a, b in zip(n_ref, n_cand):
        assert abs(a - b) < 1e-4 * max(1.0, abs(a))
    assert n_ref[0] > 0.1  # the clip is active (max_norm 0.1)
    for (n, a), (_, b) in zip(ref.named_parameters(), cand.named_parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7), n
    # torch-compatible state: per-parameter exp_avg / exp_avg_sq / step, and a state_dict that round-trips
    sd = o_cand.state_dict()
    assert len(sd["state"]) == len(list(cand.parameters()))
    p0 = next(cand.parameters())
    assert torch.allclose(o_cand.state[p0]["exp_avg"], o_ref.state[next(ref.parameters())]["exp_avg"], rtol=1e-4, atol=1e-7)
    cand2 = copy.deepcopy(ref)
    o2 = ro.FusedAdamW(_groups(cand2), lr=1e-2, weight_decay=1e-2)
    o2.load_state_dict(copy.deepcopy(o_ref.state_dict()))
    _run(ref, o_ref, torch.nn.utils.clip_grad_norm_, 2, device=device)
    _run(cand2, o2, ro.clip_grad_norm_, 2, device=device)
    for (n, a), (_, b) in zip(ref.named_parameters(), cand2.named_parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7), n
    return ro


def test_fused_adamw_host_logic_matches_torch(monkeypatch):
    _check("cpu", monkeypatch)


def test_fused_adamw_consumes_flat_gradients_without_copy(monkeypatch):
    """Gradients that already are views of one flat buffer at the optimizer's offsets (what HotPathFunction hands autograd) are
    used in place; anything else is gathered."""
    import reftr_b200.optim as ro
    monkeypatch.setattr(ro, "ops", emu_ops)
    m = _toy()
    opt = ro.FusedAdamW(_groups(m), lr=1e-2)
    flat = torch.randn(opt.n + 64)
    for _, p in opt._plist:
        o = opt._off[id(p)]
        p.grad = flat[32:][o:o + p.numel()].view(p.shape)  # 32-element (128-byte) offset into the storage: still 16-byte aligned
    g = opt._flat_grads()
    assert g.data_ptr() == flat.data_ptr() + 32 * 4 and opt.flat_g is None
    next(iter(m.parameters())).grad = torch.randn_like(next(iter(m.parameters())))  # one gradient elsewhere -> gather
    g = opt._flat_grads()
    assert opt.flat_g is not None and g.data_ptr() == opt.flat_g.data_ptr()


@pytest.mark.gpu
def test_fused_adamw_kernels_match_torch_gpu():
    _check("cuda")


@pytest.mark.gpu
def test_fused_adamw_on_the_engine_flat_gradients_gpu():
    """cfg1 model, train mode off: three optimisation steps with FusedAdamW on the engine's gradient layout (zero-copy) against
    torch.optim.AdamW + clip_grad_norm_ on an identical model; the 4 LR groups of main_vg.py:234-262."""
    import reftr_b200.optim as ro
    from oracle.cases import CASES
    from oracle.reftr_oracle import total_box_loss
    from reftr_b200.synthetic import synthetic_samples, synthetic_targets
    from util_build import build_candidate
    case = CASES["cfg1_box"]
    s = synthetic_samples(**case["inputs"], device="cuda")
    tgt = synthetic_targets(case["inputs"]["B"], 1, device="cuda")

    def groups(model):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        bb = [p for n, p in named if "img_backbone.0" in n]
        bert = [p for n, p in named if "lang_backbone" in n]
        rest = [p for n, p in named if "img_backbone.0" not in n and "lang_backbone" not in n]
        return [{"params": rest, "lr": 1e-4}, {"params": bb, "lr": 1e-5}, {"params": bert, "lr": 1e-5}]

    # the torch optimizer runs on a twin set of parameters that receives the SAME gradients every step (two engine runs differ in the
    # order of their fp32 atomics, and Adam's g / |g| normalisation turns that noise into +-lr on near-cancelling elements)
    model = build_candidate(case, device="cuda")
    twin = {n: torch.nn.Parameter(p.detach().clone()) for n, p in model.named_parameters() if p.requires_grad}

    class _Twin:
        def named_parameters(self):
            return list(twin.items())

    init = {n: p.detach().clone() for n, p in twin.items()}
    opt = ro.FusedAdamW.for_model(model, groups(model), lr=1e-4, weight_decay=1e-4)
    opt_t = torch.optim.AdamW(groups(_Twin()), lr=1e-4, weight_decay=1e-4)
    for _ in range(3):
        opt.zero_grad()
        total_box_loss(model(s), tgt).backward()
        for n, p in model.named_parameters():
            if p.requires_grad:
                twin[n].grad = p.grad.detach().clone()
        n_f = ro.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 0.1)
        assert opt.flat_g is None  # the engine's flat gradient buffer was consumed in place
        n_t = torch.nn.utils.clip_grad_norm_(list(twin.values()), 0.1)
        assert abs(float(n_f) - float(n_t)) < 1e-4 * float(n_t)
        opt.step()
        opt_t.step()
    torch.cuda.synchronize()
    # the engine's packed 16-bit weight copies followed the updates (re-packed from a captured graph from the 2nd update on): the
    # updated model computes what a freshly built model with the same state_dict computes
    assert model.engine()._repack_graphs
    fresh = build_candidate(case, device="cuda")
    fresh.load_state_dict(model.state_dict())
    with torch.no_grad():
        assert torch.equal(model(s)["pred_boxes"], fresh(s)["pred_boxes"])
    moved = 0
    for n, p in model.named_parameters():
        if p.requires_grad:
            assert torch.allclose(p, twin[n], rtol=1e-5, atol=1e-8), (n, (p - twin[n]).abs().max().item())
            moved += int((p - init[n]).abs().max().item() > 0)
    assert moved > 150, moved


def test_fused_adamw_on_the_engine_layout_cpu(monkeypatch):
    """Host logic of the engine <-> optimizer hand-over without a GPU (kernels emulated in torch): the optimizer adopts the engine's
    gradient-slot layout, the gradients HotPathFunction returns are consumed in place (no gather), the language backbone's slots are
    one contiguous slice (what the overlapped all-reduce relies on), and two steps match torch.optim.AdamW + clip_grad_norm_."""
    import reftr_b200.bert as rbert
    import reftr_b200.engine as rengine
    import reftr_b200.optim as ro
    import reftr_b200.pack as rpack
    from oracle.cases import CASES
    from reftr_b200.synthetic import synthetic_samples
    from util_build import build_candidate
    for m in (rengine, rpack, rbert, ro):
        monkeypatch.setattr(m, "ops", emu_ops)
    case = CASES["cfg1_box"]
    model = build_candidate(case)
    eng = model.engine()
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    b0, b1 = eng._bert_slice
    for n, _ in named:
        off = eng.slots[n][0]
        assert (b0 <= off < b1) == n.startswith("lang_backbone."), n
    groups = [{"params": [p for n, p in named if "lang_backbone" not in n], "lr": 1e-3},
              {"params": [p for n, p in named if "lang_backbone" in n], "lr": 1e-4}]
    opt = ro.FusedAdamW.for_model(model, groups, lr=1e-3, weight_decay=1e-2)
    assert all(opt._off[id(p)] == eng.slots[n][0] for n, p in named) and opt.n == eng.n_grad
    twin = {n: torch.nn.Parameter(p.detach().clone()) for n, p in named}
    opt_t = torch.optim.AdamW([{"params": [twin[n] for n, _ in named if "lang_backbone" not in n], "lr": 1e-3},
                               {"params": [twin[n] for n, _ in named if "lang_backbone" in n], "lr": 1e-4}], lr=1e-3, weight_decay=1e-2)
    s = synthetic_samples(**case["inputs"])
    for _ in range(2):
        opt.zero_grad()
        out = model(s)
        (out["pred_boxes"] * torch.linspace(-1, 1, out["pred_boxes"].numel()).view_as(out["pred_boxes"])).sum().mul(30.0).backward()
        for n, p in named:
            twin[n].grad = p.grad.detach().clone()
        n_f = ro.clip_grad_norm_([p for _, p in named], 0.1)
        assert opt.flat_g is None  # zero-copy: the gradients are views of the engine's flat buffer at the optimizer's offsets
        n_t = torch.nn.utils.clip_grad_norm_(list(twin.values()), 0.1)
        assert abs(float(n_f) - float(n_t)) < 1e-4 * float(n_t) and float(n_t) > 0.1
        opt.step()
        opt_t.step()
    for n, p in named:
        assert torch.allclose(p, twin[n], rtol=2e-5, atol=1e-7), n
