"""GPU parity of train-mode dropout (rb_dropout, include/reftr_b200.h): every kernel that draws a mask is compared with plain
PyTorch using the SAME counter-based mask restated in tests/dropout_ref.py -- keep decisions bit-exact, values to the
kernel's usual tolerance -- and the whole model in train mode is compared with the oracle drawing the same masks."""
import os

import pytest
import torch
import torch.nn.functional as F

import dropout_ref
import emu_ops

pytestmark = pytest.mark.gpu
dev = "cuda"
from reftr_b200 import ops as _ops
BF = _ops.t16()
SEED = 0x0123456789ABCDEF


def _rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / (b.float().abs().max().cpu() + 1e-6)).item()


def _drop(name, p=0.1, seed=SEED):
    from reftr_b200 import ops
    sd = torch.full((1,), seed, dtype=torch.int64, device=dev)
    return ops.Drop(sd, name, p)


def _mask(name, rows, cols, p=0.1, seed=SEED):
    return dropout_ref.mask_scale(seed, name, rows, cols, p, device=dev)


@pytest.mark.parametrize("M,N,p", [(300, 256, 0.1), (1000, 2048, 0.1), (129, 768, 0.5), (64, 64, 0.25)])
def test_gemm_epilogue_mask_is_bit_exact(M, N, p):
    """A = 0, bias = 1: the output IS the mask (0 or 1/(1-p)) -> the keep decisions must match the restated generator exactly."""
    from reftr_b200 import ops
    A = torch.zeros(M, 64, device=dev, dtype=BF)
    W = torch.zeros(N, 64, device=dev, dtype=BF)
    bias = torch.ones(N, device=dev)
    out = torch.empty(M, N, device=dev)
    d = _drop("site.a", p)
    ops.gemm(A, W, M, N, 64, bias=bias, out32=out, drop=d)
    ref = _mask("site.a", M, N, p)
    assert torch.equal(out > 0, ref > 0)
    assert (out - ref).abs().max().item() < 1e-6
    keep = (out > 0).float().mean().item()
    assert abs(keep - (1 - d.thr / 65536.0)) < 5 * (p * (1 - p) / (M * N)) ** 0.5 + 1e-4
    # another site / another seed give another mask
    out2 = torch.empty_like(out)
    ops.gemm(A, W, M, N, 64, bias=bias, out32=out2, drop=_drop("site.b", p))
    assert not torch.equal(out2 > 0, out > 0)
    d.seed.fill_(SEED + 1)  # the seed is read from device memory at run time (CUDA-graph replays draw new masks)
    ops.gemm(A, W, M, N, 64, bias=bias, out32=out2, drop=d)
    assert torch.equal(out2 > 0, _mask("site.a", M, N, p, seed=SEED + 1) > 0)


def test_gemm_dropout_semantics():
    """v = res32 + dropout(acc + bias);  v = dropout(relu(acc + bias));  head-group dropout;  mask_src with mask_scale."""
    from reftr_b200 import ops
    M, N, K = 777, 256, 256
    A = torch.randn(M, K, device=dev).to(BF)
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    lin = A.float() @ W.float().t() + bias
    out = torch.empty(M, N, device=dev)
    ops.gemm(A, W, M, N, K, bias=bias, res32=res, out32=out, drop=_drop("g.res"))
    assert _rel(out, res + lin * _mask("g.res", M, N)) < 2e-3
    outb = torch.empty(M, N, device=dev, dtype=BF)
    ops.gemm(A, W, M, N, K, bias=bias, relu=True, out=outb, drop=_drop("g.relu"))
    assert _rel(outb, F.relu(lin) * _mask("g.relu", M, N)) < 1e-2
    ops.gemm(A, W, M, N, K, bias=bias, out=outb, drop=_drop("g.head"), drop_gshift=5)
    hm = _mask("g.head", M, N // 32).repeat_interleave(32, dim=1)
    assert _rel(outb, lin * hm) < 1e-2
    src = torch.randn(M, N, device=dev).to(BF)
    ops.gemm(A, W, M, N, K, mask_src=src, out=outb, mask_scale=1.25)
    assert _rel(outb, torch.where(src.float() > 0, (lin - bias) * 1.25, torch.zeros((), device=dev))) < 1e-2


def test_layernorm_dropout_fwd_bwd():
    from reftr_b200 import ops
    rows, D = 333, 256
    x = torch.randn(rows, D, device=dev) * 2 + 0.3
    g = 1 + 0.1 * torch.randn(D, device=dev)
    b = 0.1 * torch.randn(D, device=dev)
    d = _drop("ln.relu")
    y32 = torch.empty(rows, D, device=dev); yb = torch.empty(rows, D, device=dev, dtype=BF)
    mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
    ops.layernorm_fwd(x, g, b, rows, y32=y32, yb=yb, relu=True, mean=mean, rstd=rstd, drop=d)
    xr = x.clone().requires_grad_()
    m = _mask("ln.relu", rows, D)
    ref = F.relu(F.layer_norm(xr, (D,), g, b)) * m
    assert _rel(y32, ref) < 1e-5 and _rel(yb, ref) < 8e-3
    dy = torch.randn(rows, D, device=dev)
    ref.backward(dy)
    dx = torch.empty(rows, D, device=dev); dxb = torch.empty(rows, D, device=dev, dtype=BF)
    od = _drop("ln.out", 0.2)
    ops.layernorm_bwd(dy, x, g, mean, rstd, rows, y_relu=y32, relu_scale=d.scale, dx32=dx, dxb=dxb, dxb_drop=od)
    assert _rel(dx, xr.grad) < 1e-4
    assert _rel(dxb, xr.grad * _mask("ln.out", rows, D, 0.2)) < 8e-3


def test_ln_wide_dropout_fwd_bwd():
    from reftr_b200 import ops
    rows, D = 100, 768
    x = torch.randn(rows, D, device=dev)
    g = 1 + 0.1 * torch.randn(D, device=dev)
    b = 0.1 * torch.randn(D, device=dev)
    y32 = torch.empty(rows, D, device=dev); yb = torch.empty(rows, D, device=dev, dtype=BF)
    mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
    ops.ln_wide_fwd(x, g, b, rows, y32=y32, yb=yb, mean=mean, rstd=rstd, eps=1e-12, drop=_drop("w.emb"))
    xr = x.clone().requires_grad_()
    ref = F.layer_norm(xr, (D,), g, b, eps=1e-12) * _mask("w.emb", rows, D)
    assert _rel(y32, ref) < 1e-5 and _rel(yb, ref) < 8e-3
    dy = torch.randn(rows, D, device=dev)
    ref.backward(dy)
    dx = torch.empty(rows, D, device=dev); dxb = torch.empty(rows, D, device=dev, dtype=BF)
    ops.ln_wide_bwd(dy, x, g, mean, rstd, rows, dx32=dx, dxb=dxb, dy_drop=_drop("w.emb"), dxb_drop=_drop("w.out"))
    assert _rel(dx, xr.grad) < 1e-4
    assert _rel(dxb, xr.grad * _mask("w.out", rows, D)) < 8e-3


def _attn_ref(q, k, v, kpm, H, scale, mask):
    B, Tq, d = q.shape
    Sk = k.shape[1]
    dh = d // H
    qh = (q * scale).view(B, Tq, H, dh).transpose(1, 2)
    kh = k.view(B, Sk, H, dh).transpose(1, 2)
    vh = v.view(B, Sk, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if kpm is not None:
        s = s.masked_fill(kpm.bool()[:, None, None, :], float("-inf"))
    p = s.softmax(-1) * mask.view(B, H, Tq, Sk)
    return (p @ vh).transpose(1, 2).reshape(B, Tq, d)


# tcgen05 kernels (Tq >= 64), odd and even key counts, SIMT kernels (short query counts)
@pytest.mark.parametrize("B,H,Tq,Sk", [(2, 8, 420, 420), (2, 8, 665, 665), (1, 8, 300, 57), (2, 8, 64, 64), (3, 8, 1, 77), (2, 8, 5, 5), (2, 8, 3, 420),
                                        (2, 8, 16, 490)])
def test_attention_dropout_fwd_bwd(B, H, Tq, Sk):
    from reftr_b200 import ops
    d = H * 32
    scale = 32 ** -0.5
    q = torch.randn(B, Tq, d, device=dev).to(BF); k = torch.randn(B, Sk, d, device=dev).to(BF); v = torch.randn(B, Sk, d, device=dev).to(BF)
    kpm = torch.zeros(B, Sk, dtype=torch.uint8, device=dev)
    kpm[:, Sk - Sk // 4:] = 1
    kpm[:, 0] = 0
    dr = _drop("attn.p")
    o = torch.empty(B * Tq, d, device=dev, dtype=BF)
    lse = torch.empty(B, H, Tq, device=dev)
    ops.attn_fwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), kpm, o, lse, B, H, Tq, Sk, scale, drop=dr)
    qr, kr, vr = (t.float().requires_grad_() for t in (q, k, v))
    m = _mask("attn.p", B * H * Tq, Sk)
    ref = _attn_ref(qr, kr, vr, kpm, H, scale, m)
    assert _rel(o.view(B, Tq, d), ref) < 1e-2
    # without dropout the same call gives a different result (the mask really is applied)
    o0 = torch.empty_like(o)
    ops.attn_fwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), kpm, o0, lse, B, H, Tq, Sk, scale)
    assert _rel(o0.view(B, Tq, d), ref) > 2e-2
    do = torch.randn(B, Tq, d, device=dev).to(BF)
    ref.backward(do.float())
    dq = torch.empty_like(q).view(-1, d); dk = torch.empty_like(k).view(-1, d); dv = torch.empty_like(v).view(-1, d)
    Dbuf = torch.empty(B, H, Tq, device=dev)
    ops.attn_bwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), kpm, o, do.view(-1, d), lse, dq, dk, dv, Dbuf, B, H, Tq, Sk, scale, drop=dr)
    assert _rel(dq.view(B, Tq, d), qr.grad) < 2e-2
    assert _rel(dk.view(B, Sk, d), kr.grad) < 2e-2
    assert _rel(dv.view(B, Sk, d), vr.grad) < 2e-2


@pytest.mark.parametrize("B,S", [(3, 20), (2, 22), (2, 90)])
def test_attn_small_dropout(B, S):
    from reftr_b200 import ops
    H, D = 12, 768
    qkv = (torch.randn(B * S, 3 * D) * 0.5).to(BF)
    mask = torch.zeros(B, S, dtype=torch.uint8)
    mask[0, S - 3:] = 1
    sd = torch.full((1,), SEED, dtype=torch.int64)
    d_c, d_g = emu_ops.Drop(sd, "b.attn", 0.1), _drop("b.attn")
    o_c, o_g = torch.empty(B * S, D, dtype=BF), torch.empty(B * S, D, dtype=BF, device=dev)
    P_c, P_g = torch.empty(B, H, S, S), torch.empty(B, H, S, S, device=dev)
    emu_ops.attn_small_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], mask, o_c, P_c, B, H, S, 0.125, drop=d_c)
    qg = qkv.to(dev)
    ops.attn_small_fwd(qg[:, :D], qg[:, D:2 * D], qg[:, 2 * D:], mask.to(dev), o_g, P_g, B, H, S, 0.125, drop=d_g)
    assert _rel(P_g, P_c) < 1e-4 and _rel(o_g, o_c) < 1e-2
    do = torch.randn(B * S, D).to(BF)
    g_c, g_g = torch.empty(B * S, 3 * D, dtype=BF), torch.empty(B * S, 3 * D, dtype=BF, device=dev)
    emu_ops.attn_small_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], do, P_c, g_c[:, :D], g_c[:, D:2 * D], g_c[:, 2 * D:], B, H, S, 0.125, drop=d_c)
    ops.attn_small_bwd(qg[:, :D], qg[:, D:2 * D], qg[:, 2 * D:], do.to(dev), P_g, g_g[:, :D], g_g[:, D:2 * D], g_g[:, 2 * D:], B, H, S, 0.125, drop=d_g)
    assert _rel(g_g, g_c) < 1.5e-2


@pytest.mark.parametrize("name", ["cfg1_box", "multi_phrase"])
def test_e2e_train_mode_matches_oracle_with_same_masks(name):
    """model.train(): eager step, CUDA-graph capture step and a replay, each under a NEW seed; the oracle (fp32, CPU, train mode)
    draws the same masks through tests/dropout_ref.py.  Tolerances are those of the eval-mode parity test."""
    import oracle.reftr_oracle as orc
    from oracle.cases import CASES, build_oracle
    from oracle.reftr_oracle import total_box_loss
    from reftr_b200.synthetic import synthetic_samples, synthetic_targets
    from util_build import build_candidate, compare_grads, rel_l2
    case = dict(CASES[name])
    case["oracle_kw"] = dict(case["oracle_kw"], dropout=0.1)
    torch.set_num_threads(os.cpu_count())
    n_ph = max(case["inputs"].get("n_ph", 0), 1)
    s_cpu = synthetic_samples(**case["inputs"])
    s = synthetic_samples(**case["inputs"], device=dev)
    cand = build_candidate(case, device=dev).train()
    eng = cand.engine()
    prev = None

    def smooth_loss(out, device):
        # a fixed random linear functional of every layer's boxes: L1 + GIoU has a sign() in its gradient, so a coordinate that
        # lands within the forward tolerance of its target flips a whole gradient component -- which seed does that is luck
        g = torch.Generator().manual_seed(5)
        boxes = [out["pred_boxes"]] + [a["pred_boxes"] for a in out.get("aux_outputs", [])]
        return sum((b * torch.randn(b.shape, generator=g).to(device)).sum() for b in boxes)

    for step in range(3):
        seed = (SEED + 7919 * step) & 0x7FFFFFFFFFFFFFFF
        oracle = build_oracle(case).train()
        hook = dropout_ref.OracleDropoutHook(seed)
        undo = dropout_ref.hook_hf_bert(oracle.lang_backbone, hook, orc.DROP_CTX)
        orc.DROPOUT_HOOK = hook
        try:
            out_o = oracle(s_cpu)
            smooth_loss(out_o, "cpu").backward()
        finally:
            orc.DROPOUT_HOOK = None
            undo()
        cand.zero_grad(set_to_none=True)
        eng.next_seed = seed
        out_c = cand(s)
        smooth_loss(out_c, dev).backward()
        torch.cuda.synchronize()
        assert eng.train_mode and eng.last_seed == seed
        assert set(hook.seen) == set(eng._drops)
        err = (out_c["pred_boxes"].cpu() - out_o["pred_boxes"]).abs().max().item()
        print(name, "train step", step, "pred_boxes max abs err", err, "rel-L2", rel_l2(out_c["pred_boxes"], out_o["pred_boxes"]))
        assert err < 4e-3 and rel_l2(out_c["pred_boxes"], out_o["pred_boxes"]) < 2e-3  # measured 0.7e-3..1.5e-3 / 0.8e-3..1.1e-3
        for a, b in zip(out_c["aux_outputs"], out_o["aux_outputs"]):
            assert (a["pred_boxes"].cpu() - b["pred_boxes"]).abs().max().item() < 4e-3
        if prev is not None:  # a new seed gives a new output, also from a replayed graph
            assert (out_c["pred_boxes"] - prev).abs().max().item() > 1e-4
        prev = out_c["pred_boxes"].detach().clone()
        errs = compare_grads(cand, oracle)
        norms = {n: p.grad.norm().item() for n, p in oracle.named_parameters() if p.grad is not None}
        big = max(norms.values())
        live = {n: e for n, e in errs.items() if norms[n] > 1e-6 * big}
        print(name, "train step", step, "worst grads", sorted(live.items(), key=lambda kv: -kv[1])[:6])
        bad = {n: e for n, e in live.items() if e != e or e > 0.9}
        assert not bad, bad
        assert sorted(live.values())[len(live) // 2] < 0.15
    if eng.use_graphs:
        assert any(st["fwd"] is not None and st["bwd"] is not None for st in eng._states.values())
    # eval mode afterwards: dropout off again, deterministic
    cand.eval()
    a = cand(s)["pred_boxes"].clone()
    b = cand(s)["pred_boxes"].clone()
    assert torch.equal(a, b)
