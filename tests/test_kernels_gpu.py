"""GPU parity of the non-GEMM kernels (through the C ABI) against plain PyTorch fp32 references of the same ops."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
from reftr_b200 import ops as _ops
T16 = _ops.t16()  # the library's 16-bit operand type (IEEE half by default)
dev = "cuda"


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6)).item()


def test_layernorm_fwd_bwd():
    from reftr_b200 import ops
    rows, D = 777, 256
    x = torch.randn(rows, D, device=dev) * 2 + 0.3
    g = 1 + 0.1 * torch.randn(D, device=dev)
    b = 0.1 * torch.randn(D, device=dev)
    pos = torch.randn(rows, D, device=dev)
    for relu in (False, True):
        y32 = torch.empty(rows, D, device=dev); yb = torch.empty(rows, D, device=dev, dtype=T16)
        ypb = torch.empty_like(yb); mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
        ops.layernorm_fwd(x, g, b, rows, y32=y32, yb=yb, pos32=pos, ypb=ypb, relu=relu, mean=mean, rstd=rstd)
        xr = x.clone().requires_grad_(); gr = g.clone().requires_grad_(); br = b.clone().requires_grad_()
        ref = F.layer_norm(xr, (D,), gr, br)
        if relu:
            ref = F.relu(ref)
        assert _rel(y32, ref) < 1e-5
        assert _rel(yb, ref) < 8e-3 and _rel(ypb, ref + pos) < 8e-3
        dy = torch.randn(rows, D, device=dev)
        ref.backward(dy)
        dx = torch.empty(rows, D, device=dev); dxb = torch.empty(rows, D, device=dev, dtype=T16)
        dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
        ops.layernorm_bwd(dy, x, g, mean, rstd, rows, y_relu=y32 if relu else None, dx32=dx, dxb=dxb, dgamma=dg, dbeta=db)
        assert _rel(dx, xr.grad) < 1e-4 and _rel(dg, gr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4
        assert _rel(dxb, xr.grad) < 8e-3


def test_layernorm_rowmap():
    from reftr_b200 import ops
    B, L, S, D = 3, 5, 11, 256
    x = torch.randn(B * L, D, device=dev)
    g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
    y = torch.zeros(B * S, D, device=dev)
    ops.layernorm_fwd(x, g, b, B * L, y32=y, rowmap=(L, S, 0))
    ref = F.layer_norm(x, (D,))
    assert _rel(y.view(B, S, D)[:, :L].reshape(B * L, D), ref) < 1e-5
    assert y.view(B, S, D)[:, L:].abs().max().item() == 0


def test_groupnorm_tokens():
    from reftr_b200 import ops
    B, h, w, L = 2, 5, 7, 3
    S = L + h * w
    x = torch.randn(B, 256, h, w, device=dev) * 1.5 + 0.2
    g = 1 + 0.1 * torch.randn(256, device=dev); be = 0.1 * torch.randn(256, device=dev)
    xp = F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous()
    pos = torch.randn(B * S, 256, device=dev)
    y32 = torch.zeros(B * S, 256, device=dev); yb = torch.zeros(B * S, 256, device=dev, dtype=T16); ypb = torch.zeros_like(yb)
    mean = torch.empty(B * 32, device=dev); rstd = torch.empty(B * 32, device=dev)
    ops.groupnorm_tokens_fwd(xp, g, be, B, h, w, S, L, y32, yb, pos, ypb, mean, rstd)
    xr = x.clone().requires_grad_(); gr = g.clone().requires_grad_(); br = be.clone().requires_grad_()
    ref = F.group_norm(xr, 32, gr, br)
    ref_tok = ref.flatten(2).transpose(1, 2)  # [B, hw, 256]
    got = y32.view(B, S, 256)[:, L:]
    assert _rel(got, ref_tok) < 1e-4
    assert _rel(ypb.view(B, S, 256)[:, L:], ref_tok + pos.view(B, S, 256)[:, L:]) < 8e-3
    dy = torch.randn(B * S, 256, device=dev)
    ref_tok.backward(dy.view(B, S, 256)[:, L:])
    dx = torch.zeros(B, h + 2, w + 2, 256, device=dev, dtype=T16)
    dg = torch.zeros(256, device=dev); db = torch.zeros(256, device=dev)
    ops.groupnorm_tokens_bwd(dy, None, xp, g, mean, rstd, B, h, w, S, L, dx, dg, db)
    assert _rel(dx[:, 1:-1, 1:-1].permute(0, 3, 1, 2), xr.grad) < 1e-2
    assert _rel(dg, gr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4


def _attn_ref(q, k, v, kpm, H, scale):
    B, Tq, d = q.shape
    Sk = k.shape[1]
    dh = d // H
    qh = (q.float() * scale).view(B, Tq, H, dh).transpose(1, 2)
    kh = k.float().view(B, Sk, H, dh).transpose(1, 2)
    vh = v.float().view(B, Sk, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if kpm is not None:
        s = s.masked_fill(kpm.bool()[:, None, None, :], float("-inf"))
    p = s.softmax(-1)
    return (p @ vh).transpose(1, 2).reshape(B, Tq, d)


@pytest.mark.parametrize("B,H,Tq,Sk", [(2, 8, 57, 57), (1, 8, 420, 420), (3, 8, 1, 77), (2, 8, 5, 5), (1, 8, 300, 665), (2, 8, 3, 420),
                                        (2, 8, 64, 64), (3, 8, 128, 200), (2, 8, 420, 420), (2, 8, 665, 665), (2, 8, 100, 7)])
def test_attention_fwd_bwd(B, H, Tq, Sk):
    from reftr_b200 import ops
    d = H * 32
    scale = 32 ** -0.5
    q = torch.randn(B, Tq, d, device=dev).to(T16); k = torch.randn(B, Sk, d, device=dev).to(T16); v = torch.randn(B, Sk, d, device=dev).to(T16)
    kpm = torch.zeros(B, Sk, dtype=torch.uint8, device=dev)
    kpm[:, Sk - Sk // 4:] = 1
    kpm[0, 1::3] = 1
    kpm[:, 0] = 0
    o = torch.empty(B * Tq, d, device=dev, dtype=T16)
    lse = torch.empty(B, H, Tq, device=dev)
    ops.attn_fwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), kpm, o, lse, B, H, Tq, Sk, scale)
    qr, kr, vr = (t.float().requires_grad_() for t in (q, k, v))
    ref = _attn_ref(qr, kr, vr, kpm, H, scale)
    assert _rel(o.view(B, Tq, d), ref) < 1e-2
    do = torch.randn(B, Tq, d, device=dev).to(T16)
    ref.backward(do.float())
    dq = torch.empty_like(q).view(-1, d); dk = torch.empty_like(k).view(-1, d); dv = torch.empty_like(v).view(-1, d)
    Dbuf = torch.empty(B, H, Tq, device=dev)
    ops.attn_bwd(q.view(-1, d), k.view(-1, d), v.view(-1, d), kpm, o, do.view(-1, d), lse, dq, dk, dv, Dbuf, B, H, Tq, Sk, scale)
    assert _rel(dq.view(B, Tq, d), qr.grad) < 2e-2
    assert _rel(dk.view(B, Sk, d), kr.grad) < 2e-2
    assert _rel(dv.view(B, Sk, d), vr.grad) < 2e-2


def test_qenc_pool():
    from reftr_b200 import ops
    B, L, n_ph = 3, 9, 2
    k = torch.randn(B, 256, device=dev) * 0.2; q = torch.randn(B * L, 256, device=dev) * 0.2; v = torch.randn(B * L, 256, device=dev)
    mask = torch.zeros(B, n_ph, L, dtype=torch.uint8, device=dev)
    mask[:, :, 0] = 1; mask[:, 1, 5:] = 1
    att = torch.empty(B, n_ph, L, device=dev); c = torch.empty(B * n_ph, 256, device=dev)
    ops.qenc_pool_fwd(k, q, v, mask, B, L, n_ph, att, c)
    kr, qr, vr = k.clone().requires_grad_(), q.clone().requires_grad_(), v.clone().requires_grad_()
    s = torch.bmm(kr.view(B, 1, 256), qr.view(B, L, 256).transpose(1, 2)).expand(-1, n_ph, -1).masked_fill(mask.bool(), float("-inf"))
    a = s.softmax(-1)
    ref = (vr.view(B, 1, L, 256) * a.unsqueeze(-1)).sum(-2)
    assert _rel(c.view(B, n_ph, 256), ref) < 1e-4
    dc = torch.randn(B * n_ph, 256, device=dev)
    ref.backward(dc.view(B, n_ph, 256))
    dk = torch.empty_like(k); dq = torch.empty_like(q); dv = torch.empty_like(v)
    ops.qenc_pool_bwd(dc, k, q, v, att, B, L, n_ph, dk, dq, dv)
    assert _rel(dk, kr.grad) < 1e-3 and _rel(dq, qr.grad) < 1e-3 and _rel(dv, vr.grad) < 1e-3


def test_pos_and_mask():
    from oracle.reftr_oracle import sine_position_embedding
    from reftr_b200 import ops
    B, H, W, h, w, L = 2, 96, 128, 3, 4, 5
    img_mask = torch.zeros(B, H, W, dtype=torch.bool, device=dev)
    img_mask[1, :, 90:] = True
    img_mask[1, 70:, :] = True
    sent_mask = torch.ones(B, L, dtype=torch.int64, device=dev); sent_mask[1, 3:] = 0
    lang_pos = torch.randn(128, 256, device=dev); tt = torch.randn(2, 256, device=dev); lvl = torch.randn(1, 256, device=dev)
    S = L + h * w
    pos = torch.empty(B * S, 256, device=dev); kpm = torch.empty(B, S, dtype=torch.uint8, device=dev)
    ops.build_pos_mask(img_mask, B, H, W, h, w, sent_mask, L, lang_pos, tt, lvl, pos, kpm)
    m_small = F.interpolate(img_mask[None].float(), size=(h, w)).to(torch.bool)[0]
    ref_vis = sine_position_embedding(m_small).flatten(2).transpose(1, 2) + lvl[0] + tt[1]
    ref_lang = (lang_pos[:L] + tt[0]).unsqueeze(0).expand(B, -1, -1)
    ref = torch.cat([ref_lang, ref_vis], 1)
    assert (pos.view(B, S, 256) - ref).abs().max().item() < 2e-5
    ref_kpm = torch.cat([sent_mask == 0, m_small.flatten(1)], 1)
    assert torch.equal(kpm.bool(), ref_kpm)
    dpos = torch.randn(B * S, 256, device=dev)
    dl = torch.zeros(128, 256, device=dev); dt = torch.zeros(2, 256, device=dev); dv = torch.zeros(1, 256, device=dev)
    ops.embed_grad(dpos, B, S, L, dl, dt, dv)
    d3 = dpos.view(B, S, 256)
    assert _rel(dl[:L], d3[:, :L].sum(0)) < 1e-5 and dl[L:].abs().max().item() == 0
    assert _rel(dt[0], d3[:, :L].sum((0, 1))) < 1e-5 and _rel(dt[1], d3[:, L:].sum((0, 1))) < 1e-5 and _rel(dv[0], dt[1]) < 1e-6


def test_stem_maxpool_parity_pack():
    from reftr_b200 import ops
    B, H, W = 2, 70, 90
    img = torch.randn(B, 3, H, W, device=dev)
    w = torch.randn(64, 3, 7, 7, device=dev) * 0.1
    bn = [1 + 0.1 * torch.randn(64, device=dev), 0.1 * torch.randn(64, device=dev), 0.1 * torch.randn(64, device=dev), 0.5 + torch.rand(64, device=dev)]
    H1, W1 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    H2, W2 = (H1 + 2 - 3) // 2 + 1, (W1 + 2 - 3) // 2 + 1
    col = torch.empty(B * H1 * W1, 160, device=dev, dtype=T16)
    ops.stem_im2col(img, col, B, H, W, H1, W1)
    wf = torch.empty(64, 160, device=dev, dtype=T16); sc = torch.empty(64, device=dev); bi = torch.empty(64, device=dev)
    ops.pack_conv(w, bn, None, wf, 160, None, sc, bi)
    c1 = torch.empty(B * H1 * W1, 64, device=dev, dtype=T16)
    ops.gemm(col, wf, B * H1 * W1, 64, 160, bias=bi, relu=True, out=c1)
    scale = bn[0] * (bn[3] + 1e-5).rsqrt()
    ref = F.relu(F.conv2d(img, w, stride=2, padding=3) * scale.view(1, -1, 1, 1) + (bn[1] - bn[2] * scale).view(1, -1, 1, 1))
    assert _rel(c1.view(B, H1, W1, 64).permute(0, 3, 1, 2), ref) < 2e-2
    pooled = torch.empty(B, H2 + 2, W2 + 2, 64, device=dev, dtype=T16)
    ops.maxpool_3x3s2(c1, pooled, B, H1, W1, 64, H2, W2)
    refp = F.max_pool2d(c1.view(B, H1, W1, 64).permute(0, 3, 1, 2).float(), 3, 2, 1)
    assert _rel(pooled[:, 1:-1, 1:-1].permute(0, 3, 1, 2), refp) == 0
    assert pooled[:, 0].abs().max().item() == 0 and pooled[:, :, -1].abs().max().item() == 0
    # parity split / merge round trip on an odd-sized grid
    Hh, Ww, C = H2, W2, 64
    Ho, Wo = (Hh + 1) // 2, (Ww + 1) // 2
    xs = torch.empty(4, B, Ho + 2, Wo + 2, C, device=dev, dtype=T16)
    ops.parity_split(pooled, xs, B, Hh, Ww, C, Ho, Wo)
    for p in range(2):
        for q in range(2):
            sub = pooled[:, p::2, q::2]
            assert torch.equal(xs[2 * p + q, :, :sub.shape[1], :sub.shape[2]], sub)
    back = torch.empty_like(pooled)
    ops.parity_merge(xs, None, None, back, B, Hh, Ww, C, Ho, Wo)
    assert torch.equal(back, pooled)
    # 3x3 pack: dgrad copy is the flipped transpose
    w3 = torch.randn(128, 64, 3, 3, device=dev)
    wf3 = torch.empty(128, 576, device=dev, dtype=T16); wd3 = torch.empty(64, 9 * 128, device=dev, dtype=T16)
    ops.pack_conv(w3, None, None, wf3, 576, wd3, None, None)
    assert torch.equal(wf3.view(128, 3, 3, 64), w3.permute(0, 2, 3, 1).to(T16))
    assert torch.equal(wd3.view(64, 3, 3, 128), w3.flip(2, 3).permute(1, 2, 3, 0).to(T16))
    # linear pack + colsum + cast
    wl = torch.randn(100, 72, device=dev)
    wb = torch.empty(100, 72, device=dev, dtype=T16); wt = torch.empty(72, 100, device=dev, dtype=T16)
    ops.pack_linear(wl, wb, wt)
    assert torch.equal(wb, wl.to(T16)) and torch.equal(wt, wl.t().to(T16))
    x = torch.randn(1000, 72, device=dev)
    cs = torch.zeros(72, device=dev)
    ops.colsum(x, cs)
    assert _rel(cs, x.sum(0)) < 1e-5
    assert torch.equal(ops.cast_bf16(x), x.to(T16))


@pytest.mark.parametrize("rows,N,dt", [(6720, 256, T16), (6720, 768, T16), (107584, 128, T16), (321, 2048, T16),
                                       (6720, 256, torch.float32), (96, 4, torch.float32), (33, 72, torch.float32)])
def test_colsum(rows, N, dt):
    from reftr_b200 import ops
    ld = 64 if N == 4 else N
    x = torch.randn(rows, ld, device=dev).to(dt)
    out = torch.ones(N, device=dev)
    ops.colsum(x, out, rows=rows, N=N)
    ref = 1 + x[:, :N].float().sum(0)
    assert (out - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("padded", [False, True])
def test_fused_box_criterion_matches_torch(padded):
    """rb_box_loss (all layers, L1 + GIoU, values and gradients) against the plain-torch restatement of criterion.py:113-153."""
    from reftr_b200.criterion import CriterionVGMultiPhrase
    nl, B, n_ph, k = 6, 5, 3, 2
    g = torch.Generator().manual_seed(0)
    c = 0.25 + 0.5 * torch.rand(nl, B, n_ph, k, 2, generator=g)
    wh = 0.05 + 0.4 * torch.rand(nl, B, n_ph, k, 2, generator=g)
    boxes = torch.cat([c, wh], -1).to(dev)
    pm = torch.ones(B, n_ph * k, dtype=torch.bool, device=dev)
    if padded:
        pm.view(B, n_ph, k)[1, 2] = False
        pm.view(B, n_ph, k)[3, 1:] = False
    targets = []
    for b in range(B):
        nv = int(pm.view(B, n_ph, k)[b, :, 0].sum().item())
        t = torch.cat([0.3 + 0.4 * torch.rand(nv, 2, generator=g), 0.1 + 0.3 * torch.rand(nv, 2, generator=g)], -1).to(dev)
        targets.append({"boxes": t, "labels": [0] * nv})
    wd = {"loss_bbox": 1.0, "loss_giou": 1.0}
    crit = CriterionVGMultiPhrase(wd, ["boxes"])
    res = {}
    for fused in (False, True):
        bx = boxes.clone().requires_grad_()
        if fused:   # slices of one tensor, as reftr_b200.modules.RefTR.forward returns them: the criterion takes the one-kernel path
            out = {"pred_boxes": bx[-1], "phrase_mask": pm, "aux_outputs": [{"pred_boxes": bx[i], "phrase_mask": pm} for i in range(nl - 1)]}
            assert crit._all_layer_boxes(out) is not None
        else:       # separate tensors per layer (what the reference's model returns): the plain-torch restatement runs
            out = {"pred_boxes": bx[-1] * 1.0, "phrase_mask": pm, "aux_outputs": [{"pred_boxes": bx[i] * 1.0, "phrase_mask": pm} for i in range(nl - 1)]}
            assert crit._all_layer_boxes(out) is None
        ld = crit(out, targets)
        w = torch.linspace(0.5, 1.5, len(ld)).tolist()
        sum(v * wi for v, wi in zip((ld[k_] for k_ in sorted(ld)), w)).backward()
        res[fused] = ({k_: v.item() for k_, v in ld.items()}, bx.grad.clone())
    assert set(res[True][0]) == set(res[False][0]) and len(res[True][0]) == 2 * nl
    for k_, v in res[False][0].items():
        assert abs(res[True][0][k_] - v) < 1e-5 * max(1.0, abs(v)), (k_, res[True][0][k_], v)
    assert (res[True][1] - res[False][1]).abs().max().item() < 1e-5


@pytest.mark.parametrize("B,H,W", [(2, 64, 96), (1, 224, 224), (2, 50, 70)])
def test_stem_conv_fused(B, H, W):
    from reftr_b200 import ops
    H1, W1 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    img = torch.randn(B, 3, H, W, device=dev)
    w = torch.randn(64, 3, 7, 7, device=dev) * 0.1
    bias = torch.randn(64, device=dev) * 0.1
    wf = torch.zeros(64, 160, device=dev, dtype=T16)
    wf[:, :147] = w.permute(0, 2, 3, 1).reshape(64, 147).to(T16)
    out = torch.empty(B * H1 * W1, 64, device=dev, dtype=T16)
    ops.stem_conv(img, wf, bias, out, B, H, W, H1, W1)
    ref = F.relu(F.conv2d(img.to(T16).float(), w.to(T16).float(), bias, stride=2, padding=3)).permute(0, 2, 3, 1).reshape(B * H1 * W1, 64)
    assert _rel(out, ref) < 1e-2


@pytest.mark.parametrize("B,H,W", [(2, 64, 96), (1, 224, 224), (2, 50, 70), (1, 640, 640), (3, 130, 514), (16, 96, 128)])
def test_stem_pool_fused(B, H, W):
    """conv1 + bn1 + relu + maxpool in one pass (rb_stem_pool: TMA-staged HWC4 rows read through overlapping UMMA descriptors, pool in
    shared memory) against PyTorch; padded NHWC output with an exact zero border."""
    from reftr_b200 import ops
    from reftr_b200.pack import PackedStem
    H1, W1 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    H2, W2 = (H1 + 2 - 3) // 2 + 1, (W1 + 2 - 3) // 2 + 1
    torch.manual_seed(B * 1000 + H + W)
    img = torch.randn(B, 3, H, W, device=dev)
    conv = torch.nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False).to(dev)
    with torch.no_grad():
        conv.weight.mul_(3.0)
    st = PackedStem(conv, None, need_dgrad=False, ldk=160)
    st.refresh()
    assert st.wrow.shape == (7, 4, 64, 8)
    hwc4 = torch.full((B * H * (W + 2), 4), 3.0, device=dev, dtype=T16)
    out = torch.full((B * (H2 + 2) * (W2 + 2), 64), 7.0, device=dev, dtype=T16)
    ops.stem_pool(img, st.wrow, st.bias, hwc4, out, B, H, W, H1, W1, H2, W2)
    torch.cuda.synchronize()
    hv = hwc4.view(B, H, W + 2, 4)
    assert torch.equal(hv[:, :, 1:-1, :3], img.permute(0, 2, 3, 1).to(T16))
    assert hv[..., 3].abs().max().item() == 0 and hv[:, :, 0].abs().max().item() == 0 and hv[:, :, -1].abs().max().item() == 0
    w16 = st.wf[:, :147].float().reshape(64, 7, 7, 3).permute(0, 3, 1, 2)
    y = F.relu(F.conv2d(img.to(T16).float(), w16, st.bias, stride=2, padding=3))
    ref = F.max_pool2d(y, 3, 2, 1).permute(0, 2, 3, 1)
    o = out.view(B, H2 + 2, W2 + 2, 64).float()
    assert o[:, 0].abs().max().item() == 0 and o[:, -1].abs().max().item() == 0
    assert o[:, :, 0].abs().max().item() == 0 and o[:, :, -1].abs().max().item() == 0
    assert _rel(o[:, 1:-1, 1:-1], ref) < 2e-3
    assert (o[:, 1:-1, 1:-1] - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    # and against the two-kernel path it replaces
    c1 = torch.empty(B * H1 * W1, 64, device=dev, dtype=T16)
    x2 = torch.empty_like(out)
    ops.stem_conv(img, st.wf, st.bias, c1, B, H, W, H1, W1)
    ops.maxpool_3x3s2(c1, x2, B, H1, W1, 64, H2, W2)
    assert _rel(out, x2) < 2e-3
