"""GPU parity of the BERT support kernels (through the C ABI) against their plain-PyTorch emulation (tests/emu_ops.py)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu_ops  # noqa: E402

pytestmark = pytest.mark.gpu
dev = "cuda"
from reftr_b200 import ops as _ops
BF = _ops.t16()


def _close(a, b, tol):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs().max().item() / (b.abs().max().item() + 1e-6)
    assert err < tol, err


def test_embed_ln_gelu_tanh():
    from reftr_b200 import ops
    rows, L, D, V = 44, 11, 768, 1000
    ids = torch.randint(0, V, (rows,))
    word, pos, typ = torch.randn(V, D), torch.randn(64, D), torch.randn(2, D)
    o_c, o_g = torch.empty(rows, D), torch.empty(rows, D, device=dev)
    emu_ops.bert_embed_fwd(ids, L, word, pos, typ[0], o_c)
    ops.bert_embed_fwd(ids.to(dev), L, word.to(dev), pos.to(dev), typ.to(dev)[0], o_g)
    _close(o_g, o_c, 1e-6)
    d = torch.randn(rows, D)
    dw_c, dp_c, dt_c = torch.zeros(V, D), torch.zeros(64, D), torch.zeros(D)
    dw_g, dp_g, dt_g = torch.zeros(V, D, device=dev), torch.zeros(64, D, device=dev), torch.zeros(D, device=dev)
    emu_ops.bert_embed_bwd(d, ids, L, dw_c, dp_c, dt_c)
    ops.bert_embed_bwd(d.to(dev), ids.to(dev), L, dw_g, dp_g, dt_g)
    _close(dw_g, dw_c, 1e-5); _close(dp_g, dp_c, 1e-5); _close(dt_g, dt_c, 1e-5)
    # LayerNorm 768
    x = torch.randn(rows, D) * 2 + 0.5
    g, b = 1 + 0.1 * torch.randn(D), 0.1 * torch.randn(D)
    y_c, yb_c, m_c, r_c = torch.empty(rows, D), torch.empty(rows, D, dtype=BF), torch.empty(rows), torch.empty(rows)
    y_g, yb_g, m_g, r_g = (t.to(dev) for t in (torch.empty(rows, D), torch.empty(rows, D, dtype=BF), torch.empty(rows), torch.empty(rows)))
    emu_ops.ln_wide_fwd(x, g, b, rows, y32=y_c, yb=yb_c, mean=m_c, rstd=r_c)
    ops.ln_wide_fwd(x.to(dev), g.to(dev), b.to(dev), rows, y32=y_g, yb=yb_g, mean=m_g, rstd=r_g)
    _close(y_g, y_c, 1e-5); _close(yb_g, yb_c, 1e-2); _close(r_g, r_c, 1e-5)
    dy, dy2 = torch.randn(rows, D), torch.randn(rows, D)
    dx_c, dg_c, db_c = torch.empty(rows, D), torch.zeros(D), torch.zeros(D)
    dx_g, dg_g, db_g = torch.empty(rows, D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    emu_ops.ln_wide_bwd(dy, x, g, m_c, r_c, rows, dy2=dy2, dx32=dx_c, dgamma=dg_c, dbeta=db_c)
    ops.ln_wide_bwd(dy.to(dev), x.to(dev), g.to(dev), m_g, r_g, rows, dy2=dy2.to(dev), dx32=dx_g, dgamma=dg_g, dbeta=db_g)
    _close(dx_g, dx_c, 1e-4); _close(dg_g, dg_c, 1e-4); _close(db_g, db_c, 1e-4)
    # GELU / tanh
    xb = (torch.randn(rows, 3072) * 2).to(BF)
    h_c, h_g = torch.empty_like(xb), torch.empty_like(xb, device=dev)
    emu_ops.gelu_fwd(xb, h_c); ops.gelu_fwd(xb.to(dev), h_g)
    _close(h_g, h_c, 1e-2)
    dh = torch.randn(rows, 3072).to(BF)
    e_c, e_g = torch.empty_like(xb), torch.empty_like(xb, device=dev)
    emu_ops.gelu_bwd(dh, xb, e_c); ops.gelu_bwd(dh.to(dev), xb.to(dev), e_g)
    _close(e_g, e_c, 1e-2)
    t = torch.randn(rows, D)
    t_c, t_g = torch.empty_like(t), torch.empty_like(t, device=dev)
    emu_ops.tanh_fwd(t, t_c); ops.tanh_fwd(t.to(dev), t_g)
    _close(t_g, t_c, 1e-5)


@pytest.mark.parametrize("B,S", [(3, 20), (2, 22), (2, 90), (1, 1)])
def test_attn_small(B, S):
    from reftr_b200 import ops
    H, D = 12, 768
    qkv = (torch.randn(B * S, 3 * D) * 0.5).to(BF)
    mask = torch.zeros(B, S, dtype=torch.uint8)
    if S > 4:
        mask[0, S - 3:] = 1
    o_c, o_g = torch.empty(B * S, D, dtype=BF), torch.empty(B * S, D, dtype=BF, device=dev)
    P_c, P_g = torch.empty(B, H, S, S), torch.empty(B, H, S, S, device=dev)
    emu_ops.attn_small_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], mask, o_c, P_c, B, H, S, 0.125)
    qg = qkv.to(dev)
    ops.attn_small_fwd(qg[:, :D], qg[:, D:2 * D], qg[:, 2 * D:], mask.to(dev), o_g, P_g, B, H, S, 0.125)
    _close(P_g, P_c, 1e-4); _close(o_g, o_c, 1e-2)
    do = torch.randn(B * S, D).to(BF)
    d_c, d_g = torch.empty(B * S, 3 * D, dtype=BF), torch.empty(B * S, 3 * D, dtype=BF, device=dev)
    emu_ops.attn_small_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], do, P_c, d_c[:, :D], d_c[:, D:2 * D], d_c[:, 2 * D:], B, H, S, 0.125)
    ops.attn_small_bwd(qg[:, :D], qg[:, D:2 * D], qg[:, 2 * D:], do.to(dev), P_g, d_g[:, :D], d_g[:, D:2 * D], d_g[:, 2 * D:], B, H, S, 0.125)
    _close(d_g, d_c, 1.5e-2)
