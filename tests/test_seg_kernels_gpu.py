"""GPU parity of the segmentation-head kernels (through the C ABI) against their plain-PyTorch emulation (tests/emu_ops.py,
itself pinned against the oracle by tests/test_engine_emulated.py)."""
import pytest
import torch

import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu_ops

pytestmark = pytest.mark.gpu
dev = "cuda"
from reftr_b200 import ops as _ops
BF = _ops.t16()


def _close(a, b, tol):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs().max().item() / (b.abs().max().item() + 1e-6)
    assert err < tol, err


def test_tokens_grid_roundtrip_and_attn_map():
    from reftr_b200 import ops
    B, L, h, w = 2, 5, 6, 7
    S, hw = L + h * w, h * w
    tok = torch.randn(B * S, 256)
    q = torch.randn(B, 256) * 0.3
    k = torch.randn(B * S, 256) * 0.3
    kpm = torch.zeros(B, S, dtype=torch.uint8)
    kpm[1, L + 30:] = 1
    R = B * (h + 2) * (w + 2)
    grid_c, grid_g = torch.zeros(R, 520, dtype=BF), torch.zeros(R, 520, dtype=BF, device=dev)
    att_c, att_g = torch.empty(B, 8, hw), torch.empty(B, 8, hw, device=dev)
    emu_ops.tokens_to_grid(tok, B, S, L, h, w, 256, grid_c, 256)
    ops.tokens_to_grid(tok.to(dev), B, S, L, h, w, 256, grid_g, 256)
    emu_ops.attn_map_fwd(q, k, kpm, B, S, L, hw, w, 32 ** -0.5, att_c, grid_c, 512)
    ops.attn_map_fwd(q.to(dev), k.to(dev), kpm.to(dev), B, S, L, hw, w, 32 ** -0.5, att_g, grid_g, 512)
    _close(att_g, att_c, 1e-4)
    _close(grid_g, grid_c, 1e-2)
    dgrid = (torch.randn(R, 520) * 0.1).to(BF)
    dext = torch.randn(B, 8, hw) * 0.1
    dq_c, dk_c = torch.empty(B, 256), torch.empty(B * S, 256)
    dq_g, dk_g = torch.empty(B, 256, device=dev), torch.full((B * S, 256), 7.0, device=dev)
    emu_ops.attn_map_bwd(dext, dgrid, 512, att_c, q, k, B, S, L, hw, w, 32 ** -0.5, dq_c, dk_c)
    ops.attn_map_bwd(dext.to(dev), dgrid.to(dev), 512, att_g, q.to(dev), k.to(dev), B, S, L, hw, w, 32 ** -0.5, dq_g, dk_g)
    _close(dq_g, dq_c, 1e-3)
    _close(dk_g, dk_c, 1e-3)
    t_c, t_g = torch.zeros(B * S, 256), torch.zeros(B * S, 256, device=dev)
    emu_ops.grid_to_tokens(dgrid, 256, B, S, L, h, w, 256, t_c)
    ops.grid_to_tokens(dgrid.to(dev), 256, B, S, L, h, w, 256, t_g)
    assert torch.equal(t_g.cpu(), t_c)


@pytest.mark.parametrize("C,G,H,W", [(520, 8, 5, 6), (128, 8, 10, 12), (16, 8, 40, 48)])
def test_groupnorm_nhwc(C, G, H, W):
    from reftr_b200 import ops
    B = 2
    R = B * (H + 2) * (W + 2)
    x = torch.randn(R, C) * 1.5 + 0.3
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    y_c, y_g = torch.empty(R, C, dtype=BF), torch.empty(R, C, dtype=BF, device=dev)
    m_c, r_c = torch.empty(B * G), torch.empty(B * G)
    m_g, r_g = torch.empty(B * G, device=dev), torch.empty(B * G, device=dev)
    emu_ops.groupnorm_nhwc_fwd(x, gamma, beta, B, H, W, C, G, y_c, m_c, r_c)
    ops.groupnorm_nhwc_fwd(x.to(dev), gamma.to(dev), beta.to(dev), B, H, W, C, G, y_g, m_g, r_g)
    _close(m_g, m_c, 1e-4)
    _close(r_g, r_c, 1e-4)
    _close(y_g, y_c, 1e-2)
    dy = (torch.randn(R, C) * 0.2).to(BF)
    dx_c, dx_g = torch.empty(R, C, dtype=BF), torch.empty(R, C, dtype=BF, device=dev)
    dg_c, db_c = torch.zeros(C), torch.zeros(C)
    dg_g, db_g = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    emu_ops.groupnorm_nhwc_bwd(dy, y_c, x, gamma, m_c, r_c, B, H, W, C, G, dx_c, dg_c, db_c)
    ops.groupnorm_nhwc_bwd(dy.to(dev), y_c.to(dev), x.to(dev), gamma.to(dev), m_c.to(dev), r_c.to(dev), B, H, W, C, G, dx_g, dg_g, db_g)
    _close(dx_g, dx_c, 1.5e-2)
    _close(dg_g, dg_c, 1e-3)
    _close(db_g, db_c, 1e-3)


@pytest.mark.parametrize("h,w,H,W", [(5, 6, 10, 12), (10, 12, 20, 24), (5, 6, 9, 13)])
def test_upsample_add_and_bwd(h, w, H, W):
    from reftr_b200 import ops
    B, C = 2, 32
    lo = torch.randn(B * (h + 2) * (w + 2), C).to(BF)
    cur = torch.randn(B * (H + 2) * (W + 2), C).to(BF)
    y_c, y_g = torch.empty_like(cur), torch.empty_like(cur, device=dev)
    emu_ops.upsample_add(lo, cur, y_c, B, h, w, H, W, C)
    ops.upsample_add(lo.to(dev), cur.to(dev), y_g, B, h, w, H, W, C)
    _close(y_g, y_c, 1e-2)
    dy = torch.randn_like(cur)
    d_c, d_g = torch.empty_like(lo), torch.empty_like(lo, device=dev)
    emu_ops.upsample_bwd(dy, d_c, B, h, w, H, W, C)
    ops.upsample_bwd(dy.to(dev), d_g, B, h, w, H, W, C)
    _close(d_g, d_c, 1e-2)
