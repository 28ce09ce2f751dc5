"""GPU parity of the tcgen05 GEMM / implicit-conv kernel (rb_gemm, through the C ABI) against a torch fp32 reference
of the same op computed from the same bf16-rounded operands.  Tolerance: fp32 accumulation order only -> 2e-3 of the
output scale (bf16 output rounding adds 2^-8 relative)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
from reftr_b200 import ops as _ops
T16 = _ops.t16()  # the library's 16-bit operand type (IEEE half by default)


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6)).item()


@pytest.mark.parametrize("M,N,K,bn", [(300, 96, 160, 0), (128, 32, 64, 32), (1000, 64, 576, 64), (777, 256, 2048, 128),
                                      (512, 512, 256, 256), (16, 256, 256, 0), (4100, 2048, 256, 0)])
def test_gemm_nt_linear(M, N, K, bn):
    from reftr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(T16)
    B = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(T16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(T16)
    out = torch.full((M, N), 7.0, device="cuda", dtype=T16)
    out32 = torch.full((M, N), 7.0, device="cuda")
    ops.gemm(A, B, M, N, K, bias=bias, res=res, relu=True, out=out, out32=out32, block_n=bn)
    ref = torch.relu(A.float() @ B.float().t() + bias + res.float())
    assert _rel(out32, ref) < 2e-3
    assert _rel(out, ref) < 8e-3


def test_gemm_nt_ragged_n_and_res32():
    from reftr_b200 import ops
    M, N, K = 50, 4, 256
    A = torch.randn(M, K, device="cuda").to(T16)
    B = torch.zeros(8, K, device="cuda", dtype=T16)
    B[:N] = (torch.randn(N, K, device="cuda") / 16).to(T16)
    bias = torch.randn(N, device="cuda")
    res32 = torch.randn(M, N, device="cuda")
    out32 = torch.zeros(M, N, device="cuda")
    ops.gemm(A, B, M, N, K, bias=bias, out32=out32)
    ref = A.float() @ B[:N].float().t() + bias
    assert _rel(out32, ref) < 2e-3


def _pad_nhwc(x):  # NCHW fp32 -> padded NHWC bf16 [N, H+2, W+2, C]
    return F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).to(T16).contiguous()


@pytest.mark.parametrize("Nb,H,W,Cin,Cout", [(2, 20, 20, 64, 64), (3, 14, 10, 128, 256), (1, 40, 40, 256, 128)])
def test_conv3x3_as_shifted_gemm(Nb, H, W, Cin, Cout):
    from reftr_b200 import ops
    x = torch.randn(Nb, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (Cin * 9) ** 0.5
    bias = torch.randn(Cout, device="cuda")
    xp = _pad_nhwc(x)
    Wp, Hp = W + 2, H + 2
    R = Nb * Hp * Wp
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(T16).contiguous()  # [Cout, (r,s), Cin]
    taps = [((r - 1) * Wp + (s - 1), (r * 3 + s) * Cin) for r in range(3) for s in range(3)]
    out = torch.full((R, Cout), 5.0, device="cuda", dtype=T16)
    ops.gemm(xp.view(R, Cin), wk, R, Cout, Cin, taps=taps, bias=bias, relu=True, out=out,
             geom=ops.make_geom(1, Wp, Hp * Wp, H, W))
    ref = F.relu(F.conv2d(xp[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float(), wk.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).float(),
                          bias, padding=1))
    got = out.view(Nb, Hp, Wp, Cout)
    assert _rel(got[:, 1:-1, 1:-1].permute(0, 3, 1, 2), ref) < 1e-2
    border = got.clone()
    border[:, 1:-1, 1:-1] = 0
    assert border.abs().max().item() == 0.0  # padding rows must be written as exact zeros


@pytest.mark.parametrize("R,Mo,No,splits,bn", [(1000, 128, 192, 3, 64), (6400, 256, 128, 8, 128), (200, 64, 64, 1, 64),
                                             (333, 512, 256, 2, 256)])
def test_gemm_tn_wgrad(R, Mo, No, splits, bn):
    from reftr_b200 import ops
    dY = torch.randn(R, Mo, device="cuda").to(T16)
    X = torch.randn(R, No, device="cuda").to(T16)
    out32 = torch.zeros(Mo, No, device="cuda")
    ops.gemm(dY, X, Mo, No, R, mode=1, out32=out32, atomic=True, splits=splits, block_n=bn)
    ref = dY.float().t() @ X.float()
    assert _rel(out32, ref) < 2e-3


@pytest.mark.parametrize("R,Mo,No,splits,bn", [(1000, 128, 192, 3, 0), (6400, 256, 512, 8, 0), (200, 64, 64, 1, 64), (333, 768, 768, 2, 256),
                                             (16, 4, 256, 1, 0), (6720, 2048, 256, 14, 0), (320, 3072, 768, 1, 0)])
def test_gemm_tn_wgrad_with_bias_gradient_and_row_scale(R, Mo, No, splits, bn):
    """bias_grad (column sums of dY from an extra N=16 MMA against ones in the same launch), row_scale and out_scale of the atomic
    epilogue; the accumulation targets start non-zero (the kernel ADDS)."""
    from reftr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(R + Mo)
    dY = torch.randn(R, max(Mo, 64), device="cuda", generator=g).to(T16)[:, :Mo]   # (row pitch stays a multiple of 16 bytes: the 4-wide box head)
    X = torch.randn(R, No, device="cuda", generator=g).to(T16)
    rs = torch.rand(Mo, device="cuda", generator=g) + 0.5
    out32 = torch.full((Mo, No), 0.25, device="cuda")
    bg = torch.full((Mo,), -1.0, device="cuda")
    ops.gemm(dY, X, Mo, No, R, mode=1, out32=out32, atomic=True, splits=splits if bn else 0, block_n=bn, bias_grad=bg, row_scale=rs, out_scale=0.5)
    ref = 0.25 + 0.5 * rs[:, None] * (dY.float().t() @ X.float())
    assert _rel(out32, ref) < 2e-3
    ref_b = -1.0 + 0.5 * dY.float().sum(0)
    assert (bg - ref_b).abs().max().item() < 2e-3 * max(1.0, ref_b.abs().max().item())


def test_conv3x3_wgrad_taps_with_bias_gradient():
    """With several taps the bias gradient must be accumulated once (by tap 0's tiles), not once per tap."""
    from reftr_b200 import ops
    Nb, H, W, Cin, Cout = 2, 12, 9, 128, 128
    Wp, Hp = W + 2, H + 2
    R = Nb * Hp * Wp
    dyp = _pad_nhwc(torch.randn(Nb, Cout, H, W, device="cuda"))
    xp = _pad_nhwc(torch.randn(Nb, Cin, H, W, device="cuda"))
    taps = [(0, (r - 1) * Wp + (s - 1)) for r in range(3) for s in range(3)]
    out32 = torch.zeros(Cout, 9 * Cin, device="cuda")
    bg = torch.zeros(Cout, device="cuda")
    ops.gemm(dyp.view(R, Cout), xp.view(R, Cin), Cout, Cin, R, mode=1, taps=taps, out32=out32, atomic=True, splits=2, out32_z_stride=Cin, bias_grad=bg)
    ref_b = dyp.view(R, Cout).float().sum(0)
    assert (bg - ref_b).abs().max().item() < 2e-3 * ref_b.abs().max().item()


def test_conv3x3_wgrad_taps():
    from reftr_b200 import ops
    Nb, H, W, Cin, Cout = 2, 12, 9, 128, 128
    Wp, Hp = W + 2, H + 2
    R = Nb * Hp * Wp
    x = torch.randn(Nb, Cin, H, W, device="cuda")
    dy = torch.randn(Nb, Cout, H, W, device="cuda")
    xp, dyp = _pad_nhwc(x), _pad_nhwc(dy)
    taps = [(0, (r - 1) * Wp + (s - 1)) for r in range(3) for s in range(3)]
    out32 = torch.zeros(Cout, 9, Cin, device="cuda")
    ops.gemm(dyp.view(R, Cout), xp.view(R, Cin), Cout, Cin, R, mode=1, taps=taps, out32=out32.view(Cout, 9 * Cin),
             atomic=True, splits=2, out32_z_stride=Cin)
    xr = xp[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().requires_grad_()
    wz = torch.zeros(Cout, Cin, 3, 3, device="cuda", requires_grad=True)
    F.conv2d(xr, wz, padding=1).backward(dyp[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float())
    ref = wz.grad.permute(0, 2, 3, 1).reshape(Cout, 9, Cin)
    assert _rel(out32, ref) < 2e-3


@pytest.mark.parametrize("M,N,K", [(16, 256, 256), (16, 2048, 256), (16, 256, 2048), (96, 64, 256), (96, 256, 64), (5, 4, 256), (80, 768, 768), (128, 256, 512),
                                   (1, 256, 256)])
def test_gemm_skinny_matches_torch_and_big_kernel(M, N, K):
    """M <= 128 rows dispatch to the mma.sync latency kernel (gemm_skinny.cu): same results as torch and as the tcgen05 kernel for
    every epilogue combination the decoder / query encoder / box head use."""
    import os
    from reftr_b200 import ops
    import dropout_ref
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(T16)
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(T16)
    bias = torch.randn(N, device="cuda", generator=g)
    pad = (N + 7) // 8 * 8  # 16-bit operands of the epilogue need a 16-byte row pitch
    res = torch.randn(M, pad, device="cuda", generator=g).to(T16)[:, :N]
    res32 = torch.randn(M, pad, device="cuda", generator=g)[:, :N]
    msk = torch.randn(M, pad, device="cuda", generator=g).to(T16)[:, :N]
    lin = A.float() @ W.float().t() + bias
    seed = torch.full((1,), 1234567, dtype=torch.int64, device="cuda")
    cases = [
        (dict(bias=bias, res=res, relu=True), torch.relu(lin + res.float())),
        (dict(bias=bias, res32=res32), lin + res32),
        (dict(mask_src=msk, mask_scale=1.5), torch.where(msk.float() > 0, (lin - bias) * 1.5, torch.zeros((), device="cuda"))),
    ]
    if N % 2 == 0:
        d1 = ops.Drop(seed, "sk.a", 0.1)
        m1 = dropout_ref.mask_scale(1234567, "sk.a", M, N, 0.1, device="cuda")
        cases.append((dict(bias=bias, res32=res32, drop=d1), res32 + lin * m1))
        cases.append((dict(bias=bias, relu=True, drop=d1), torch.relu(lin) * m1))
    if N % 32 == 0:
        d2 = ops.Drop(seed, "sk.h", 0.3)
        m2 = dropout_ref.mask_scale(1234567, "sk.h", M, N // 32, 0.3, device="cuda").repeat_interleave(32, dim=1)
        cases.append((dict(bias=bias, drop=d2, drop_gshift=5), lin * m2))
    for kw, ref in cases:
        outs = []
        for no_skinny in ("", "1"):
            if no_skinny:
                os.environ["RB_GEMM_NO_SKINNY"] = "1"
            try:
                ob = torch.full((M, pad), 3.0, device="cuda", dtype=T16)[:, :N]
                o32 = torch.full((M, pad), 3.0, device="cuda")[:, :N]
                ops.gemm(A, W, M, N, K, out=ob, out32=o32, **kw)
            finally:
                os.environ.pop("RB_GEMM_NO_SKINNY", None)
            assert _rel(o32, ref) < 2e-3, (kw.keys(), no_skinny)
            assert _rel(ob, ref) < 8e-3
            outs.append(o32)
        assert _rel(outs[0], outs[1]) < 1e-4  # the two kernels agree to fp32 summation order


@pytest.mark.parametrize("M,N,K", [(6720, 256, 256), (6720, 256, 2048), (320, 768, 3072), (320, 2304, 768), (16, 256, 256)])
def test_gemm_hilo_weight_pairs(M, N, K):
    """(hi | residual) weight pairs (rb_pack_linear_hilo + a two-tap rb_gemm that reads the same A rows twice): the product with the
    fp32 WEIGHT, i.e. what remains is the 16-bit rounding of the activations alone.  Checked against fp32 matmul of the 16-bit
    activations with the fp32 weight -- to accumulation-order accuracy, ~50x tighter than the single-16-bit-weight product is."""
    from reftr_b200 import ops
    from reftr_b200.pack import PackedLinear, lin_taps
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(T16)
    w = torch.nn.Parameter(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    bias = torch.nn.Parameter(torch.randn(N, device="cuda", generator=g))
    pk = PackedLinear(w, bias, hilo=True)
    pk.refresh()
    assert pk.wb2.shape == (N, 2 * K)
    assert torch.equal(pk.wb2[:, :K], pk.wb[:N])                                   # hi = the plain 16-bit weight
    assert torch.equal(pk.wb2[:, K:], (w.detach() - pk.wb[:N].float()).to(T16))   # residual
    B, taps = lin_taps(pk)
    assert len(taps) == 2 and taps[1] == (0, K)
    out32 = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, M, N, K, taps=taps, bias=pk.bias, out32=out32)
    ref = A.float() @ w.detach().t() + bias.detach()
    single = A.float() @ pk.wb[:N].float().t() + bias.detach()
    e_pair = ((out32 - ref).norm() / ref.norm()).item()
    e_single = ((single - ref).norm() / ref.norm()).item()
    print(f"hi/lo M{M} N{N} K{K}: rel-L2 vs fp32 weights {e_pair:.2e} (one 16-bit weight: {e_single:.2e})")
    assert e_pair < 2e-5 and e_pair < 0.1 * e_single


@pytest.mark.parametrize("limit", [2, 8, 36, 200])
def test_gemm_sm_limit_gives_the_same_result(limit):
    """rb_gemm_args.sm_limit only narrows the persistent grid (and the tile / split choice): NT output bit-identical for a fixed tile
    width, TN weight gradient equal up to the order of the fp32 atomics."""
    from reftr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(limit)
    M, N, K = 1500, 768, 512
    A = torch.randn(M, K, device="cuda", generator=g).to(T16)
    B = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(T16)
    o0, o1 = torch.empty(M, N, device="cuda", dtype=T16), torch.empty(M, N, device="cuda", dtype=T16)
    ops.gemm(A, B, M, N, K, out=o0, block_n=128)
    ops.gemm(A, B, M, N, K, out=o1, block_n=128, sm_limit=limit)
    assert torch.equal(o0, o1)
    with ops.sm_limit_scope(limit):
        o2 = torch.empty(M, N, device="cuda", dtype=T16)
        ops.gemm(A, B, M, N, K, out=o2, block_n=128)
    assert torch.equal(o0, o2)
    R, Mo, No = 2000, 256, 384
    dY = torch.randn(R, Mo, device="cuda", generator=g).to(T16)
    X = torch.randn(R, No, device="cuda", generator=g).to(T16)
    gw = torch.zeros(Mo, No, device="cuda")
    gb = torch.zeros(Mo, device="cuda")
    ops.gemm(dY, X, Mo, No, R, mode=1, out32=gw, atomic=True, splits=0, bias_grad=gb, sm_limit=limit)
    ref = dY.float().t() @ X.float()
    assert _rel(gw, ref) < 1e-4
    assert _rel(gb, dY.float().sum(0)) < 1e-4
