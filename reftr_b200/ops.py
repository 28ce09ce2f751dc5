"""Python-side launchers for the C-ABI kernels (include/reftr_b200.h).  torch is used for device memory and the
current stream only; all arithmetic happens in libreftr_b200.so."""
import ctypes as C

import torch

from . import _lib
from ._lib import Dropout, GemmArgs, Geom


SM_LIMIT = 0  # default rb_gemm_args.sm_limit of gemm(); set for a region by sm_limit_scope()
PROFILE = None  # bench.py sets this to a list: every rb_gemm launch descriptor is then recorded -> (args, flops, signature)


def require_device(t):
    """The kernels exist for sm_100a only; there is no CPU or PyTorch fallback."""
    if not t.is_cuda:
        raise RuntimeError("reftr_b200 runs on CUDA (sm_100a) only: got a tensor on %s" % t.device)
    _lib.lib()


def launch_count():
    return _lib.LAUNCHES


_T16 = None


def t16():
    """torch dtype of the library's 16-bit operand / activation type (rb_act_dtype: IEEE half by default, bf16 with -DRB_ACT_BF16)."""
    global _T16
    if _T16 is None:
        _T16 = torch.float16 if _lib.lib().rb_act_dtype() == 0 else torch.bfloat16
    return _T16


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def site_id(name):
    """Stable 31-bit id of a named dropout site (rb_dropout.site)."""
    import zlib
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


class Drop:
    """Host handle of one dropout site: ``seed`` is a 1-element int64 device tensor the engine rewrites every step."""

    def __init__(self, seed, name, p):
        assert seed.dtype == torch.int64 and seed.numel() == 1
        self.seed, self.name, self.p = seed, name, float(p)
        self.site = site_id(name)
        thr = min(int(self.p * 65536.0 + 0.5), 65535)
        self.thr = thr
        self.scale = 65536.0 / (65536 - thr)   # what kept values are multiplied by (1 / (1 - p) with p rounded to 1/65536)
        self.c = Dropout(seed.data_ptr(), self.site, self.p)

    def ptr(self):
        return C.addressof(self.c)


def _dp(drop):
    return None if drop is None else drop.ptr()


def make_geom(mode=0, Wp=0, HpWp=0, H=0, W=0, Rs=0):
    return Geom(mode, Wp, HpWp, H, W, Rs)


class sm_limit_scope:
    """``with sm_limit_scope(n):`` GEMMs launched inside use at most n SMs (rb_gemm_args.sm_limit); n <= 0 leaves it unchanged."""

    def __init__(self, n):
        self.n = int(n)

    def __enter__(self):
        global SM_LIMIT
        self.prev = SM_LIMIT
        if self.n > 0:
            SM_LIMIT = self.n

    def __exit__(self, *exc):
        global SM_LIMIT
        SM_LIMIT = self.prev
        return False


def _check_2d(t, dtype, name):
    assert t.is_cuda and t.dtype == dtype and t.dim() == 2 and t.stride(1) == 1, (name, t.dtype, t.shape, t.stride())


def gemm(A, B, M, N, K, *, mode=0, taps=((0, 0),), bias=None, res=None, res32=None, mask_src=None, relu=False,
         out=None, out32=None, atomic=False, splits=1, geom=None, out_row_off=0, out32_z_stride=0, block_n=0, drop=None, drop_gshift=0,
         mask_scale=1.0, bias_grad=None, row_scale=None, out_scale=1.0, sm_limit=None):
    """See rb_gemm in include/reftr_b200.h.  ``taps`` is a sequence of (a_rowoff, b_koff) pairs.  ``sm_limit`` None: the value of the
    enclosing ``sm_limit_scope`` (0 = all SMs)."""
    _check_2d(A, t16(), "A")
    _check_2d(B, t16(), "B")
    a = GemmArgs()
    a.mode = mode
    a.A, a.a_rows, a.a_cols, a.lda = A.data_ptr(), A.shape[0], A.shape[1], A.stride(0)
    a.B, a.b_rows, a.b_cols, a.ldb = B.data_ptr(), B.shape[0], B.shape[1], B.stride(0)
    a.M, a.N, a.K = M, N, K
    a.taps = len(taps)
    for i, (ro, ko) in enumerate(taps):
        a.a_rowoff[i] = ro
        a.b_koff[i] = ko
    a.splits = splits
    a.block_n = block_n
    a.out_row_off = out_row_off
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        a.bias = bias.data_ptr()
    if res is not None:
        _check_2d(res, t16(), "res")
        a.res, a.ldres = res.data_ptr(), res.stride(0)
    if res32 is not None:
        _check_2d(res32, torch.float32, "res32")
        a.res32, a.ldres32 = res32.data_ptr(), res32.stride(0)
    if mask_src is not None:
        _check_2d(mask_src, t16(), "mask_src")
        a.mask_src, a.ldmask = mask_src.data_ptr(), mask_src.stride(0)
    if out is not None:
        _check_2d(out, t16(), "out")
        a.out, a.ldo = out.data_ptr(), out.stride(0)
    if out32 is not None:
        assert out32.dtype == torch.float32 and out32.is_cuda
        a.out32, a.ldo32 = out32.data_ptr(), (out32.stride(-2) if out32.dim() >= 2 else N)
    a.out32_z_stride = out32_z_stride
    a.relu = int(relu)
    a.atomic = int(atomic)
    if geom is not None:
        a.geom = geom
    if drop is not None:
        a.drop, a.drop_gshift = drop.ptr(), drop_gshift
    a.mask_scale = mask_scale
    if bias_grad is not None:
        assert bias_grad.dtype == torch.float32 and bias_grad.is_contiguous() and bias_grad.numel() >= M
        a.bias_grad = bias_grad.data_ptr()
    if row_scale is not None:
        assert row_scale.dtype == torch.float32 and row_scale.is_contiguous() and row_scale.numel() >= M
        a.row_scale = row_scale.data_ptr()
    a.out_scale = out_scale
    a.sm_limit = SM_LIMIT if sm_limit is None else sm_limit
    if PROFILE is not None:  # bench.py: keep the launch descriptor so the launch can be re-issued and timed in isolation
        # algorithmic FLOPs: taps that read the SAME A rows are the (hi | residual) halves of one weight, counted once
        PROFILE.append((a, 2.0 * M * N * K * (len({ro for ro, _ in taps}) if mode == 0 else len(taps)), (mode, M, N, K, len(taps), bool(atomic), res is not None, res32 is not None,
                                                       mask_src is not None, out32 is not None)))
    _lib.check(_lib.lib().rb_gemm(C.byref(a), _stream()), "rb_gemm")
    return out if out is not None else out32


def relaunch_gemm(a):
    """Re-issues a recorded rb_gemm launch (same pointers) on the current stream."""
    _lib.check(_lib.lib().rb_gemm(C.byref(a), _stream()), "rb_gemm")


# ---------------------------------------------------------------------------------------------------------------
# thin wrappers (argument order == include/reftr_b200.h)
# ---------------------------------------------------------------------------------------------------------------
def _p(t):
    return None if t is None else t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def stem_im2col(img, out, B, H, W, H1, W1):
    _lib.call("rb_stem_im2col", _p(img), _p(out), B, H, W, H1, W1, _s())


def stem_conv(img, wf, bias, out, B, H, W, H1, W1):
    _lib.call("rb_stem_conv", _p(img), _p(wf), wf.stride(0), _p(bias), _p(out), B, H, W, H1, W1, _s())


def stem_pool(img, wpk, bias, hwc4, out, B, H, W, H1, W1, H2, W2):
    _lib.call("rb_stem_pool", _p(img), _p(wpk), _p(bias), _p(hwc4), _p(out), B, H, W, H1, W1, H2, W2, _s())


def maxpool_3x3s2(x, out, B, H1, W1, C, H2, W2):
    _lib.call("rb_maxpool_3x3s2", _p(x), _p(out), B, H1, W1, C, H2, W2, _s())


def parity_split(x, xs, B, H, W, C, Ho, Wo):
    _lib.call("rb_parity_split", _p(x), _p(xs), B, H, W, C, Ho, Wo, _s())


def parity_merge(dxs, add, mask_src, dx, B, H, W, C, Ho, Wo):
    _lib.call("rb_parity_merge", _p(dxs), _p(add), _p(mask_src), _p(dx), B, H, W, C, Ho, Wo, _s())


def parity_split_plane(x, xs_plane, B, H, W, C, Ho, Wo, plane):
    _lib.call("rb_parity_split_plane", _p(x), _p(xs_plane), B, H, W, C, Ho, Wo, plane, _s())


def parity_merge_plane(dxs_plane, plane, add, mask_src, dx, B, H, W, C, Ho, Wo):
    _lib.call("rb_parity_merge_plane", _p(dxs_plane), plane, _p(add), _p(mask_src), _p(dx), B, H, W, C, Ho, Wo, _s())


def pack_conv(w, bn, conv_bias, fwd, ldk, dgr, scale_out, bias_out, eps=1e-5):
    Cout, Cin, kh, kw = w.shape
    bw, bb, brm, brv = bn if bn is not None else (None, None, None, None)
    _lib.call("rb_pack_conv", _p(w), Cout, Cin, kh, kw, _p(bw), _p(bb), _p(brm), _p(brv), eps, _p(conv_bias), _p(fwd), ldk, _p(dgr),
              _p(scale_out), _p(bias_out), _s())


def pack_linear(w, wb, wt):
    N, K = w.shape
    assert w.is_contiguous()
    _lib.call("rb_pack_linear", _p(w), N, K, _p(wb), wb.stride(0) if wb is not None else 0, _p(wt), wt.stride(0) if wt is not None else 0, _s())


def pack_linear_hilo(w, w2):
    N, K = w.shape
    assert w.is_contiguous() and w2.shape[1] == 2 * K
    _lib.call("rb_pack_linear_hilo", _p(w), N, K, _p(w2), w2.stride(0), _s())


def unpack_conv_grad(dwf, scale, grad, Cout, Cin, taps):
    _lib.call("rb_unpack_conv_grad", _p(dwf), _p(scale), _p(grad), Cout, Cin, taps, _s())


def cast_bf16(x, out=None):
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=t16(), device=x.device)
    _lib.call("rb_cast_bf16", _p(x), _p(out), x.numel(), _s())
    return out


def colsum(x, out, rows=None, N=None):
    """out[N] += column sums of x [rows, N] (bf16 or fp32)."""
    rows = x.shape[0] if rows is None else rows
    N = x.shape[1] if N is None else N
    _lib.call("rb_colsum", _p(x), int(x.dtype == t16()), x.stride(0), rows, N, _p(out), _s())


def add(a, b, y=None, yb=None):
    _lib.call("rb_add", _p(a), _p(b), _p(y), _p(yb), a.numel(), _s())


def layernorm_fwd(x, gamma, beta, rows, *, y32=None, yb=None, pos32=None, ypb=None, relu=False, mean=None, rstd=None, rowmap=(0, 0, 0),
                  eps=1e-5, drop=None):
    _lib.call("rb_layernorm_fwd", _p(x), _p(gamma), _p(beta), rows, x.shape[-1], eps, _p(y32), _p(yb), _p(pos32), _p(ypb), int(relu),
              _p(mean), _p(rstd), rowmap[0], rowmap[1], rowmap[2], _dp(drop), _s())


def layernorm_bwd(dy, x, gamma, mean, rstd, rows, *, dy2=None, y_relu=None, relu_scale=1.0, dx32=None, dxb=None, dgamma=None, dbeta=None,
                  rowmap=(0, 0, 0), dxb_drop=None):
    _lib.call("rb_layernorm_bwd", _p(dy), _p(dy2), _p(y_relu), relu_scale, _p(x), _p(gamma), _p(mean), _p(rstd), rows, x.shape[-1], _p(dx32),
              _p(dxb), _p(dgamma), _p(dbeta), rowmap[0], rowmap[1], rowmap[2], _dp(dxb_drop), _s())


def groupnorm_tokens_fwd(x, gamma, beta, B, h, w, S, L, y32, yb, pos32, ypb, mean, rstd, eps=1e-5):
    _lib.call("rb_groupnorm_tokens_fwd", _p(x), _p(gamma), _p(beta), B, h, w, S, L, eps, _p(y32), _p(yb), _p(pos32), _p(ypb), _p(mean),
              _p(rstd), _s())


def groupnorm_tokens_bwd(dy, dy2, x, gamma, mean, rstd, B, h, w, S, L, dx, dgamma, dbeta):
    _lib.call("rb_groupnorm_tokens_bwd", _p(dy), _p(dy2), _p(x), _p(gamma), _p(mean), _p(rstd), B, h, w, S, L, _p(dx), _p(dgamma),
              _p(dbeta), _s())


def build_pos_mask(img_mask, B, H, W, h, w, sent_mask, L, lang_pos, token_type, level_embed, pos32, kpm):
    assert img_mask.dtype == torch.bool and sent_mask.dtype == torch.int64
    _lib.call("rb_build_pos_mask", _p(img_mask), B, H, W, h, w, _p(sent_mask), L, _p(lang_pos), _p(token_type), _p(level_embed), _p(pos32),
              _p(kpm), _s())


def embed_grad(dpos, B, S, L, d_lang_pos, d_token_type, d_level):
    _lib.call("rb_embed_grad", _p(dpos), B, S, L, _p(d_lang_pos), _p(d_token_type), _p(d_level), _s())


def attn_fwd(Q, K, V, kpm, O, LSE, B, H, Tq, Sk, scale, drop=None):
    _lib.call("rb_attn_fwd", _p(Q), _p(K), _p(V), _p(kpm), _p(O), _p(LSE), B, H, 32, Tq, Sk, Q.stride(0), K.stride(0), V.stride(0),
              O.stride(0), scale, _dp(drop), _s())


def attn_bwd(Q, K, V, kpm, O, dO, LSE, dQ, dK, dV, Dbuf, B, H, Tq, Sk, scale, drop=None):
    _lib.call("rb_attn_bwd", _p(Q), _p(K), _p(V), _p(kpm), _p(O), _p(dO), _p(LSE), _p(dQ), _p(dK), _p(dV), _p(Dbuf), B, H, 32, Tq, Sk,
              Q.stride(0), K.stride(0), V.stride(0), O.stride(0), dO.stride(0), dQ.stride(0), dK.stride(0), dV.stride(0), scale, _dp(drop), _s())


def qenc_pool_fwd(k, q, v, mask, B, L, n_ph, att, c):
    _lib.call("rb_qenc_pool_fwd", _p(k), _p(q), _p(v), _p(mask), B, L, n_ph, _p(att), _p(c), _s())


def qenc_pool_bwd(dc, k, q, v, att, B, L, n_ph, dk, dq, dv):
    _lib.call("rb_qenc_pool_bwd", _p(dc), _p(k), _p(q), _p(v), _p(att), B, L, n_ph, _p(dk), _p(dq), _p(dv), _s())


def _map(m):
    """(group, stride, inner, offset) -> host int[4] (kept alive by the caller's frame) or NULL."""
    if m is None:
        return None, None
    arr = (C.c_int * 4)(*m)
    return arr, C.addressof(arr)


def rows_add(a, b, rows, D, *, y32=None, yb=None, map_a=None, map_b=None, map_y=None):
    """y[my(r), :D] = a[ma(r), :D] + b[mb(r), :D]; tensors are 2-D fp32 views (pitch = stride(0))."""
    ka, pa = _map(map_a)
    kb, pb = _map(map_b)
    ky, py = _map(map_y)
    _lib.call("rb_rows_add", _p(a), a.stride(0), pa, _p(b), b.stride(0) if b is not None else 0, pb, _p(y32),
              y32.stride(0) if y32 is not None else 0, _p(yb), yb.stride(0) if yb is not None else 0, py, rows, D, _s())


def rows_scatter_add(src, dst, rows, D, *, map_src=None, map_dst=None):
    ks, ps = _map(map_src)
    kd, pd = _map(map_dst)
    _lib.call("rb_rows_scatter_add", _p(src), src.stride(0), ps, _p(dst), dst.stride(0), pd, rows, D, _s())


# ---------------------------------------------------------------------------------------------------------------
# segmentation head
# ---------------------------------------------------------------------------------------------------------------
def tokens_to_grid(tok, B, S, L, h, w, C, grid, col0):
    _lib.call("rb_tokens_to_grid", _p(tok), B, S, L, h, w, C, _p(grid), grid.stride(0), col0, _s())


def grid_to_tokens(grid, col0, B, S, L, h, w, C, dtok):
    _lib.call("rb_grid_to_tokens", _p(grid), grid.stride(0), col0, B, S, L, h, w, C, _p(dtok), _s())


def attn_map_fwd(q, k, kpm, B, S, L, hw, w, scale, att, grid, col0):
    _lib.call("rb_attn_map_fwd", _p(q), _p(k), _p(kpm), B, S, L, hw, w, scale, _p(att), _p(grid), grid.stride(0), col0, _s())


def attn_map_bwd(datt_ext, dgrid, col0, att, q, k, B, S, L, hw, w, scale, dq, dk):
    _lib.call("rb_attn_map_bwd", _p(datt_ext), _p(dgrid), dgrid.stride(0), col0, _p(att), _p(q), _p(k), B, S, L, hw, w, scale, _p(dq), _p(dk), _s())


def groupnorm_nhwc_fwd(x, gamma, beta, B, H, W, C, G, y, mean, rstd, relu=True, eps=1e-5):
    _lib.call("rb_groupnorm_nhwc_fwd", _p(x), _p(gamma), _p(beta), B, H, W, C, G, eps, int(relu), _p(y), _p(mean), _p(rstd), _s())


def groupnorm_nhwc_bwd(dy, y, x, gamma, mean, rstd, B, H, W, C, G, dx, dgamma, dbeta, relu=True):
    _lib.call("rb_groupnorm_nhwc_bwd", _p(dy), _p(y), _p(x), _p(gamma), _p(mean), _p(rstd), B, H, W, C, G, int(relu), _p(dx), _p(dgamma), _p(dbeta), _s())


def upsample_add(lo, cur, y, B, h, w, H, W, C):
    _lib.call("rb_upsample_add", _p(lo), _p(cur), _p(y), B, h, w, H, W, C, _s())


def upsample_bwd(dy, dlo, B, h, w, H, W, C):
    _lib.call("rb_upsample_bwd", _p(dy), _p(dlo), B, h, w, H, W, C, _s())


# ---------------------------------------------------------------------------------------------------------------
# BERT
# ---------------------------------------------------------------------------------------------------------------
def bert_embed_fwd(ids, L, word, pos, type0, out):
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    _lib.call("rb_bert_embed_fwd", _p(ids), ids.numel(), L, word.shape[1], _p(word), _p(pos), _p(type0), _p(out), _s())


def bert_embed_bwd(d, ids, L, dword, dpos, dtype0):
    _lib.call("rb_bert_embed_bwd", _p(d), _p(ids), ids.numel(), L, d.shape[1], _p(dword), _p(dpos), _p(dtype0), _s())


def ln_wide_fwd(x, gamma, beta, rows, *, y32=None, yb=None, mean=None, rstd=None, eps=1e-12, drop=None):
    _lib.call("rb_ln_wide_fwd", _p(x), _p(gamma), _p(beta), rows, x.shape[-1], eps, _p(y32), _p(yb), _p(mean), _p(rstd), _dp(drop), _s())


def ln_wide_bwd(dy, x, gamma, mean, rstd, rows, *, dy2=None, dx32=None, dxb=None, dgamma=None, dbeta=None, dy_drop=None, dxb_drop=None):
    _lib.call("rb_ln_wide_bwd", _p(dy), _p(dy2), _p(x), _p(gamma), _p(mean), _p(rstd), rows, x.shape[-1], _p(dx32), _p(dxb), _p(dgamma), _p(dbeta),
              _dp(dy_drop), _dp(dxb_drop), _s())


def gelu_fwd(x, y):
    _lib.call("rb_gelu_fwd", _p(x), _p(y), x.numel(), _s())


def gelu_bwd(dy, x, dx):
    _lib.call("rb_gelu_bwd", _p(dy), _p(x), _p(dx), x.numel(), _s())


def tanh_fwd(x, y):
    _lib.call("rb_tanh_fwd", _p(x), _p(y), x.numel(), _s())


def tanh_bwd(dy, y, dx=None, dxb=None):
    _lib.call("rb_tanh_bwd", _p(dy), _p(y), _p(dx), _p(dxb), y.numel(), _s())


def attn_small_fwd(Q, K, V, mask, O, P, B, H, S, scale, drop=None):
    _lib.call("rb_attn_small_fwd", _p(Q), _p(K), _p(V), _p(mask), _p(O), _p(P), B, H, 64, S, Q.stride(0), K.stride(0), V.stride(0), O.stride(0), scale,
              _dp(drop), _s())


def attn_small_bwd(Q, K, V, dO, P, dQ, dK, dV, B, H, S, scale, drop=None):
    _lib.call("rb_attn_small_bwd", _p(Q), _p(K), _p(V), _p(dO), _p(P), _p(dQ), _p(dK), _p(dV), B, H, 64, S, Q.stride(0), K.stride(0), V.stride(0),
              dO.stride(0), dQ.stride(0), dK.stride(0), dV.stride(0), scale, _dp(drop), _s())


def box_loss(boxes, tgt, valid, inv_norm, inv_norm_dev, losses, dl1, dgiou):
    n_layers, N = boxes.shape[0], boxes.shape[1]
    _lib.call("rb_box_loss", _p(boxes), _p(tgt), _p(valid), n_layers, N, float(inv_norm), _p(inv_norm_dev), _p(losses), _p(dl1), _p(dgiou), _s())


# ---------------------------------------------------------------------------------------------------------------
# optimizer step on flat buffers (reftr_b200/optim.py)
# ---------------------------------------------------------------------------------------------------------------
def sumsq(x, out):
    """out (device scalar) += sum(x^2); x: flat fp32, numel % 4 == 0."""
    _lib.call("rb_sumsq", _p(x), x.numel(), _p(out), _s())


def scale_copy_check(src, dst, scale, flag):
    """dst = src * scale; flag (device int32[1]) |= any(src non-finite)."""
    assert src.numel() == dst.numel() and flag.dtype == torch.int32
    _lib.call("rb_scale_copy_check", _p(src), _p(dst), src.numel(), float(scale), _p(flag), _s())


def zero_if(x, flag, counter=None):
    """x = 0 and counter += 1 when flag != 0; nothing otherwise."""
    _lib.call("rb_zero_if", _p(x), x.numel(), _p(flag), _p(counter), _s())


def adamw_flat(p, g, m, v, seg_end, seg_group, lrs, wds, beta1, beta2, eps, step, sumsq_dev=None, max_norm=0.0):
    segs = _lib.AdamwSegments()
    segs.nseg = len(seg_end)
    for i, (e, grp) in enumerate(zip(seg_end, seg_group)):
        segs.end[i] = e
        segs.group[i] = grp
    for i, (lr, wd) in enumerate(zip(lrs, wds)):
        segs.lr[i] = lr
        segs.weight_decay[i] = wd
    _lib.call("rb_adamw_flat", _p(p), _p(g), _p(m), _p(v), p.numel(), C.addressof(segs), beta1, beta2, eps, int(step), _p(sumsq_dev), float(max_norm), _s())


def collate_u8(packed, table, B, H, W, mean, std, out, mask):
    _lib.call("rb_collate_u8", _p(packed), _p(table), B, H, W, float(mean[0]), float(mean[1]), float(mean[2]), float(std[0]), float(std[1]),
              float(std[2]), _p(out), _p(mask), _s())


def resize_u8(src, h, w, dst, oh, ow, tab_h, tab_v, tmp):
    """tab_* = (bounds int32 [out, 2], kk int32 [out, ksize]) on the device, or None for a pass that does not change the size."""
    bh, kh = tab_h if tab_h is not None else (None, None)
    bv, kv = tab_v if tab_v is not None else (None, None)
    _lib.call("rb_resize_u8", _p(src), h, w, _p(dst), oh, ow, _p(bh), _p(kh), kh.shape[1] if kh is not None else 0, _p(bv), _p(kv),
              kv.shape[1] if kv is not None else 0, _p(tmp), _s())
