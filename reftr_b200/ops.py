"""Python-side launchers for the C-ABI kernels (include/reftr_b200.h).  torch is used for device memory and the
current stream only; all arithmetic happens in libreftr_b200.so."""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, Geom


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def make_geom(mode=0, Wp=0, HpWp=0, H=0, W=0, Rs=0):
    return Geom(mode, Wp, HpWp, H, W, Rs)


def _check_2d(t, dtype, name):
    assert t.is_cuda and t.dtype == dtype and t.dim() == 2 and t.stride(1) == 1, (name, t.dtype, t.shape, t.stride())


def gemm(A, B, M, N, K, *, mode=0, taps=((0, 0),), bias=None, res=None, res32=None, mask_src=None, relu=False,
         out=None, out32=None, atomic=False, splits=1, geom=None, out_row_off=0, out32_z_stride=0, block_n=0):
    """See rb_gemm in include/reftr_b200.h.  ``taps`` is a sequence of (a_rowoff, b_koff) pairs."""
    _check_2d(A, torch.bfloat16, "A")
    _check_2d(B, torch.bfloat16, "B")
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
        _check_2d(res, torch.bfloat16, "res")
        a.res, a.ldres = res.data_ptr(), res.stride(0)
    if res32 is not None:
        _check_2d(res32, torch.float32, "res32")
        a.res32, a.ldres32 = res32.data_ptr(), res32.stride(0)
    if mask_src is not None:
        _check_2d(mask_src, torch.bfloat16, "mask_src")
        a.mask_src, a.ldmask = mask_src.data_ptr(), mask_src.stride(0)
    if out is not None:
        _check_2d(out, torch.bfloat16, "out")
        a.out, a.ldo = out.data_ptr(), out.stride(0)
    if out32 is not None:
        assert out32.dtype == torch.float32 and out32.is_cuda
        a.out32, a.ldo32 = out32.data_ptr(), (out32.stride(-2) if out32.dim() >= 2 else N)
    a.out32_z_stride = out32_z_stride
    a.relu = int(relu)
    a.atomic = int(atomic)
    if geom is not None:
        a.geom = geom
    _lib.check(_lib.lib().rb_gemm(C.byref(a), _stream()), "rb_gemm")
    return out if out is not None else out32
