"""Breakdown of the epilogue cost (RB_GEMM_DEBUG bits: 1 no stores, 4 no math, 8 no TMEM loads, 16 no staging writes, 32 no fence/barrier)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reftr_b200 import ops
T16 = ops.t16()
dev = "cuda"
def run(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:72s} {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us", flush=True)
def nt(M, N, K, res, dbg, bias=True, relu=True, tag=""):
    A = torch.randn(M, K, device=dev).to(T16); W = torch.randn(N, K, device=dev).to(T16)
    b = torch.randn(N, device=dev) if bias else None
    out = torch.empty(M, N, device=dev, dtype=T16)
    r = torch.randn(M, N, device=dev).to(T16) if res else None
    os.environ["RB_GEMM_DEBUG"] = str(dbg)
    run(f"M{M} N{N} K{K} res{int(res)} bias{int(bias)} relu{int(relu)} dbg{dbg:2d} {tag}", lambda: ops.gemm(A, W, M, N, K, bias=b, res=r, relu=relu, out=out))
    os.environ["RB_GEMM_DEBUG"] = "0"
B = 16
R1, R3 = B * 162 * 162, B * 42 * 42
for (M, N, K) in ((R1, 256, 64), (R3, 1024, 256)):
    for res in (False, True):
        nt(M, N, K, res, 0, tag="full")
        nt(M, N, K, res, 8, tag="no TMEM ld")
        nt(M, N, K, res, 16, tag="no staging writes")
        nt(M, N, K, res, 32, tag="no fence/barrier")
        nt(M, N, K, res, 1 | 16, tag="no stores, no staging")
        nt(M, N, K, res, 1 | 8 | 16, tag="no stores, no staging, no TMEM ld (ALU only)")
        nt(M, N, K, res, 1 | 8 | 16 | 32, tag="... and no barrier")
        nt(M, N, K, res, 0, bias=False, relu=False, tag="full, no bias/relu")
        nt(M, N, K, res, 5, tag="no math no stores (loads + sync)")
