"""Times rb_gemm on the conv / linear shapes of the cfg2 step (CUDA events, L2-cold by rotating buffers is not needed: operands >> L2 for the big ones)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reftr_b200 import ops
T16 = ops.t16()
dev = "cuda"
def run(name, fn, flops, bytes_):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{name:44s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {bytes_ / us / 1e3:7.0f} GB/s", flush=True)

def nt(M, N, K, taps=1, res=False, mask=False, relu=True, f32=False, res32=False):
    A = torch.randn(M + 2048, K, device=dev).to(T16)[1024:1024 + M]
    W = torch.randn(N, K * taps, device=dev).to(T16)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else T16)
    r = torch.randn(M, N, device=dev).to(T16) if res else None
    m = torch.randn(M, N, device=dev).to(T16) if mask else None
    r32 = torch.randn(M, N, device=dev) if res32 else None
    tp = [((t // 3 - 1) * 162 + (t % 3 - 1), t * K) for t in range(taps)] if taps > 1 else [(0, 0)]
    kw = dict(out32=out) if f32 else dict(out=out)
    fn = lambda: ops.gemm(A, W, M, N, K, taps=tp, bias=bias, res=r, mask_src=m, res32=r32, relu=relu, **kw)
    by = M * K * 2 + M * N * (4 if f32 else 2) + (M * N * 2 if res else 0) + (M * N * 2 if mask else 0) + (M * N * 4 if res32 else 0)
    run(f"NT M{M} N{N} K{K} t{taps} res{int(res)} mask{int(mask)} f32{int(f32)}", fn, 2.0 * M * N * K * taps, by)

def tn(R, Mo, No, taps=1, splits=1):
    dY = torch.randn(R + 2048, Mo, device=dev).to(T16)[1024:1024 + R]
    X = torch.randn(R + 2048, No, device=dev).to(T16)[1024:1024 + R]
    out = torch.zeros(Mo, taps * No, device=dev)
    tp = [(0, (t // 3 - 1) * 42 + (t % 3 - 1)) for t in range(taps)] if taps > 1 else [(0, 0)]
    fn = lambda: ops.gemm(dY, X, Mo, No, R, mode=1, taps=tp, out32=out, atomic=True, splits=splits, out32_z_stride=No)
    run(f"TN R{R} M{Mo} N{No} t{taps} s{splits}", fn, 2.0 * R * Mo * No * taps, R * (Mo + No) * 2)

B = 16
R1, R2, R3, R4 = B * 162 * 162, B * 82 * 82, B * 42 * 42, B * 22 * 22
nt(R1, 64, 64); nt(R1, 64, 256); nt(R1, 64, 64, taps=9); nt(R1, 256, 64, relu=False); nt(R1, 256, 64, res=True)
nt(R1, 128, 256)
nt(R2, 128, 128, taps=9); nt(R2, 512, 128, res=True); nt(R2, 128, 512); nt(R2, 512, 256, relu=False)
nt(R3, 256, 256, taps=9); nt(R3, 1024, 256, res=True); nt(R3, 256, 1024)
nt(R4, 512, 512, taps=9); nt(R4, 2048, 512, res=True); nt(R4, 512, 2048)
# dgrad-style (mask)
nt(R3, 256, 1024, mask=True, relu=False); nt(R3, 1024, 256, res=True, mask=True, relu=False); nt(R2, 128, 512, mask=True, relu=False)
# transformer
T = B * 420
nt(T, 512, 256, relu=False); nt(T, 256, 256, relu=False, f32=True, res32=True); nt(T, 2048, 256); nt(T, 256, 2048, relu=False, f32=True, res32=True)
nt(T, 256, 768, relu=False, f32=True, res32=True)
# wgrad
tn(R3, 256, 256, taps=9, splits=13); tn(R3, 256, 1024, splits=28); tn(R3, 1024, 256, splits=28); tn(R2, 128, 128, taps=9, splits=12)
tn(R2, 512, 128, splits=56); tn(R4, 512, 512, taps=9, splits=7); tn(T, 2048, 256, splits=14); tn(T, 256, 2048, splits=14); tn(T, 256, 256, splits=14)
# round 2: splits = 0 -> tile width and K splits chosen together by rb_gemm (single-wave configurations)
print("--- auto splits")
tn(R3, 256, 256, taps=9, splits=0); tn(R3, 256, 1024, splits=0); tn(R3, 1024, 256, splits=0); tn(R2, 128, 128, taps=9, splits=0)
tn(R2, 512, 128, splits=0); tn(R2, 128, 512, splits=0); tn(R4, 512, 512, taps=9, splits=0); tn(R4, 512, 2048, splits=0); tn(T, 2048, 256, splits=0); tn(T, 256, 2048, splits=0); tn(T, 256, 256, splits=0)
tn(320, 768, 768, splits=0); tn(320, 3072, 768, splits=0); tn(320, 768, 3072, splits=0)
