"""GPU-side cost of small rb_gemm launches, measured inside a CUDA graph (no host launch overhead)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reftr_b200 import ops
T16 = ops.t16()
dev = "cuda"
def timed_graph(fn, n=40):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def nt(M, N, K, f32=False, res32=False, relu=False):
    A = torch.randn(M, K, device=dev).to(T16); W = torch.randn(N, K, device=dev).to(T16); bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else T16)
    r32 = torch.randn(M, N, device=dev) if res32 else None
    kw = dict(out32=out) if f32 else dict(out=out)
    us = timed_graph(lambda: ops.gemm(A, W, M, N, K, bias=bias, res32=r32, relu=relu, **kw))
    print(f"NT M{M} N{N} K{K} f32{int(f32)} res32{int(res32)}: {us:.1f} us  {2.0*M*N*K/us/1e6:.1f} TF/s", flush=True)
def tn(R, Mo, No, splits):
    dY = torch.randn(R, Mo, device=dev).to(T16); X = torch.randn(R, No, device=dev).to(T16); out = torch.zeros(Mo, No, device=dev)
    us = timed_graph(lambda: ops.gemm(dY, X, Mo, No, R, mode=1, out32=out, atomic=True, splits=splits))
    print(f"TN R{R} M{Mo} N{No} s{splits}: {us:.1f} us  {2.0*R*Mo*No/us/1e6:.1f} TF/s", flush=True)
nt(16, 256, 256, f32=True); nt(320, 768, 768, f32=True, res32=True); nt(320, 2304, 768); nt(320, 3072, 768); nt(320, 768, 3072, f32=True, res32=True)
nt(6720, 512, 256); nt(6720, 256, 256, f32=True, res32=True); nt(6720, 2048, 256, relu=True); nt(6720, 256, 2048, f32=True, res32=True); nt(6720, 1536, 256)
tn(320, 768, 768, 1); tn(320, 3072, 768, 1); tn(320, 768, 3072, 1); tn(6720, 256, 256, 14); tn(6720, 2048, 256, 14); tn(6720, 512, 256, 14); tn(16, 256, 256, 1)
x = torch.randn(6720, 256, device=dev).to(T16); o = torch.zeros(256, device=dev)
print("colsum bf16 6720x256:", timed_graph(lambda: ops.colsum(x, o)), "us")
x2 = torch.randn(320, 3072, device=dev).to(T16); o2 = torch.zeros(3072, device=dev)
print("colsum bf16 320x3072:", timed_graph(lambda: ops.colsum(x2, o2)), "us")
