"""per-k-iteration cost of the GEMM main loop: same problem at different tile widths / K (inside a CUDA graph)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reftr_b200 import ops
T16 = ops.t16()
dev = "cuda"
def timed_graph(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for M, N in ((320, 768), (128, 64), (128, 256), (18944, 64), (18944, 256)):
    for K in (256, 1024, 3072):
        A = torch.randn(M, K, device=dev).to(T16); W = torch.randn(N, K, device=dev).to(T16)
        out = torch.empty(M, N, device=dev, dtype=T16)
        for bn in (64, 128, 256):
            if bn > N and bn > 64: continue
            us = timed_graph(lambda: ops.gemm(A, W, M, N, K, out=out, block_n=bn))
            tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
            per_sm = -(-tiles // 148)
            print(f"M{M} N{N} K{K} bn{bn}: {us:6.1f} us  tiles {tiles} ({per_sm}/SM)  k-iters/SM {per_sm * K // 64}  -> {us / (per_sm * K / 64) * 1e3:.0f} ns per k-iter", flush=True)
