"""A few representative rb_gemm launches for `ncu --set full` (stall reasons of the epilogue-bound and the tensor-bound shapes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reftr_b200 import ops
T16 = ops.t16(); dev = "cuda"
def nt(M, N, K, taps=1, res=False):
    A = torch.randn(M + 2048, K, device=dev).to(T16)[1024:1024 + M]; W = torch.randn(N, K * taps, device=dev).to(T16)
    bias = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev, dtype=T16)
    r = torch.randn(M, N, device=dev).to(T16) if res else None
    tp = [((t // 3 - 1) * 162 + (t % 3 - 1), t * K) for t in range(taps)] if taps > 1 else [(0, 0)]
    for _ in range(2):
        ops.gemm(A, W, M, N, K, taps=tp, bias=bias, res=r, relu=True, out=out)
B = 16
nt(B * 162 * 162, 256, 64, res=True)      # layer1 conv3 + residual: epilogue / HBM bound
nt(B * 42 * 42, 1024, 256, res=True)      # layer3 conv3 + residual
nt(B * 42 * 42, 256, 256, taps=9)         # layer3 3x3: tensor bound
nt(B * 420, 2048, 256)                    # encoder FFN linear1
torch.cuda.synchronize()
