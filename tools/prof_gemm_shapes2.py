"""Representative rb_gemm launches for `ncu --set full`: what bounds the L2-bound shapes (operand stream vs shared-memory bandwidth vs issue)?
Order: NT layer3 conv3 (+res), NT layer4 conv3 (+res), NT layer3 3x3, TN layer3 3x3 wgrad (auto), TN layer3 1x1 wgrad (auto), NT layer1 conv3 (+res)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reftr_b200 import ops
T16 = ops.t16(); dev = "cuda"
def nt(M, N, K, taps=1, res=False, Wp=42):
    A = torch.randn(M + 2048, K, device=dev).to(T16)[1024:1024 + M]; W = torch.randn(N, K * taps, device=dev).to(T16)
    bias = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev, dtype=T16)
    r = torch.randn(M, N, device=dev).to(T16) if res else None
    tp = [((t // 3 - 1) * Wp + (t % 3 - 1), t * K) for t in range(taps)] if taps > 1 else [(0, 0)]
    for _ in range(2):
        ops.gemm(A, W, M, N, K, taps=tp, bias=bias, res=r, relu=True, out=out)
def tn(R, Mo, No, taps=1):
    dY = torch.randn(R + 2048, Mo, device=dev).to(T16)[1024:1024 + R]; X = torch.randn(R + 2048, No, device=dev).to(T16)[1024:1024 + R]
    out = torch.zeros(Mo, taps * No, device=dev)
    tp = [(0, (t // 3 - 1) * 42 + (t % 3 - 1)) for t in range(taps)] if taps > 1 else [(0, 0)]
    for _ in range(2):
        ops.gemm(dY, X, Mo, No, R, mode=1, taps=tp, out32=out, atomic=True, splits=0, out32_z_stride=No)
B = 16
R1, R3, R4 = B * 162 * 162, B * 42 * 42, B * 22 * 22
nt(R3, 1024, 256, res=True); nt(R4, 2048, 512, res=True); nt(R3, 256, 256, taps=9); tn(R3, 256, 256, taps=9); tn(R3, 1024, 256); nt(R1, 256, 64, res=True, Wp=162)
torch.cuda.synchronize()
