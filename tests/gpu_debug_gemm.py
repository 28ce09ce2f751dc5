"""Diagnostic (not a pytest): prints error structure of rb_gemm for tiny structured problems."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reftr_b200 import ops

torch.manual_seed(0)


def report(name, got, ref):
    err = (got.float() - ref.float()).abs()
    print(f"{name}: max_err={err.max().item():.4g} ref_max={ref.abs().max().item():.4g} "
          f"bad_rows={(err.max(1).values > 1e-2 * ref.abs().max()).sum().item()}/{got.shape[0]} "
          f"bad_cols={(err.max(0).values > 1e-2 * ref.abs().max()).sum().item()}/{got.shape[1]}", flush=True)


for (M, N, K, bn) in [(128, 32, 64, 32), (128, 64, 64, 64), (128, 128, 64, 128), (128, 256, 64, 256), (128, 64, 128, 64),
                      (256, 64, 256, 64), (300, 96, 160, 0)]:
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    o = torch.zeros(M, N, device="cuda")
    try:
        ops.gemm(A, B, M, N, K, out32=o, block_n=bn)
        torch.cuda.synchronize()
        report(f"NT M{M} N{N} K{K} bn{bn}", o, A.float() @ B.float().t())
    except Exception as e:
        print("NT", M, N, K, bn, "EXC", e, flush=True)

for (R, Mo, No, bn, splits) in [(64, 128, 64, 64, 1), (64, 128, 128, 128, 1), (128, 128, 64, 64, 1), (256, 128, 64, 64, 2),
                                (64, 64, 64, 64, 1), (1000, 256, 256, 256, 3)]:
    dY = torch.randn(R, Mo, device="cuda").bfloat16()
    X = torch.randn(R, No, device="cuda").bfloat16()
    o = torch.zeros(Mo, No, device="cuda")
    try:
        ops.gemm(dY, X, Mo, No, R, mode=1, out32=o, atomic=True, splits=splits, block_n=bn)
        torch.cuda.synchronize()
        report(f"TN R{R} M{Mo} N{No} bn{bn} s{splits}", o, dY.float().t() @ X.float())
    except Exception as e:
        print("TN", R, Mo, No, bn, "EXC", e, flush=True)

# throughput probe
for (M, N, K, bn) in [(16384, 4096, 4096, 256), (16384, 4096, 4096, 128), (419904, 64, 576, 64), (107584, 128, 1152, 128),
                      (107584, 512, 128, 128), (28224, 256, 2304, 128), (28224, 1024, 256, 256), (7744, 512, 4608, 128)]:
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(A, B, M, N, K, out=o, block_n=bn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(A, B, M, N, K, out=o, block_n=bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"perf NT M{M} N{N} K{K} bn{bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s  "
          f"{(M*K+N*K+M*N)*2/ms/1e6:.0f} GB/s", flush=True)
