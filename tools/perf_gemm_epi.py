"""Where does the epilogue time of the output-heavy 1x1 convolutions go?  (RB_GEMM_DEBUG knobs: timing only, results wrong.)"""
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
    print(f"{name:64s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {bytes_ / us / 1e3:7.0f} GB/s", flush=True)
def nt(M, N, K, res=False, relu=True, bias=True, bn=0, dbg=0, f32=False, tag=""):
    A = torch.randn(M, K, device=dev).to(T16); W = torch.randn(N, K, device=dev).to(T16)
    b = torch.randn(N, device=dev) if bias else None
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else T16)
    r = torch.randn(M, N, device=dev).to(T16) if res else None
    kw = dict(out32=out) if f32 else dict(out=out)
    os.environ["RB_GEMM_DEBUG"] = str(dbg)
    fn = lambda: ops.gemm(A, W, M, N, K, bias=b, res=r, relu=relu, block_n=bn, **kw)
    by = M * K * 2 + M * N * (4 if f32 else 2) + (M * N * 2 if res else 0)
    run(f"M{M} N{N} K{K} res{int(res)} bias{int(bias)} bn{bn} dbg{dbg} f32{int(f32)} {tag}", fn, 2.0 * M * N * K, by)
    os.environ["RB_GEMM_DEBUG"] = "0"
B = 16
R1, R2, R3, R4 = B * 162 * 162, B * 82 * 82, B * 42 * 42, B * 22 * 22
for (M, N, K) in ((R3, 1024, 256), (R1, 256, 64), (R2, 512, 128)):
    for res in (False, True):
        for bn in (0, 64, 128, 256):
            nt(M, N, K, res=res, bn=bn)
    nt(M, N, K, res=True, bias=False)
    nt(M, N, K, res=True, dbg=1, tag="no-store")
    nt(M, N, K, res=True, dbg=4, tag="no-math")
    nt(M, N, K, res=True, dbg=5, tag="no-store no-math")
    nt(M, N, K, res=False, dbg=5, tag="no-store no-math")
    nt(M, N, K, res=False, dbg=1, tag="no-store")
# copy-bandwidth reference points on this box
x = torch.empty(R1, 256, device=dev, dtype=T16); y = torch.empty_like(x)
run("torch copy 215 MB -> 215 MB", lambda: y.copy_(x), 0, 2 * x.numel() * 2)
z = torch.empty_like(x)
run("torch add (2 reads, 1 write) 215 MB each", lambda: torch.add(x, y, out=z), 0, 3 * x.numel() * 2)
