"""Times the attention kernels at the cfg2 encoder shape (B=16, H=8, S=420) with CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reftr_b200 import ops
T16 = ops.t16()
B, H, S = 16, 8, 420
d = 256
dev = "cuda"
qkv = torch.randn(B * S, 3 * d, device=dev).to(T16)
kpm = torch.zeros(B, S, dtype=torch.uint8, device=dev)
o = torch.empty(B * S, d, device=dev, dtype=T16)
do = torch.randn(B * S, d, device=dev).to(T16)
dqkv = torch.empty_like(qkv)
lse = torch.empty(B, H, S, device=dev); Dbuf = torch.empty(B, H, S, device=dev)
q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
def fwd(): ops.attn_fwd(q, k, v, kpm, o, lse, B, H, S, S, 32 ** -0.5)
def bwd(): ops.attn_bwd(q, k, v, kpm, o, do, lse, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], Dbuf, B, H, S, S, 32 ** -0.5)
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"attn {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
# train mode: dropout 0.1 on the probabilities (counter-based hash inside the kernels)
seed = torch.full((1,), 12345, dtype=torch.int64, device=dev)
dr = ops.Drop(seed, "perf.attn", 0.1)
def fwd_t(): ops.attn_fwd(q, k, v, kpm, o, lse, B, H, S, S, 32 ** -0.5, drop=dr)
def bwd_t(): ops.attn_bwd(q, k, v, kpm, o, do, lse, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], Dbuf, B, H, S, S, 32 ** -0.5, drop=dr)
for name, fn in (("fwd train", fwd_t), ("bwd train", bwd_t)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"attn {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
