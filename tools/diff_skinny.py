"""Debug: which workspace buffer first differs between the skinny-GEMM path and the tcgen05 path (train mode, same seed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle.cases import CASES
from oracle.reftr_oracle import total_box_loss
from reftr_b200.synthetic import synthetic_samples, synthetic_targets
from util_build import build_candidate
os.environ["REFTR_B200_GRAPHS"] = "0"
os.environ["REFTR_B200_SIDE_STREAM"] = "0"
case = dict(CASES["cfg1_box"]); case["oracle_kw"] = dict(case["oracle_kw"], dropout=0.1)
s = synthetic_samples(**case["inputs"], device="cuda")
def run(no_skinny):
    if no_skinny: os.environ["RB_GEMM_NO_SKINNY"] = "1"
    else: os.environ.pop("RB_GEMM_NO_SKINNY", None)
    m = build_candidate(case, device="cuda").train()
    m.engine().next_seed = 0x0123456789ABCDEF
    out = m(s)
    total_box_loss(out, synthetic_targets(case["inputs"]["B"], 1, device="cuda")).backward()
    torch.cuda.synchronize()
    bufs = {k: v.detach().float().clone() for k, v in m.engine().ws.bufs.items()}
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    return bufs, grads
a, ga = run(False)
b, gb = run(True)
rows = []
for k in a:
    if k in b and a[k].shape == b[k].shape:
        d = (a[k] - b[k]).norm().item() / (b[k].norm().item() + 1e-20)
        rows.append((d, k, tuple(a[k].shape)))
for d, k, sh in rows:
    if d > 2e-3:
        print(f"{d:10.3e}  {k:28s} {sh}")
print("---- grads")
gr = sorted(((ga[n] - gb[n]).norm().item() / (gb[n].norm().item() + 1e-20), n) for n in ga)
for d, n in gr[-12:]:
    print(f"{d:10.3e}  {n}")
