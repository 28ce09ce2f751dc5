"""One eager (no CUDA graph) fwd+bwd step of the bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Usage: python tools/profile_step.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("REFTR_B200_GRAPHS", "0")
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
model, crit, _ = bench.build_ours(dev)
model.train() if os.environ.get("REFTR_B200_BENCH_EVAL") != "1" else model.eval()
s, t = bench.host_batch(B, pinned=False)
s, t = bench.to_device(s, t, dev)

def step():
    model.zero_grad(set_to_none=True)
    ld = crit(model(s), bench.targets_list(t))
    sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict).backward()

for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step, C-ABI launches per step:", model.engine().launches // 3)
