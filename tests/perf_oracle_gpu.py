"""Times the ORACLE (plain PyTorch fp32 restatement == the reference's stock PyTorch-CUDA path) on the GPU: the
denominator of the north-star "5x" target.  Not a pytest; run under gpurun:  python tests/perf_oracle_gpu.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformers import BertConfig, BertModel
from oracle.reftr_oracle import RefTROracle, total_box_loss
from reftr_b200.synthetic import synthetic_samples, synthetic_targets, synthetic_weights

B, H, W, L = 16, 640, 640, 20
dev = "cuda"
torch.manual_seed(0)
model = RefTROracle(BertModel(BertConfig()), dropout=0.1, aux_loss=True)
synthetic_weights(model, 0)
model.to(dev)
s = synthetic_samples(B, H, W, L, device=dev)
tgt = synthetic_targets(B, device=dev)
res = {}
for mode in ("eval", "train"):
    model.train(mode == "train")
    for tf32 in (True,):
        torch.backends.cudnn.allow_tf32 = tf32  # stock default: TF32 conv on, fp32 matmul
        def step():
            model.zero_grad(set_to_none=True)
            out = model(s)
            loss = total_box_loss(out, tgt)
            loss.backward()
            return loss
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 15
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[f"{mode}_fwdbwd_ms"] = ms
        res[f"{mode}_samples_per_s"] = B / ms * 1e3
        print(mode, "fwd+bwd ms", ms, "samples/s", B / ms * 1e3, flush=True)
# forward only
model.eval()
with torch.no_grad():
    for _ in range(3):
        model(s)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(10):
        model(s)
    torch.cuda.synchronize()
    res["eval_fwd_ms"] = (time.time() - t0) / 10 * 1e3
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/oracle_gpu_baseline.json", "w"))
