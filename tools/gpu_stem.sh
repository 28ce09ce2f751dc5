#!/bin/bash
# fused stem + max-pool (rb_stem_pool) vs the two-kernel path: parity tests, then the step A/B
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py -x -q -k "stem or cfg1_box or seg" > gpurun_out/r02_pytest_stem.log 2>&1; tail -3 gpurun_out/r02_pytest_stem.log
for rep in 1 2; do
for v in 0 1; do
  REFTR_B200_STEM_FUSED=$v timeout 300 python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_stem$v.json 2> gpurun_out/r02_bench_stem$v.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_stem$v.json") if l.startswith("{")][-1])
print("fused=$v", round(d["value"],1), round(d["e2e"]["value"],1), d["windows_ms_per_step"])
P
done
done
timeout 300 python - <<'P'
import torch, time
from reftr_b200 import ops
from reftr_b200.pack import PackedStem
B,H,W=16,640,640
H1=W1=320; H2=W2=160
img=torch.randn(B,3,H,W,device="cuda")
conv=torch.nn.Conv2d(3,64,7,stride=2,padding=3,bias=False).cuda()
st=PackedStem(conv,None,need_dgrad=False,ldk=160); st.refresh()
T=ops.t16()
hwc4=torch.empty(B*H*(W+2),4,device="cuda",dtype=T); out=torch.empty(B*(H2+2)*(W2+2),64,device="cuda",dtype=T)
c1=torch.empty(B*H1*W1,64,device="cuda",dtype=T)
flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
def t(f,n=20):
    tot=0
    for _ in range(n):
        flush.zero_(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); tot+=e0.elapsed_time(e1)
    return tot/n*1e3
f1=lambda: ops.stem_pool(img,st.wrow,st.bias,hwc4,out,B,H,W,H1,W1,H2,W2)
f2=lambda: (ops.stem_conv(img,st.wf,st.bias,c1,B,H,W,H1,W1), ops.maxpool_3x3s2(c1,out,B,H1,W1,64,H2,W2))
f1(); f2(); torch.cuda.synchronize()
print("rb_stem_pool (convert + fused): %.1f us;  rb_stem_conv + rb_maxpool_3x3s2: %.1f us" % (t(f1), t(f2)))
P
