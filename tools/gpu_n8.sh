#!/bin/bash
# N GPUs of one box: the driver's scaling line (bench.py under torchrun) with --verify
N=${1:-8}
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
python bench.py --no-cpu-baseline --windows 3 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=1:', round(d['value'],1), d['windows_ms_per_step'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --verify > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"
tail -1 gpurun_out/r02_bench_n$N.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','windows_ms_per_step')}); print(d['e2e']['value']); v=d.get('verify') or {}; print({k:v.get(k) for k in ('world','rel_l2_flat_gradient','worst_tensor_rel_l2','ok')})" || tail -20 gpurun_out/r02_bench_n$N.err
