#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
for cfg in "-1 1" "0 0" "-1 0" "-1 1"; do
set -- $cfg
REFTR_B200_MAIN_PRIORITY=$1 REFTR_B200_NCCL_HIGH_PRIORITY=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}_p.json 2> gpurun_out/r02_bench_n${N}_p.err; echo "main priority $1, NCCL high-priority group $2: rc=$?"
tail -1 gpurun_out/r02_bench_n${N}_p.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','windows_ms_per_step')}, 'e2e', round(d['e2e']['value'],1), d['e2e']['windows_ms_per_step'])" || tail -5 gpurun_out/r02_bench_n${N}_p.err
done
