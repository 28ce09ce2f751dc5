#!/bin/bash
# N GPUs of one box: bench (with --verify: all-reduced gradient vs one process on the global batch) and a kernel timeline of one step on rank 0
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --verify > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"
tail -1 gpurun_out/r02_bench_n$N.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','windows_ms_per_step')}); print(d['e2e']['value']); print(d.get('verify'))" || tail -20 gpurun_out/r02_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/timeline.py --out gpurun_out/r02_timeline_n$N > gpurun_out/r02_timeline_n$N.log 2>&1; echo "timeline rc=$?"
grep -i "nccl\|span\|busy (union" gpurun_out/r02_timeline_n${N}_summary.txt | head -12
REFTR_B200_SPLIT_BWD=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --windows 3 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('single-graph backward + one all-reduce:', d['value'], d['ms_per_step'])"
