#!/bin/bash
# (now: REFTR_B200_SPLIT_BERT_HALVES 1 / 0 -- BERT backward as two graphs / one)
# 2 GPUs: split backward with BERT's part launched early (beside layer4) vs late (round 2a: after layer4), + the N=1 step on the same box
N=${1:-2}
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
timeout 900 python -m pytest tests/test_e2e_gpu.py -x -q -k "split" > gpurun_out/r02_pytest_split.log 2>&1; tail -2 gpurun_out/r02_pytest_split.log
python bench.py --no-cpu-baseline --windows 3 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=1:', round(d['value'],1), d['windows_ms_per_step'])"
for early in 1 0 1; do
REFTR_B200_SPLIT_BERT_HALVES=$early timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$early bench.py --gpus $N --steps 20 --warmup 5 --windows 3 --verify > gpurun_out/r02_bench_n${N}_early$early.json 2> gpurun_out/r02_bench_n${N}_early$early.err; echo "bench N=$N early=$early rc=$?"
tail -1 gpurun_out/r02_bench_n${N}_early$early.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','windows_ms_per_step')}); print(d['e2e']['value']); print(d.get('verify'))" || tail -20 gpurun_out/r02_bench_n${N}_early$early.err
done
