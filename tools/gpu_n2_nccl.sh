#!/bin/bash
# NCCL protocol / CTA-count variants of the N-GPU bench (same box), then N=1 on the same box for the efficiency
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --windows 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), d['windows_ms_per_step'])"
}
run default X=1
run simple NCCL_PROTO=Simple
run maxctas8 NCCL_MAX_CTAS=8
run simple_maxctas8 NCCL_PROTO=Simple NCCL_MAX_CTAS=8
run simple_maxctas4 NCCL_PROTO=Simple NCCL_MAX_CTAS=4
run nosplit REFTR_B200_SPLIT_BWD=0
REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 timeout 300 python bench.py --steps 20 --warmup 5 --windows 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 same box', round(d['value'],1), d['windows_ms_per_step'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --windows 1 --verify 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('verify'))"
