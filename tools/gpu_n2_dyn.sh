#!/bin/bash
# N GPUs: dynamic-scheduler build vs the previous (static) kernels, same box; N=1 for the efficiency
N=${1:-2}
run() { name=$1; lib=$2; shift 2
  env REFTR_B200_LIB=$lib "$@" REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 --windows 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name N=$N', round(d['value'],1), d['windows_ms_per_step'])"
}
run1() { name=$1; lib=$2; shift 2
  env REFTR_B200_LIB=$lib "$@" REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 timeout 300 python bench.py --steps 20 --warmup 5 --windows 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name N=1', round(d['value'],1), d['windows_ms_per_step'])"
}
run base $PWD/build/base/libreftr_b200.so X=1
run dyn $PWD/reftr_b200/libreftr_b200.so X=1
run base $PWD/build/base/libreftr_b200.so X=1
run dyn $PWD/reftr_b200/libreftr_b200.so X=1
run1 base $PWD/build/base/libreftr_b200.so X=1
run1 dyn $PWD/reftr_b200/libreftr_b200.so X=1
