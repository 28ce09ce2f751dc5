#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tools/nccl_probe.py 2>&1 | grep -E "MB|rror" | tee gpurun_out/nccl_probe_n2.log
nvidia-smi topo -m | head -8
