#!/bin/bash
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/ddp_host_probe.py 2>&1 | grep -E "host ms|rror"
