#!/bin/bash
mkdir -p gpurun_out
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/allreduce_probe.py 2>&1 | grep "all_reduce"; }
run
NCCL_ALGO=NVLS run
NCCL_ALGO=Ring run
NCCL_ALGO=Tree run
NCCL_DEBUG=INFO timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/allreduce_probe.py > gpurun_out/r02_nccl_info.log 2>&1; grep -i "nvls\|algo\|channels" gpurun_out/r02_nccl_info.log | head -12 | cut -c1-200
