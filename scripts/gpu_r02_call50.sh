#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check_final.log 2>&1; echo "dist rc=$?"; grep -i "dist overlap\|dist nvls\|error" gpurun_out/r02_dist_check_final.log | tail -5 | cut -c1-200
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --allreduce nvls > gpurun_out/r02_final_n2_nvls.log 2>&1; tail -1 gpurun_out/r02_final_n2_nvls.log | cut -c1-160
