#!/bin/bash
# 2 GPUs: staged-backward tests, multi-GPU parity incl. overlapped all-reduce, bench N=2 overlapped vs serial
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_staged_backward.py tests/test_gpu_parity.py tests/test_gpu_zz_uneven_units.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu25.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu25.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check.log 2>&1; echo "dist rc=$?"; tail -12 gpurun_out/r02_dist_check.log | cut -c1-300
for mode in 4 0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --grad-chunks $mode > gpurun_out/r02_bench_n2_chunks$mode.log 2>&1; echo "bench chunks=$mode rc=$?"; tail -1 gpurun_out/r02_bench_n2_chunks$mode.log | cut -c1-250
done
