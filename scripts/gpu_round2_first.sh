#!/bin/bash
# First GPU call of round 2 (1 GPU, ~6 min): confirm the tree is green on the box, then run the
# experiment that decides the shape of the dense_tc_bwd rewrite.
#   gpurun --timeout 600 -- scripts/gpu_round2_first.sh
mkdir -p gpurun_out
if [ ! -x scripts/micro/mn_major_probe ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cirkit_b200/csrc -I include \
       -o scripts/micro/mn_major_probe scripts/micro/mn_major_probe.cu
fi
timeout 60 scripts/micro/mn_major_probe > gpurun_out/r02_mn_major_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02_mn_major_probe.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r02_pytest_gpu.log
CKB_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zzz_dense128.py -m gpu -q -x > gpurun_out/r02_dense128.log 2>&1; echo "dense128 rc=$?"; tail -15 gpurun_out/r02_dense128.log
CKB_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zzz_complex_kernels.py -m gpu -q > gpurun_out/r02_complex_kernels.log 2>&1; echo "complex kernels rc=$?"; tail -15 gpurun_out/r02_complex_kernels.log
timeout 400 python bench.py > gpurun_out/r02_bench_cp.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r02_bench_cp.log | cut -c1-200
# K = 128: FP32 SIMT route vs the experimental tcgen05 kernels (only meaningful if the dense128 test above passed)
timeout 400 python bench.py --workload pd32_cp_k128 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r02_bench_pd32_simt.log 2>&1; tail -1 gpurun_out/r02_bench_pd32_simt.log | cut -c1-200
timeout 400 python bench.py --workload pd32_cp_k128 --no-cpu-baseline --steps 5 --warmup 3 --tc-flags 515 > gpurun_out/r02_bench_pd32_tc128.log 2>&1; tail -1 gpurun_out/r02_bench_pd32_tc128.log | cut -c1-200
