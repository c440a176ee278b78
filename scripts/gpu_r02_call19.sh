#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_complex.py -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu19.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02_pytest_gpu19.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --profile-out gpurun_out/r02_prof_cp_bwd3b.json > gpurun_out/r02_bench_cp_bwd3b.log 2>&1; tail -1 gpurun_out/r02_bench_cp_bwd3b.log | cut -c1-200
