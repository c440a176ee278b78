#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -m gpu -q -x --timeout 120 > gpurun_out/r02_pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02_pytest_gpu7.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --profile-out gpurun_out/r02_prof_cp_bwd3.json > gpurun_out/r02_bench_cp_bwd3.log 2>&1; tail -1 gpurun_out/r02_bench_cp_bwd3.log | cut -c1-300
