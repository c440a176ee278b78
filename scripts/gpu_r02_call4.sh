#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 > gpurun_out/r02_pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu4.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --profile-out gpurun_out/r02_prof_cp_bwd2.json > gpurun_out/r02_bench_cp_bwd2.log 2>&1; tail -1 gpurun_out/r02_bench_cp_bwd2.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --tc-flags 1539 --profile-out gpurun_out/r02_prof_cp_bwd1.json > gpurun_out/r02_bench_cp_bwd1.log 2>&1; tail -1 gpurun_out/r02_bench_cp_bwd1.log | cut -c1-300
