#!/bin/bash
mkdir -p gpurun_out
{
for fl in 0 0 16384; do echo "=== bwd3 flags $fl"; timeout 200 python scripts/stress_bwd3b.py 2048 150 $fl 2>&1 | grep -v "^frame" | grep "MISMATCH\|done\|Error" | tail -6; done
} > gpurun_out/r02_stress_bwd3e.txt 2>&1; cut -c1-200 gpurun_out/r02_stress_bwd3e.txt
timeout 600 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu15.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu15.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --profile-out gpurun_out/r02_prof_cp_bwd3.json > gpurun_out/r02_bench_cp_bwd3.log 2>&1; tail -1 gpurun_out/r02_bench_cp_bwd3.log | cut -c1-300
