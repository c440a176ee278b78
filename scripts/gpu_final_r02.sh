#!/bin/bash
# Round-2 evidence run (1 GPU): tests, smoke, bench lines, ncu launch list and full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/final2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/final2_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final2_smoke.log | cut -c1-200
timeout 400 python bench.py --profile-out gpurun_out/final2_prof_cp.json > gpurun_out/final2_bench_cp.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/final2_bench_cp.log | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final2_bench_reference.log 2>&1; tail -1 gpurun_out/final2_bench_reference.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_r02.log 2>&1; echo "launch list rc=$?"
scripts/gpu_ncu_one.sh r02 dense_tc_bwd3_kernel 18
scripts/gpu_ncu_one.sh r02 dense_tc_fwd_kernel 12
scripts/gpu_ncu_one.sh r02 table_bwd_kernel 1
scripts/gpu_ncu_one.sh r02 table_pair_gather_smem_kernel 1
ls -la gpurun_out | grep "r02\|final2" | tail -12
