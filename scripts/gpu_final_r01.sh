#!/bin/bash
# Round-end evidence run (1 GPU): tests, smoke, bench lines, ncu launch lists and full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/final_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py > gpurun_out/final_bench_cp.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/final_bench_cp.log | cut -c1-160
timeout 400 python bench.py --workload qt28_tucker_k64 --steps 10 --warmup 3 > gpurun_out/final_bench_tucker.log 2>&1; tail -1 gpurun_out/final_bench_tucker.log | cut -c1-160
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.log 2>&1; tail -1 gpurun_out/final_bench_reference.log | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_final.log 2>&1; echo "launch list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r01_tucker_final.csv \
    python bench.py --workload qt28_tucker_k64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_tucker_final.log 2>&1; echo "tucker launch list rc=$?"
scripts/gpu_ncu_one.sh r01f dense_tc_bwd_kernel 8
scripts/gpu_ncu_one.sh r01f dense_tc_fwd_kernel 1
scripts/gpu_ncu_one.sh r01f table_bwd_kernel 0
scripts/gpu_ncu_wl.sh r01f qt28_tucker_k64 tucker_tc_bwd_dx_kernel 8
scripts/gpu_ncu_wl.sh r01f qt28_tucker_k64 tucker_tc_bwd_dw_kernel 8
scripts/gpu_ncu_wl.sh r01f qt28_tucker_k64 tucker_tc_fwd_kernel 0
