#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu20.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu20.log | cut -c1-300
timeout 300 python bench.py --workload rbt64_sos_k64 --steps 20 --warmup 5 --profile-out gpurun_out/r02_prof_sos.json > gpurun_out/r02_bench_sos.log 2>&1; tail -2 gpurun_out/r02_bench_sos.log | cut -c1-1500
timeout 300 python bench.py --workload qt28_cp_k32 --batch 512 --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02_bench_k32b.log 2>&1; tail -1 gpurun_out/r02_bench_k32b.log | cut -c1-200
