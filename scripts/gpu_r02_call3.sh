#!/bin/bash
mkdir -p gpurun_out
timeout 60 scripts/micro/tmem_shape_probe > gpurun_out/r02_tmem_shape_probe.txt 2>&1; cat gpurun_out/r02_tmem_shape_probe.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu3.log | cut -c1-300
timeout 400 python bench.py --workload pd32_cp_k128 --no-cpu-baseline --steps 5 --warmup 3 --profile-out gpurun_out/r02_prof_pd32_tc128.json > gpurun_out/r02_bench_pd32_b.log 2>&1; tail -1 gpurun_out/r02_bench_pd32_b.log | cut -c1-300
timeout 400 python bench.py --workload qt28_cp_k32 --batch 512 --no-cpu-baseline --steps 20 --warmup 5 --profile-out gpurun_out/r02_prof_k32_b512.json > gpurun_out/r02_bench_k32.log 2>&1; tail -1 gpurun_out/r02_bench_k32.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --profile-out gpurun_out/r02_prof_cp.json > gpurun_out/r02_bench_cp_b.log 2>&1; tail -1 gpurun_out/r02_bench_cp_b.log | cut -c1-300
