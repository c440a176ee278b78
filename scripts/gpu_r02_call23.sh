#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu23.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu23.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r02_prof_cp23.json > gpurun_out/r02_bench_cp23.log 2>&1; tail -1 gpurun_out/r02_bench_cp23.log | cut -c1-300
