#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu53.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu53.log | cut -c1-200
timeout 200 python bench.py --profile-out gpurun_out/final3_prof_cp.json > gpurun_out/final3_bench_cp.log 2>&1; tail -1 gpurun_out/final3_bench_cp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step']['frac'],3), 'cpu', round(d['cpu_baseline']['value']))"
