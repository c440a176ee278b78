#!/bin/bash
# round 2, call 2: layout probes for the backward rewrite, new parity tests, reference arm from baseline/_ref
mkdir -p gpurun_out
for i in 0 1 2 3 4 5 6 7; do timeout 60 scripts/micro/umma_probe2 $i; done > gpurun_out/r02_umma_probe2.txt 2>&1
cat gpurun_out/r02_umma_probe2.txt
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r02_pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02_pytest_gpu2.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_pytest_gpu2_all.log 2>&1; echo "pytest(all) rc=$?"; tail -40 gpurun_out/r02_pytest_gpu2_all.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r02_bench_ref.log | cut -c1-900
