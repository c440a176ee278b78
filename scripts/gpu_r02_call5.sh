#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/timeline_bwd2.py 2048 1 > gpurun_out/r02_timeline_bwd2.txt 2>&1; cat gpurun_out/r02_timeline_bwd2.txt
timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu5.log | cut -c1-300
