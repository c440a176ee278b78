#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_staged_backward.py tests/test_gpu_dropin.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu24.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu24.log | cut -c1-300
