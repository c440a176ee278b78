#!/bin/bash
mkdir -p gpurun_out
{
echo "=== stress 2048"; timeout 300 python scripts/stress_bwd3.py 2048 60 2>&1 | grep -v "^frame" | tail -15
echo "=== bench blocking"; CUDA_LAUNCH_BLOCKING=1 timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>&1 | grep -v "^frame" | tail -12 | cut -c1-400
} > gpurun_out/r02_stress_bwd3.txt 2>&1; cut -c1-400 gpurun_out/r02_stress_bwd3.txt
