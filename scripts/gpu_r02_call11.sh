#!/bin/bash
mkdir -p gpurun_out
{
echo "=== stress perturb"; timeout 300 python scripts/stress_bwd3b.py 2048 100 1 2>&1 | grep -v "^frame" | tail -12
echo "=== stress no perturb"; timeout 300 python scripts/stress_bwd3b.py 2048 100 0 2>&1 | grep -v "^frame" | tail -12
} > gpurun_out/r02_stress_bwd3b.txt 2>&1; cut -c1-300 gpurun_out/r02_stress_bwd3b.txt
