#!/bin/bash
mkdir -p gpurun_out
{
echo "=== stress with ring check"; timeout 300 python scripts/stress_bwd3b.py 2048 80 0 2>&1 | grep -v "^frame" | tail -14
} > gpurun_out/r02_stress_bwd3c.txt 2>&1; cut -c1-300 gpurun_out/r02_stress_bwd3c.txt
