#!/bin/bash
mkdir -p gpurun_out
{
for fl in 0 0; do echo "=== bwd3 flags $fl"; timeout 200 python scripts/stress_bwd3b.py 2048 150 $fl 2>&1 | grep -v "^frame" | grep "MISMATCH\|done\|Error\|leaf" | tail -30; done
} > gpurun_out/r02_stress_bwd3f.txt 2>&1; cut -c1-250 gpurun_out/r02_stress_bwd3f.txt
