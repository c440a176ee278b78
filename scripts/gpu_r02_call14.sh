#!/bin/bash
mkdir -p gpurun_out
{
for fl in 0 8192 16384 24576; do echo "=== bwd3 flags $fl"; timeout 200 python scripts/stress_bwd3b.py 2048 100 $fl 2>&1 | grep -v "^frame" | grep "MISMATCH\|done\|Error" | tail -6; done
echo "=== bwd2 vs bwd1"; timeout 200 python scripts/stress_bwd3b.py 2048 100 2048 3587 2>&1 | grep -v "^frame" | grep "MISMATCH\|done\|Error" | tail -6
} > gpurun_out/r02_stress_bwd3d.txt 2>&1; cut -c1-200 gpurun_out/r02_stress_bwd3d.txt
