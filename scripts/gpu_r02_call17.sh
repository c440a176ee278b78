#!/bin/bash
mkdir -p gpurun_out
{
for fl in 32768 65536; do echo "=== bwd3 flags $fl"; timeout 200 python scripts/stress_bwd3b.py 2048 200 $fl 2>&1 | grep -v "^frame" | grep "MISMATCH\|done\|Error" | tail -8; done
} > gpurun_out/r02_stress_bwd3g.txt 2>&1; cut -c1-250 gpurun_out/r02_stress_bwd3g.txt
