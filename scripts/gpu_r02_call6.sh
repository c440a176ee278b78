#!/bin/bash
mkdir -p gpurun_out
for fl in 0 32 64 96 2048 2080; do echo "=== flags $fl"; timeout 300 python scripts/timeline_bwd2.py 2048 1 $fl 2>&1 | tail -16; done > gpurun_out/r02_timeline_bwd2_ab.txt 2>&1; cat gpurun_out/r02_timeline_bwd2_ab.txt
