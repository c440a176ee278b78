#!/bin/bash
mkdir -p gpurun_out
for B in 1024 2048; do CUDA_LAUNCH_BLOCKING=1 timeout 200 python scripts/dbg_bwd3.py $B 2>&1 | grep -v "^frame" | tail -12; done > gpurun_out/r02_dbg_bwd3.txt 2>&1; cut -c1-1500 gpurun_out/r02_dbg_bwd3.txt
