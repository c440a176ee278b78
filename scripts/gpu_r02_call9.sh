#!/bin/bash
mkdir -p gpurun_out
{
echo "=== non-blocking dbg 2048"; timeout 200 python scripts/dbg_bwd3.py 2048 2>&1 | grep -v "^frame" | tail -8
echo "=== pytest full size"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size" 2>&1 | tail -5
echo "=== memcheck dbg 1024"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/dbg_bwd3.py 1024 2>&1 | grep -v "^frame" | tail -40
} > gpurun_out/r02_dbg_bwd3_b.txt 2>&1; cut -c1-300 gpurun_out/r02_dbg_bwd3_b.txt
