#!/bin/bash
mkdir -p gpurun_out
{
echo "=== racecheck"; timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python scripts/stress_bwd3b.py 1024 4 0 2>&1 | grep -v "^frame" | tail -60
echo "=== synccheck"; timeout 900 compute-sanitizer --tool synccheck --print-limit 30 python scripts/stress_bwd3b.py 1024 4 0 2>&1 | grep -v "^frame" | tail -40
} > gpurun_out/r02_sanitize_bwd3.txt 2>&1; cut -c1-400 gpurun_out/r02_sanitize_bwd3.txt | head -120
