#!/bin/bash
# usage: scripts/gpu_ncu_one.sh <tag> <kernel-regex> [skip]
tag=$1; k=$2; skip=${3:-0}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${k} -s ${skip} -c 1 \
    -o gpurun_out/full_${tag}_${k} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${tag}_${k}.log 2>&1
echo "full ${k} rc=$?"
