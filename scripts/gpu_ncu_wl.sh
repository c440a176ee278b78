#!/bin/bash
# ncu --set full capture of one launch of a kernel under `bench.py --workload ...`
# usage: scripts/gpu_ncu_wl.sh <tag> <workload> <kernel-regex> [skip]
tag=$1; wl=$2; k=$3; skip=${4:-0}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${k} -s ${skip} -c 1 \
    -o gpurun_out/full_${tag}_${k} -f \
    python bench.py --workload ${wl} --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${tag}_${k}.log 2>&1
echo "full ${k} rc=$?"
