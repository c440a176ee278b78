#!/bin/bash
# ncu evidence for the bench command (1 GPU): launch list + one full capture of the top kernels.
# usage: scripts/gpu_profile.sh <tag> [kernel-regex ...]
tag=${1:-r01}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${tag}.log 2>&1
echo "launch list rc=$?"
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:${k} -s 2 -c 1 \
      -o gpurun_out/full_${tag}_${k} -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${tag}_${k}.log 2>&1
  echo "full ${k} rc=$?"
done
ls -la gpurun_out | tail -20
