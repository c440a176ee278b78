#!/bin/bash
# 8 GPUs: scaling check of the bench (overlapped vs serial all-reduce), multi-GPU parity
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],4), 'host', round(d.get('host_issue_ms_per_step',0),4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
else:
    print(sys.argv[1], 'NO LINE'); import subprocess; print(open(sys.argv[1]).read()[-1500:])
PY
}
runN() { n=$1; name=$2; shift; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n $B "$@" > gpurun_out/r02_s_$name.log 2>&1; show gpurun_out/r02_s_$name.log; }
runN 8 n8_default
runN 8 n8_serial --grad-chunks 0
runN 8 n8_noar --no-grad-allreduce
runN 8 n8_c4 --grad-chunks 4
runN 4 n4_default
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check8.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r02_dist_check8.log | cut -c1-200
