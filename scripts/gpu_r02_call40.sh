#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check_nvls8.log 2>&1; echo "dist rc=$?"; grep -i "dist overlap\|dist nvls\|error" gpurun_out/r02_dist_check_nvls8.log | tail -6 | cut -c1-220
B="--steps 20 --warmup 5 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],4), 'host', round(d.get('host_issue_ms_per_step',0),4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d.get('allreduce'))
else:
    print(sys.argv[1], 'NO LINE'); print(open(sys.argv[1]).read()[-1200:])
PY
}
runN() { n=$1; name=$2; shift; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n $B "$@" > gpurun_out/r02_w_$name.log 2>&1; show gpurun_out/r02_w_$name.log; }
runN 8 n8_auto
runN 4 n4_auto
runN 4 n4_nccl --allreduce nccl
