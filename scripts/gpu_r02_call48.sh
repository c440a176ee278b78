#!/bin/bash
# 8 GPUs: configs[3] (PoonDomingos 3x32x32, K=128, global batch 4096) and the north-star line once more
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],4), 'host', round(d.get('host_issue_ms_per_step',0),4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d.get('allreduce'))
else:
    print(sys.argv[1], 'NO LINE'); print(open(sys.argv[1]).read()[-1200:])
PY
}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --workload pd32_cp_k128 > gpurun_out/r02_pd32_n8.log 2>&1; show gpurun_out/r02_pd32_n8.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_cp_n8_final.log 2>&1; show gpurun_out/r02_cp_n8_final.log
