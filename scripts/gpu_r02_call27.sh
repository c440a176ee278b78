#!/bin/bash
# 2 GPUs: where does the overlapped all-reduce lose time?  (a) staging alone at N=1, (b) N=2 with 1/2/4 chunks, (c) NCCL CTA cap
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampling.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu27.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu27.log | cut -c1-300
B="--steps 20 --warmup 5 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],4), 'host', round(d.get('host_issue_ms_per_step',0),4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
else:
    print(sys.argv[1], 'NO LINE')
PY
}
timeout 300 python bench.py $B > gpurun_out/r02_d_n1.log 2>&1; show gpurun_out/r02_d_n1.log
timeout 300 python bench.py $B --stage-only --grad-chunks 4 > gpurun_out/r02_d_n1_stage4.log 2>&1; show gpurun_out/r02_d_n1_stage4.log
run2() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $B "$@" > gpurun_out/r02_d_$name.log 2>&1; show gpurun_out/r02_d_$name.log; }
run2 n2_serial --grad-chunks 0
run2 n2_noar --no-grad-allreduce
run2 n2_stage4_noar --no-grad-allreduce --stage-only --grad-chunks 4
run2 n2_c1 --grad-chunks 1
run2 n2_c2 --grad-chunks 2
run2 n2_c4 --grad-chunks 4
NCCL_MAX_CTAS=8 run2 n2_c4_cta8 --grad-chunks 4
NCCL_MAX_CTAS=4 run2 n2_c4_cta4 --grad-chunks 4
NCCL_MAX_CTAS=8 run2 n2_serial_cta8 --grad-chunks 0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r02_dist_check.log | cut -c1-200
