#!/bin/bash
# 2 GPUs: TABLE_INPUT fusion parity + bench; op-only chunked overlap at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_staged_backward.py tests/test_gpu_edge_cases.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu28.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_gpu28.log | cut -c1-300
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
timeout 300 python bench.py $B --profile-out gpurun_out/r02_prof_cp28.json > gpurun_out/r02_e_n1.log 2>&1; show gpurun_out/r02_e_n1.log
timeout 300 python bench.py $B --workload qt28_cp_k32 --batch 512 > gpurun_out/r02_e_k32.log 2>&1; show gpurun_out/r02_e_k32.log
run2() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $B "$@" > gpurun_out/r02_e_$name.log 2>&1; show gpurun_out/r02_e_$name.log; }
run2 n2_serial --grad-chunks 0
run2 n2_noar --no-grad-allreduce
run2 n2_c1 --grad-chunks 1
run2 n2_c2 --grad-chunks 2
run2 n2_c4 --grad-chunks 4
run2 n2_c8 --grad-chunks 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r02_dist_check.log | cut -c1-200
