#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu46.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu46.log | cut -c1-300
B="--steps 20 --warmup 5 --no-cpu-baseline"
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],4), 'host', round(d.get('host_issue_ms_per_step',0),4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
else:
    print(sys.argv[1], 'NO LINE'); print(open(sys.argv[1]).read()[-800:])
PY
}
timeout 120 python bench.py $B > gpurun_out/r02_y_n1.log 2>&1; show gpurun_out/r02_y_n1.log
CKB_GRAPHS=0 timeout 120 python bench.py $B > gpurun_out/r02_y_n1_nograph.log 2>&1; show gpurun_out/r02_y_n1_nograph.log
timeout 120 python bench.py $B --workload qt28_cp_k32 --batch 512 > gpurun_out/r02_y_k32.log 2>&1; show gpurun_out/r02_y_k32.log
CKB_GRAPHS=0 timeout 120 python bench.py $B --workload qt28_cp_k32 --batch 512 > gpurun_out/r02_y_k32_nograph.log 2>&1; show gpurun_out/r02_y_k32_nograph.log
