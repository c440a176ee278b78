#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu34.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu34.log | cut -c1-300
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
timeout 120 python bench.py $B --profile-out gpurun_out/r02_prof_cp34.json > gpurun_out/r02_k_n1.log 2>&1; show gpurun_out/r02_k_n1.log
timeout 120 python bench.py $B --workload qt28_cp_k32 --batch 512 > gpurun_out/r02_k_k32.log 2>&1; show gpurun_out/r02_k_k32.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload pd32_cp_k128 > gpurun_out/r02_k_pd32.log 2>&1; show gpurun_out/r02_k_pd32.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload qt28_tucker_k64 > gpurun_out/r02_k_tucker.log 2>&1; show gpurun_out/r02_k_tucker.log
timeout 300 python bench.py $B --workload rbt64_sos_k64 > gpurun_out/r02_k_sos.log 2>&1; show gpurun_out/r02_k_sos.log
