#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_zz_uneven_units.py tests/test_gpu_tucker.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu30.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_gpu30.log | cut -c1-300
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
timeout 300 python bench.py $B --profile-out gpurun_out/r02_prof_cp30.json > gpurun_out/r02_g_n1.log 2>&1; show gpurun_out/r02_g_n1.log
timeout 300 python scripts/host_profile.py qt28_cp_k64 2048 > gpurun_out/r02_host_profile.txt 2>&1; head -60 gpurun_out/r02_host_profile.txt | cut -c1-160
