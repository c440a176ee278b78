#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_staged_backward.py -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu32.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu32.log | cut -c1-300
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
CKB_PDL=0 timeout 300 python bench.py $B --profile-out gpurun_out/r02_prof_cp32_nopdl.json > gpurun_out/r02_i_nopdl.log 2>&1; show gpurun_out/r02_i_nopdl.log
timeout 300 python bench.py $B --profile-out gpurun_out/r02_prof_cp32.json > gpurun_out/r02_i_pdl.log 2>&1; show gpurun_out/r02_i_pdl.log
timeout 300 python bench.py $B --workload qt28_cp_k32 --batch 512 > gpurun_out/r02_i_k32.log 2>&1; show gpurun_out/r02_i_k32.log
python - <<'PY'
import json
for n in ('gpurun_out/r02_prof_cp32_nopdl.json','gpurun_out/r02_prof_cp32.json'):
    d=json.load(open(n)); print(n, 'fwd', round(sum(r['fwd_ms'] for r in d),4), 'bwd', round(sum(r['bwd_ms'] for r in d),4))
    print(' '.join(f"{r['fwd_ms']*1e3:.0f}/{r['bwd_ms']*1e3:.0f}" for r in d))
PY
