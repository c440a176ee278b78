#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_staged_backward.py -m gpu -q -x --timeout 60 > gpurun_out/r02_pytest_gpu33.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu33.log | cut -c1-300
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
timeout 120 python bench.py $B --workload qt28_cp_k32 --batch 512 > gpurun_out/r02_j_k32.log 2>&1; show gpurun_out/r02_j_k32.log
timeout 120 python bench.py $B --profile-out gpurun_out/r02_prof_cp33.json > gpurun_out/r02_j_persist.log 2>&1; show gpurun_out/r02_j_persist.log
python - <<'PY'
import json
for n in ('gpurun_out/r02_prof_cp33_legacy.json','gpurun_out/r02_prof_cp33.json'):
    try:
        d=json.load(open(n)); print(n, 'fwd', round(sum(r['fwd_ms'] for r in d),4), 'bwd', round(sum(r['bwd_ms'] for r in d),4))
        print(' '.join(f"{r['fwd_ms']*1e3:.0f}/{r['bwd_ms']*1e3:.0f}" for r in d))
    except Exception as e: print(n, e)
PY
