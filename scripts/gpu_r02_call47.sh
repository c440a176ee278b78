#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r02_pytest_gpu47.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu47.log | cut -c1-300
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],4), 'host', round(d.get('host_issue_ms_per_step',0),4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
else:
    print(sys.argv[1], 'NO LINE'); print(open(sys.argv[1]).read()[-800:])
PY
}
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload pd32_cp_k128 --profile-out gpurun_out/r02_prof_pd32_47.json > gpurun_out/r02_z_pd32.log 2>&1; show gpurun_out/r02_z_pd32.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_prof_pd32_47.json'))
rows=sorted(((r.get('fwd_ms',0)+r.get('bwd_ms',0), r['step'], r.get('F'), round(r.get('fwd_ms',0),3), round(r.get('bwd_ms',0),3)) for r in d), reverse=True)
for x in rows[:8]: print(x)
print('total', sum(r.get('fwd_ms',0) for r in d), sum(r.get('bwd_ms',0) for r in d))
PY
