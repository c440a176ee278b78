#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(sys.argv[1].split('/')[-1], 'ms', round(d['ms_per_step'],4), 'host', round(d.get('host_issue_ms_per_step',0),4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'whole', round(d['roofline'].get('whole_step',{}).get('frac',0),3))
else:
    print(sys.argv[1], 'NO LINE'); print(open(sys.argv[1]).read()[-600:])
PY
}
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload qt28_cp_k32 --batch 512 > gpurun_out/final2_bench_k32.log 2>&1; show gpurun_out/final2_bench_k32.log
timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload qt28_tucker_k64 > gpurun_out/final2_bench_tucker.log 2>&1; show gpurun_out/final2_bench_tucker.log
timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload rbt64_sos_k64 > gpurun_out/final2_bench_sos.log 2>&1; show gpurun_out/final2_bench_sos.log
