#!/bin/bash
# dev tool (GPU box): bench lines of every config (both arms for the headline) -> gpurun_out/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r02}
for c in 1 0 2 3; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${T}_c$c.json 2> gpurun_out/bench_${T}_c$c.err
done
timeout 900 python bench.py --impl reference --config 1 --steps 5 --warmup 1 > gpurun_out/bench_${T}_reference_c1.json 2>> gpurun_out/bench_${T}_c1.err
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${T}_c4.json 2> gpurun_out/bench_${T}_c4.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${T}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],2), 'val', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
    except Exception as e:
        print(f, 'ERR', e)
PY
