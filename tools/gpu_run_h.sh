#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2h}
( timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_$T.log 2>&1; tail -4 gpurun_out/pytest_$T.log
run() { # name, env...
  n=$1; shift
  for c in 1 2; do
    env "$@" timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${T}_${n}_c$c.json 2> gpurun_out/bench_${T}_${n}_c$c.err
  done
}
run base X=1
run nostage ASTCENC_B200_STAGE_SETUP=0
run w16 ASTCENC_B200_WARPS_SETUP_1P=16
for c in 0 3; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${T}_base_c$c.json 2> gpurun_out/bench_${T}_base_c$c.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${T}_*_c*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],2), 'val', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
