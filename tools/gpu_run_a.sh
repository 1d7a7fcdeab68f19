#!/bin/bash
# dev tool (GPU box): round-2 first measurement - GPU tests, bench lines of configs 0-3, unroll A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_$T.txt 2>&1
python -c "import os;print('affinity',len(os.sched_getaffinity(0)),'cpus',os.cpu_count());print(open('/sys/fs/cgroup/cpu.max').read())" >> gpurun_out/smi_$T.txt 2>&1
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_$T.log 2>&1
tail -5 gpurun_out/pytest_$T.log
for c in 1 0 2 3; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_${T}_c$c.json 2> gpurun_out/bench_${T}_c$c.err
  tail -c 600 gpurun_out/bench_${T}_c$c.json; echo
done
ASTCENC_B200_LIB=$PWD/astc-encoder_b200/libastcenc_b200_unroll.so timeout 300 python bench.py --config 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${T}_unroll.json 2> gpurun_out/bench_${T}_unroll.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2a_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],2), 'val', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['stage_ms'], d.get('cpu_baseline',{}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
