#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2d}
timeout 300 python tools/diag_hdr3.py 96 > gpurun_out/diag_hdr3_$T.txt 2>&1; tail -5 gpurun_out/diag_hdr3_$T.txt
( timeout 600 compute-sanitizer --tool initcheck --print-limit 8 python tools/diag_hdr3.py 48 ) > gpurun_out/initcheck_$T.txt 2>&1; grep -E "Uninitialized|ERROR SUMMARY|at 0x|in .*\(" gpurun_out/initcheck_$T.txt | head -24
( timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python tools/diag_hdr3.py 48 ) > gpurun_out/memcheck_$T.txt 2>&1; grep -E "Invalid|ERROR SUMMARY|at 0x" gpurun_out/memcheck_$T.txt | head -12
( timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_$T.log 2>&1; tail -3 gpurun_out/pytest_$T.log
for v in "" nohot; do
  L=$PWD/astc-encoder_b200/libastcenc_b200${v:+_$v}.so
  for c in 1 2; do
    ASTCENC_B200_LIB=$L timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${T}_${v:-hot}_c$c.json 2> gpurun_out/bench_${T}_${v:-hot}_c$c.err
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${T}_*_c*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],2), 'val', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,2) for k,v in d['roofline']['stage_ms'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
