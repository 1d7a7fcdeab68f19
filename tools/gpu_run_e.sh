#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2e}
timeout 600 python tools/diag_hdr3.py 96 102 128 256 > gpurun_out/diag_hdr3_$T.txt 2>&1; tail -18 gpurun_out/diag_hdr3_$T.txt
for v in "" g3 g5 g9 g13; do
  L=$PWD/astc-encoder_b200/libastcenc_b200${v:+_$v}.so
  for c in 1 2; do
    ASTCENC_B200_LIB=$L timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${T}_${v:-g1}_c$c.json 2> gpurun_out/bench_${T}_${v:-g1}_c$c.err
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
