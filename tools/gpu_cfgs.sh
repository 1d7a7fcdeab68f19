#!/bin/bash
# dev tool (GPU box): bench lines of configs 0 2 3 for library builds:  bash tools/gpu_cfgs.sh <tag> <name|cur> ...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=$1; shift
for n in "$@"; do
  for c in 0 2 3; do
    if [ "$n" = cur ]; then L=""; else L="$PWD/astc-encoder_b200/libastcenc_b200_$n.so"; fi
    ASTCENC_B200_LIB=$L timeout 600 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline 2> gpurun_out/cfg_${T}_${n}_$c.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$n', 'config', $c, 'ms %.3f' % d['ms_per_step'], 'MT/s %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'])"
  done
done | tee gpurun_out/cfgs_$T.txt
