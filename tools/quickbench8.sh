#!/bin/bash
# tuning run: refine kernel with 26 / 28 warps x 72 registers (variant libraries) vs 24 x 80, and set-up staging off
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})"; }
cd /root/repo
run A=1
run ASTCENC_B200_STAGE_SETUP=0
cp astc-encoder_b200/libastcenc_b200.so /tmp/keep.so
for n in 832 896; do
  cp astc-encoder_b200/libastcenc_b200_r$n.so astc-encoder_b200/libastcenc_b200.so
  python tools/gpu_quick.py 2>&1 | tail -1
  run VARIANT=$n
done
cp /tmp/keep.so astc-encoder_b200/libastcenc_b200.so
