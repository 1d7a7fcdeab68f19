#!/bin/bash
# dev tool (GPU box): A/B of two library builds on one box, step statistics with tail histograms, then the ncu evidence of profile_r2.sh
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2d}
{
  for rep in 1 2; do
    [ -f astc-encoder_b200/libastcenc_b200_prev.so ] && timeout 300 python tools/sweep_knobs.py --lib astc-encoder_b200/libastcenc_b200_prev.so - 2>&1 | sed 's/^/prev /'
    timeout 300 python tools/sweep_knobs.py - 2>&1 | sed 's/^/cur  /'
  done
} > gpurun_out/ab_$T.txt 2>&1
cat gpurun_out/ab_$T.txt
if [ -f astc-encoder_b200/libastcenc_b200_stats.so ]; then
  ASTCENC_B200_LIB=$PWD/astc-encoder_b200/libastcenc_b200_stats.so timeout 300 python tools/step_stats.py > gpurun_out/step_stats_$T.txt 2>&1
  tail -24 gpurun_out/step_stats_$T.txt
fi
bash tools/profile_r2.sh $T
