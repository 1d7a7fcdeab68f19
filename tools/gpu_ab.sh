#!/bin/bash
# dev tool (GPU box): A/B of library builds (make variant NAME=...) on one box: ms per device-resident pass of config 1, output equality
#   bash tools/gpu_ab.sh <tag> <name> <name> ...      ("cur" = the shipped build); then the stats build if present
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=$1; shift
{
  for n in "$@"; do
    if [ "$n" = cur ]; then
      timeout 300 python tools/sweep_knobs.py - 2>&1 | tail -1 | sed "s/^/cur      /"
    else
      timeout 300 python tools/sweep_knobs.py --lib astc-encoder_b200/libastcenc_b200_$n.so - 2>&1 | tail -1 | sed "s/^/$n      /"
    fi
  done
} > gpurun_out/ab_$T.txt 2>&1
cat gpurun_out/ab_$T.txt
if [ -f astc-encoder_b200/libastcenc_b200_stats.so ]; then
  ASTCENC_B200_LIB=$PWD/astc-encoder_b200/libastcenc_b200_stats.so timeout 300 python tools/step_stats.py > gpurun_out/step_stats_$T.txt 2>&1
  tail -14 gpurun_out/step_stats_$T.txt | cut -c1-330
fi
