#!/bin/bash
# dev tool (GPU box): ncu evidence of the current build - launch list of one pass, full sets of the three big kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2p}
# launches per pass: 1 memset-free; kernels: wave 0: S R P; waves 1..9: S(cls 0) S(1-plane) R P; emit => 3 + 9*4 + 1 = 40
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:astc_ -s 40 -c 40 --csv --log-file gpurun_out/launches_$T.csv python tools/one_pass.py 2 > gpurun_out/ncu_launches_$T.log 2>&1
tail -2 gpurun_out/ncu_launches_$T.log
# full sets (2048^2 keeps the replays short): set-up wave 0, refine wave 1, set-up wave 2 class 0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:astc_wave_setup -s 19 -c 1 -o gpurun_out/prof_${T}_setup0 -f python tools/one_pass.py 2 2048 > gpurun_out/ncu_full_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:astc_wave_refine -s 11 -c 1 -o gpurun_out/prof_${T}_refine1 -f python tools/one_pass.py 2 2048 >> gpurun_out/ncu_full_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:astc_wave_setup -s 22 -c 1 -o gpurun_out/prof_${T}_setup2 -f python tools/one_pass.py 2 2048 >> gpurun_out/ncu_full_$T.log 2>&1
ls -la gpurun_out/prof_${T}_*.ncu-rep
cp astc-encoder_b200/libastcenc_b200.so gpurun_out/lib_$T.so
