#!/bin/bash
# Dev tool, run on the GPU box: tools/profile_round.sh <tag>
# 1. bench line (own arm, with the CPU baseline) -> gpurun_out/bench_<tag>.json
# 2. reference arm                                -> gpurun_out/bench_ref_<tag>.json
# 3. ncu launch list of the same bench command (durations + DRAM bytes per launch) -> gpurun_out/launches_<tag>.csv
tag=$1
cd /root/repo
python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_$tag.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
    --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
head -c 600 gpurun_out/bench_$tag.json; echo; head -c 400 gpurun_out/bench_ref_$tag.json; echo
