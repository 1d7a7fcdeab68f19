#!/bin/bash
# dev tool (GPU box): GPU test suite + default bench line of the current build -> gpurun_out/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2b}
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_$T.log 2>&1
tail -6 gpurun_out/pytest_$T.log
timeout 600 python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err
tail -c 1800 gpurun_out/bench_${T}.json; echo; tail -3 gpurun_out/bench_${T}.err
