#!/bin/bash
# dev tool (GPU box): HDR device-path diagnosis, step statistics, launch list of the current build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2c}
timeout 300 python tools/diag_hdr2.py > gpurun_out/diag_hdr2_$T.txt 2>&1; cat gpurun_out/diag_hdr2_$T.txt | tail -8
( timeout 900 python -m pytest tests -x -q -m gpu -k "hdr or golden" ) > gpurun_out/pytest_$T.log 2>&1; tail -3 gpurun_out/pytest_$T.log
timeout 600 python bench.py --config 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${T}_c1.json 2> gpurun_out/bench_${T}_c1.err
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${T}_c3.json 2> gpurun_out/bench_${T}_c3.err; tail -2 gpurun_out/bench_${T}_c3.err
python - <<PY
import json
for c in (1,3):
    try:
        d=json.loads(open('gpurun_out/bench_${T}_c%d.json'%c).read().strip().splitlines()[-1])
        print('c%d ms'%c, round(d['ms_per_step'],2), 'val', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['stage_ms'])
    except Exception as e:
        print('bench ERR', e)
PY
# step statistics (stats build): cycles per refinement-step part
cat > /tmp/st.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import astc_images as I
import __graft_entry__ as g
pkg = g.load_package()
img = I.photo_like(2048, 2048, seed=2024)
ctx = pkg.Context(pkg.config_init(1, 6, 6, 60.0, 32))
for _ in range(2):
    ctx.compress_image(img)
ctx.close()
PY
ASTCENC_B200_LIB=$PWD/astc-encoder_b200/libastcenc_b200_stats.so timeout 300 python /tmp/st.py > gpurun_out/step_stats_$T.txt 2>&1; tail -7 gpurun_out/step_stats_$T.txt
# launch list (one pass)
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 70 -c 36 --csv --log-file gpurun_out/launches_$T.csv python bench.py --config 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1
tail -3 gpurun_out/ncu_bench_$T.log | cut -c1-300
