#!/bin/bash
# dev tool (GPU box): the round's acceptance run of the current build -> gpurun_out/
#   GPU tests, default bench line, racecheck on tiny images of every profile, step statistics of the stats build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-chk}
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_$T.log 2>&1
tail -4 gpurun_out/pytest_$T.log
timeout 900 python bench.py > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err
tail -c 2500 gpurun_out/bench_${T}.json; echo; tail -3 gpurun_out/bench_${T}.err
cat > /tmp/rc.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import astc_images as I
import __graft_entry__ as g
pkg = g.load_package()
for (prof, b, q, gen) in [(1, 6, 60.0, 'photo_like'), (1, 4, 10.0, 'photo_like'), (3, 6, 60.0, 'hdr_noise'), (1, 8, 98.0, 'voronoi_flat')]:
    img = getattr(I, gen)(48, 48, seed=3)
    ctx = pkg.Context(pkg.config_init(prof, b, b, q, 32))
    out = ctx.compress_image(img)
    ctx.close()
    print(prof, b, q, gen, int(out.sum()))
PY
( time timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python /tmp/rc.py ) > gpurun_out/racecheck_$T.txt 2>&1
grep -E "RACECHECK SUMMARY|hazard|ERROR SUMMARY" gpurun_out/racecheck_$T.txt | sort | uniq -c | sort -rn | head -12
if [ -f astc-encoder_b200/libastcenc_b200_stats.so ]; then
  ASTCENC_B200_LIB=$PWD/astc-encoder_b200/libastcenc_b200_stats.so timeout 300 python tools/step_stats.py > gpurun_out/step_stats_$T.txt 2>&1
  tail -12 gpurun_out/step_stats_$T.txt
fi
