#!/bin/bash
# dev tool (GPU box): the round's closing run of the shipped build -> gpurun_out/
#   GPU tests, bench lines of every config (+ the reference arm), ncu launch list + full sets, racecheck on tiny images, step statistics
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r02f}
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_$T.log 2>&1
tail -4 gpurun_out/pytest_$T.log
bash tools/profile_round.sh $T
bash tools/profile_r2.sh $T 2>&1 | tail -6
cat > /tmp/rc.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
import astc_images as I
import __graft_entry__ as g
pkg = g.load_package()
for (prof, b, q, gen) in [(1, 6, 60.0, 'photo_like'), (1, 4, 10.0, 'photo_like'), (3, 6, 60.0, 'hdr_noise'), (1, 8, 98.0, 'voronoi_flat'), (1, 6, 60.0, 'uniform_noise')]:
    img = getattr(I, gen)(48, 48, seed=3)
    ctx = pkg.Context(pkg.config_init(prof, b, b, q, 32))
    out = ctx.compress_image(img)
    ctx.close()
    print(prof, b, q, gen, int(out.sum()))
vol = np.ascontiguousarray(I.photo_like(12 * 6, 14, seed=5).reshape(6, 12, 14, 4))
ctx = pkg.Context(pkg.config_init(1, 4, 4, 60.0, 32, block_z=4))
print('4x4x4', int(ctx.compress_image(vol).sum()))
ctx.close()
PY
( time timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python /tmp/rc.py ) > gpurun_out/racecheck_$T.txt 2>&1
grep -E "RACECHECK SUMMARY|hazard|ERROR SUMMARY" gpurun_out/racecheck_$T.txt | sort | uniq -c | sort -rn | head -8
if [ -f astc-encoder_b200/libastcenc_b200_stats.so ]; then
  ASTCENC_B200_LIB=$PWD/astc-encoder_b200/libastcenc_b200_stats.so timeout 300 python tools/step_stats.py > gpurun_out/step_stats_$T.txt 2>&1
fi
