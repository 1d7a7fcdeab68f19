#!/bin/bash
# dev tool (GPU box): GPU tests, bench of config 1, determinism diagnosis, racecheck on tiny images
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2b}
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_$T.log 2>&1
tail -4 gpurun_out/pytest_$T.log
timeout 600 python bench.py --config 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${T}_c1.json 2> gpurun_out/bench_${T}_c1.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${T}_c1.json').read().strip().splitlines()[-1])
    print('c1 ms', round(d['ms_per_step'],2), 'val', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['stage_ms'])
except Exception as e:
    print('bench ERR', e); print(open('gpurun_out/bench_${T}_c1.err').read()[-1500:])
PY
timeout 600 python tools/diag_hdr.py 1024 3 > gpurun_out/diag_hdr_$T.txt 2>&1; cat gpurun_out/diag_hdr_$T.txt | tail -16
cat > /tmp/rc.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
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
( time timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python /tmp/rc.py ) > gpurun_out/racecheck_$T.txt 2>&1
grep -E "RACECHECK SUMMARY|hazard|ERROR SUMMARY" gpurun_out/racecheck_$T.txt | sort | uniq -c | sort -rn | head -12
