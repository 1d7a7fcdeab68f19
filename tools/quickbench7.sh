#!/bin/bash
# tuning run (historical: the ASTCENC_B200_DENSE_LIMIT / NO_STAGE knobs it used were removed after the measurement): staged tables on/off, wavefront limit of realign_weights (ms per pass, stage split)
run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})"; }
python tools/gpu_quick.py 2>&1 | tail -2
run A=1
run ASTCENC_B200_NO_STAGE=1
run ASTCENC_B200_STAGE_SETUP=1
run ASTCENC_B200_DENSE_LIMIT=9
run ASTCENC_B200_DENSE_LIMIT=12
run ASTCENC_B200_DENSE_LIMIT=16
ASTCENC_B200_DENSE_LIMIT=12 ASTCENC_B200_STAGE_SETUP=1 python tools/gpu_quick.py 2>&1 | tail -1
