cd /root/repo
run() { echo "$@"; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  value %.1f MT/s  ms %.2f ' % (d['value'], d['ms_per_step']), d['roofline']['stage_ms'])"; }
run ASTCENC_B200_SYNC_MASK=0x000
run ASTCENC_B200_SYNC_MASK=0x100
run ASTCENC_B200_SYNC_MASK=0x200
run ASTCENC_B200_SYNC_MASK=0x300
