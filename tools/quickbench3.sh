cd /root/repo
run() { echo "$@"; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  value %.1f MT/s  ms %.2f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; }
run ASTCENC_B200_SYNC_MASK=0 ASTCENC_B200_WARPS_REFINE=32
run ASTCENC_B200_SYNC_MASK=0 ASTCENC_B200_WARPS_REFINE=28
run ASTCENC_B200_SYNC_MASK=0 ASTCENC_B200_WARPS_REFINE=24
