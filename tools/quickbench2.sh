cd /root/repo
run() { echo "$@"; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  value %.1f MT/s  ms %.2f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; }
run A=0
run ASTCENC_B200_SYNC_MASK=0x00
run ASTCENC_B200_SYNC_MASK=0x0F
run ASTCENC_B200_SYNC_MASK=0xF0
run ASTCENC_B200_SYNC_MASK=0x5F
run ASTCENC_B200_SYNC_MASK=0x1F
run ASTCENC_B200_WARPS_REFINE=16
run ASTCENC_B200_WARPS_REFINE=20
run ASTCENC_B200_WARPS_SETUP=12
run ASTCENC_B200_WARPS_SETUP=8
