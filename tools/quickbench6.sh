cd /root/repo
run() { echo "$@"; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  value %.1f MT/s  ms %.2f ' % (d['value'], d['ms_per_step']), d['roofline']['stage_ms'])"; }
run A=0
run ASTCENC_B200_CTAS_PER_SM=2 ASTCENC_B200_WARPS_REFINE=12
run ASTCENC_B200_CTAS_PER_SM=3 ASTCENC_B200_WARPS_REFINE=8
run ASTCENC_B200_CTAS_PER_SM=2 ASTCENC_B200_WARPS_REFINE=12 ASTCENC_B200_WARPS_SETUP=8
