cd /root/repo
python tools/gpu_quick.py 2>&1 | tail -2
run() { echo "$@"; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  value %.1f MT/s  ms %.2f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['roofline']['stage_ms'])"; }
run ASTCENC_B200_WARPS_REFINE=32
run ASTCENC_B200_WARPS_REFINE=28
run ASTCENC_B200_WARPS_REFINE=24
