cd /root/repo
for cfg in "8 2" "8 1" "4 1" "4 2" "2 2" "4 4"; do set -- $cfg; echo "WARPS=$1 CTAS=$2"; ASTCENC_B200_WARPS=$1 ASTCENC_B200_CTAS_PER_SM=$2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  value %.1f MT/s  ms %.1f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; done
