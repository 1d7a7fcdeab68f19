#!/usr/bin/env python3
"""Aggregate an ncu SASS page by source line using nvdisasm -g line info (dev tool).
   python tools/ncu_by_line.py report.ncu-rep lib.so kernel-substring [blocks] [top] [file-filter] [nth-matching-launch]"""
import csv, subprocess, sys, re, os, tempfile, glob
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
blocks = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
ffilter = sys.argv[6] if len(sys.argv) > 6 else ""
nth = int(sys.argv[7]) if len(sys.argv) > 7 else 0          # which matching launch of the report
by = 2 if os.environ.get("BY_SAMPLES") else 0              # BY_SAMPLES=1: sort by stall samples instead of instructions
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max(glob.glob(os.path.join(tmp, "*.cubin")), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# address -> (file, line) for the selected kernel's section
amap = {}
insec = False
cur = ("?", 0)
for l in dis.splitlines():
    if l.startswith("//--------------------- .text."):
        insec = kern in l
        continue
    if not insec:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,8})\*/", l)
    if m:
        amap[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = [i for i in starts if kern in rows[i][1]][nth]
end = next((i for i in starts if i > sel), len(rows))
rows = rows[sel:end]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
base = int(data[0][0], 16)
agg = {}
ti = ts = 0
for r in data:
    off = int(r[0], 16) - base
    key = amap.get(off, ("?", 0))
    a = agg.setdefault(key, [0, 0, 0])
    ie = int(r[ci["Instructions Executed"]]); te = int(r[ci["Thread Instructions Executed"]]); sm = int(r[ci["# Samples"]])
    a[0] += ie; a[1] += te; a[2] += sm
    ti += ie; ts += sm
srcs = {}
def srcline(f, n):
    if f not in srcs:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f))
        srcs[f] = open(cand[0]).read().splitlines() if cand else []
    s = srcs[f]
    return s[n - 1].strip()[:90] if 0 < n <= len(s) else ""
print(f"{'file:line':34s} {'instr/blk':>9s} {'%instr':>6s} {'lanes':>5s} {'%smp':>6s}  source")
for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][by]):
    if ffilter and ffilter not in f:
        continue
    if top <= 0:
        break
    top -= 1
    print(f"{f+':'+str(n):34s} {a[0]/blocks:9.1f} {100*a[0]/ti:6.2f} {a[1]/max(a[0],1):5.1f} {100*a[2]/max(ts,1):6.2f}  {srcline(f, n)}")
