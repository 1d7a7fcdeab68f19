#!/usr/bin/env python3
"""Aggregate an ncu source page by device function (dev tool).
   python tools/ncu_by_function.py report.ncu-rep lib.so [blocks] [kernel-substring] [nth-matching-launch]"""
import csv, subprocess, sys, re
rep, lib = sys.argv[1], sys.argv[2]
blocks = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
kern = sys.argv[4] if len(sys.argv) > 4 else ""
nth = int(sys.argv[5]) if len(sys.argv) > 5 else 0      # which matching launch of the report
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
syms = []
insym = False
for l in elf.splitlines():
    if l.startswith(".section .symtab"):
        insym = True; continue
    if insym and l.startswith(".section"):
        insym = False; continue
    if insym:
        p = l.split()
        if len(p) >= 7 and p[3] in ("0x2", "0x12", "0x22"):
            try:
                off = int(p[1], 16); size = int(p[2], 16)
            except ValueError:
                continue
            name = p[-1]
            m2 = re.search(r"_cu_[0-9a-f]{8}(\d+)(.*)$", name)
            short = name
            if m2:
                short = m2.group(2)[:int(m2.group(1))]
            if size and name.startswith("$") and kern in name.split("$")[1]:
                syms.append((off, size, short))
syms.sort()
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernels may be in the report: take the first whose name matches
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
def fn_of(off):
    for o, s, n in syms:
        if o <= off < o + s:
            return n
    return "<kernel body>"
tot_i = tot_s = 0
for r in data:
    if len(r) < len(hdr):
        continue
    off = int(r[0], 16) - base
    f = fn_of(off)
    a = agg.setdefault(f, [0, 0, 0, 0])
    ie = int(r[ci["Instructions Executed"]]); te = int(r[ci["Thread Instructions Executed"]]); sm = int(r[ci["# Samples"]])
    a[0] += ie; a[1] += te; a[2] += sm; a[3] += 1
    tot_i += ie; tot_s += sm
print(f"{'function':44s} {'instr/blk':>10s} {'%instr':>7s} {'lanes':>6s} {'%samples':>8s} {'static':>7s}")
for f, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    if a[0] == 0:
        continue
    print(f"{f:44s} {a[0]/blocks:10.0f} {100*a[0]/tot_i:7.2f} {a[1]/max(a[0],1):6.1f} {100*a[2]/max(tot_s,1):8.2f} {a[3]:7d}")
print("total instr/blk", tot_i / blocks)
