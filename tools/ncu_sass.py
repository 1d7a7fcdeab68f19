#!/usr/bin/env python3
"""Dev tool: print the SASS of one device function from an ncu report with executed counts and stall samples.
   python tools/ncu_sass.py report.ncu-rep lib.so kernel-substring function-name [min_exec_fraction]"""
import csv, subprocess, sys, re
rep, lib, kern, fn = sys.argv[1:5]
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
lo = hi = None
insym = False
for l in elf.splitlines():
    if l.startswith(".section .symtab"):
        insym = True; continue
    if insym and l.startswith(".section"):
        insym = False; continue
    if insym:
        p = l.split()
        if len(p) >= 7 and p[3] in ("0x2", "0x12", "0x22") and p[-1].startswith("$") and kern in p[-1].split("$")[1]:
            m2 = re.search(r"_cu_[0-9a-f]{8}(\d+)(.*)$", p[-1])
            if m2 and m2.group(2)[:int(m2.group(1))] == fn:
                lo = int(p[1], 16); hi = lo + int(p[2], 16)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = next(i for i in starts if kern in rows[i][1])
end = next((i for i in starts if i > sel), len(rows))
rows = rows[sel:end]
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; ci = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[h + 1:] if r and r[0].startswith("0x")]
base = int(data[0][0], 16)
sel = [r for r in data if lo <= int(r[0], 16) - base < hi]
mx = max(int(r[ci["Instructions Executed"]]) for r in sel) or 1
for r in sel:
    ie = int(r[ci["Instructions Executed"]])
    if ie < thr * mx:
        continue
    print("%06x %10d %5.1f %6s  %s" % (int(r[0], 16) - base - lo, ie, int(r[ci["Thread Instructions Executed"]]) / max(ie, 1), r[ci["# Samples"]], r[1].strip()))
