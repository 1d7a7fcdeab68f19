#!/usr/bin/env python3
"""Dev tool: executed local-memory instructions (LDL / STL) per device function from an ncu report.
   python tools/ncu_local_mem.py report.ncu-rep lib.so kernel-substring [nth-matching-launch]"""
import csv, subprocess, sys, re
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
nth = int(sys.argv[4]) if len(sys.argv) > 4 else 0
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
            short = m2.group(2)[:int(m2.group(1))] if m2 else name
            if size and name.startswith("$") and kern in name.split("$")[1]:
                syms.append((off, size, short))
syms.sort()
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = [i for i in starts if kern in rows[i][1]][nth]
end = next((i for i in starts if i > sel), len(rows))
rows = rows[sel:end]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
base = int(data[0][0], 16)
def fn_of(off):
    for o, s, n in syms:
        if o <= off < o + s:
            return n
    return "<kernel body>"
agg = {}
tot = 0
for r in data:
    if len(r) < len(hdr):
        continue
    op = r[1].strip()
    ie = int(r[ci["Instructions Executed"]])
    tot += ie
    if op.startswith("LDL") or op.startswith("STL") or " LDL" in op[:12] or " STL" in op[:12]:
        a = agg.setdefault(fn_of(int(r[0], 16) - base), [0, 0, 0])
        a[0 if "LDL" in op[:14] else 1] += ie
        a[2] += int(r[ci["# Samples"]])
print(f"{'function':44s} {'LDL exec':>12s} {'STL exec':>12s} {'samples':>8s}")
for f, a in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][1]))[:25]:
    print(f"{f:44s} {a[0]:12d} {a[1]:12d} {a[2]:8d}")
print("total executed instructions", tot, " local", sum(a[0] + a[1] for a in agg.values()))
