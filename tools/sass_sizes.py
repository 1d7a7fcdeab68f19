#!/usr/bin/env python3
"""Per-function code sizes of the kernel image (dev tool): python tools/sass_sizes.py lib.so [top]"""
import subprocess, sys, re
lib = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rows = []
insym = False
for l in out.splitlines():
    if l.startswith(".section .symtab"):
        insym = True
        continue
    if insym and l.startswith(".section"):
        insym = False
        continue
    if insym:
        p = l.split()
        if len(p) >= 7 and p[3] in ("0x2", "0x12", "0x22"):
            try:
                size = int(p[2], 16)
            except ValueError:
                continue
            name = p[-1]
            short = name
            m2 = re.search(r"_cu_[0-9a-f]{8}(.*)$", name)
            if m2:
                rest = m2.group(1)
                m3 = re.match(r"(\d+)", rest)
                if m3:
                    n = int(m3.group(1))
                    # the length prefix may swallow a leading digit of nothing; names never start with a digit
                    body = rest[len(m3.group(1)):]
                    short = body[:n]
            rows.append((size, short))
rows.sort(reverse=True)
print("functions", len(rows), "total bytes", sum(r[0] for r in rows))
for s, n in rows[:top]:
    print(f"{s:8d}  {n}")
