#!/usr/bin/env python3
"""Dev tool: condense an ncu launch-list CSV (tools/profile_round.sh) into
   profiles/<tag>_launches.csv  (the launches of ONE pipeline pass, one row per launch) and
   profiles/traffic.json        (DRAM bytes per pass, read by bench.py)
   python tools/launches_summary.py gpurun_out/launches_<tag>.csv <tag>"""
import csv, json, sys, os
src, tag = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(src)) if len(r) > 8]
while rows and rows[0][0] != "ID":
    rows.pop(0)
hdr = rows[0]; ci = {n: i for i, n in enumerate(hdr)}
L = {}; order = []
for r in rows[1:]:
    k = r[ci["ID"]]
    if k not in L:
        L[k] = {"name": r[ci["Kernel Name"]].split("(")[0]}; order.append(k)
    L[k][r[ci["Metric Name"]]] = float(r[ci["Metric Value"]].replace(",", ""))
ks = [k for k in order if "astc_wave" in L[k]["name"]]
passes = []; cur = []
for k in ks:
    cur.append(k)
    if "emit" in L[k]["name"]:
        passes.append(cur); cur = []
p = passes[-1]
out = os.path.join(root, "profiles", tag + "_launches.csv")
with open(out, "w") as f:
    f.write("launch,kernel,time_ms,dram_read_bytes,dram_write_bytes,warp_instructions,issue_active_pct,lanes_per_instruction\n")
    for i, k in enumerate(p):
        d = L[k]
        f.write("%d,%s,%.4f,%d,%d,%d,%.2f,%.2f\n" % (i, d["name"], d["gpu__time_duration.sum"] / 1e6, d["dram__bytes_read.sum"], d["dram__bytes_write.sum"],
                                                 d["smsp__inst_executed.sum"], d["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                                                 d.get("smsp__thread_inst_executed_per_inst_executed.ratio", d.get("smsp__thread_inst_executed.sum", 0.0) / max(d["smsp__inst_executed.sum"], 1.0))))
agg = {}
for k in p:
    d = L[k]; a = agg.setdefault(d["name"], dict(ms=0.0, rd=0.0, wr=0.0, inst=0.0, n=0))
    a["ms"] += d["gpu__time_duration.sum"] / 1e6; a["rd"] += d["dram__bytes_read.sum"]; a["wr"] += d["dram__bytes_write.sum"]; a["inst"] += d["smsp__inst_executed.sum"]; a["n"] += 1
tot = sum(a["ms"] for a in agg.values())
traffic = {"source": "ncu launch list of `python tools/one_pass.py 2` = two device-resident passes of config 1 (%s), last pipeline pass, %d launches" % (os.path.basename(src), len(p)),
           "dram_bytes_per_pass": int(sum(a["rd"] + a["wr"] for a in agg.values())),
           "pass_ms_under_ncu": tot,
           "per_kernel": {n: {"launches": a["n"], "ms": round(a["ms"], 3), "share": round(a["ms"] / tot, 4), "dram_read_bytes": int(a["rd"]), "dram_write_bytes": int(a["wr"]),
                              "warp_instructions": int(a["inst"])} for n, a in agg.items()}}
json.dump(traffic, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
