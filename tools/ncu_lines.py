#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
stall samples (with the dominant stall reasons) and warp-instructions executed, first launch only.
usage: python tools/ncu_lines.py export.csv [top_n]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
rows = list(csv.reader(open(path)))
hdr = None
cur_file = None
lines = {}
launch = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Kernel Name":
        launch += 1
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or launch > 1 and False:
        continue
    if r[0] == "" or r[0] == "-":
        continue          # SASS row
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (cur_file, ln)
    e = lines.setdefault(key, {"src": r[1], "samples": 0, "inst": 0, "stalls": defaultdict(int)})
    def num(x):
        try:
            return int(x)
        except (TypeError, ValueError):
            return 0
    e["samples"] += num(d.get("# Samples"))
    e["inst"] += num(d.get("Instructions Executed"))
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try:
                e["stalls"][k[6:]] += int(v)
            except ValueError:
                pass
tot = sum(e["samples"] for e in lines.values())
toti = sum(e["inst"] for e in lines.values())
print(f"total samples {tot}, warp-instructions {toti} (all captured launches)")
for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(e["stalls"].items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*e['samples']/tot:5.1f}% {100*e['inst']/toti:5.1f}%i {f}:{ln:<5d} {' '.join(f'{k}={v}' for k, v in st):48s} | {e['src'].strip()[:110]}")
