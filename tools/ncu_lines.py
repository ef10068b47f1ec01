#!/usr/bin/env python
"""Per-source-line warp-stall samples of one profiled launch, from `ncu -i X.ncu-rep --page source --csv --print-source
cuda,sass > src.csv`.   python tools/ncu_lines.py src.csv [launch index] [file substring] [top N]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fsub = sys.argv[3] if len(sys.argv) > 3 else "tile.cu"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(path)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = dict(file=r[1], rows=[])
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
files = []
for s in sections:
    if s["file"] not in files:
        files.append(s["file"])
per_launch = len(files)
secs = sections[launch * per_launch:(launch + 1) * per_launch]
total = 0
for s in secs:
    hdr = s["rows"][1]
    i_samp = hdr.index("# Samples")
    for r in s["rows"][2:]:
        if len(r) > i_samp and r[i_samp].isdigit():
            total += int(r[i_samp])
print("total samples (all files)", total)
for s in secs:
    if fsub not in s["file"]:
        continue
    hdr = s["rows"][1]
    i_line, i_src, i_samp, i_inst = 0, 1, hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    by_line = defaultdict(lambda: [0, 0, "", defaultdict(int)])
    line, src = None, ""
    for r in s["rows"][2:]:
        if len(r) <= i_samp:
            continue
        if r[i_line].strip():
            line, src = int(r[i_line]), r[i_src]
        if line is None:
            continue
        e = by_line[line]
        e[2] = src
        if r[i_samp].isdigit():
            e[0] += int(r[i_samp])
        if r[i_inst].isdigit():
            e[1] += int(r[i_inst])
        for i, h in stall_cols:
            if i < len(r) and r[i].isdigit():
                e[3][h] += int(r[i])
    tot_f = sum(e[0] for e in by_line.values())
    print(s["file"], "samples", tot_f)
    for line, e in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(e[3].items(), key=lambda kv: -kv[1])[:3]
        print(f"{line:5d} {100.0 * e[0] / max(total, 1):5.1f}% inst {e[1]:>10d}  {' '.join(f'{k[6:]}={v}' for k, v in st):40s} | {e[2].strip()[:90]}")
