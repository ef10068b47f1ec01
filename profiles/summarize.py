#!/usr/bin/env python
"""Turn an ncu report (`ncu --set full ... -o gpurun_out/NAME`) into the markdown summary and traffic.json kept here.

    python profiles/summarize.py gpurun_out/r01_qft30_full.ncu-rep profiles/r01_qft30_ncu_full.md 30
"""
import csv
import json
import os
import subprocess
import sys

rep, out_md, n = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "lts__t_sector_hit_rate.pct"]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
lines = [f"# ncu --set full: {os.path.basename(rep)}", "",
         "`ncu --set full --clock-control none --import-source on -k regex:\"k_tile|k_involution\" -s 12 -c 4 "
         "python bench.py --steps 1 --warmup 3 --no-cpu` under gpurun on one B200 (the passes of one timed step of the "
         f"{n}-qubit QFT).  Per-launch times under ncu are cold-cache and serialised; the un-profiled CUDA-event "
         "times are in the bench JSON next to this file.", ""]
traffic = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "")
    lines += [f"## {name}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}", "", "| metric | value | unit |",
              "|---|---|---|"]
    for w in want:
        if w in idx:
            lines.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[idx[h]] or 0))
          for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    st.sort(key=lambda x: -x[1])
    lines += ["| warp stall cycles per issued instruction (top) | " + ", ".join(f"{k} {v:.2f}" for k, v in st[:8]) + " | |", ""]
    t = sum(float(r[idx[m]].replace(",", "")) * scale.get(units[idx[m]], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    traffic.setdefault("k_tile" if "k_tile" in name else "k_involution", []).append(t)
open(out_md, "w").write("\n".join(lines))
tpath = os.path.join(os.path.dirname(out_md), "traffic.json")
old = json.load(open(tpath)) if os.path.exists(tpath) else {}
for k, v in traffic.items():
    old.setdefault(k, {})[str(n)] = sum(v) / len(v)
json.dump(old, open(tpath, "w"), indent=1)
print(json.dumps(old))
