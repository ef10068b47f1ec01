#!/usr/bin/env python
"""Round-2 summaries of an `ncu --set full` report: per profiled launch the headline metrics, the stall reasons and the
executed-instruction mix by opcode (from the SASS source page).

    python profiles/summarize2.py gpurun_out/r02c_layers30.ncu-rep profiles/r02c_layers30_ncu_full.md "command line"
"""
import collections
import csv
import os
import subprocess
import sys

rep, out_md, cmd = sys.argv[1], sys.argv[2], sys.argv[3]


def page(name, *extra):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True,
                                          text=True).stdout.splitlines()))


rows = page("raw")
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum"]
lines = [f"# ncu --set full: {os.path.basename(rep)}", "", f"`{cmd}` under gpurun on one B200.  Per-launch times under ncu are "
         "cold-cache and serialised; the un-profiled CUDA-event times are in the bench JSON of the same name.", ""]
src = page("source", "--print-source", "sass")
starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"] + [len(src)]
for k, r in enumerate(rows[2:]):
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "")
    lines += [f"## launch {k}: {name}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}", "", "| metric | value | unit |",
              "|---|---|---|"]
    for w in want:
        if w in idx:
            lines.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
    st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[idx[h]] or 0))
          for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    st.sort(key=lambda x: -x[1])
    lines += ["| warp stall cycles per issued instruction (top) | " + ", ".join(f"{a} {b:.2f}" for a, b in st[:8]) + " | |"]
    if k + 1 < len(starts):
        sh = src[starts[k] + 1]
        six = {h: i for i, h in enumerate(sh)}
        mix = collections.Counter()
        for q in src[starts[k] + 2:starts[k + 1]]:
            if len(q) <= six["Instructions Executed"]:
                continue
            toks = q[six["Source"]].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            mix[op] += int(q[six["Instructions Executed"]] or 0)
        tot = sum(mix.values()) or 1
        lines += ["| executed warp instructions by opcode | " + ", ".join(f"{o} {100.0 * n / tot:.1f} %" for o, n in mix.most_common(12)) + " | |"]
    lines.append("")
open(out_md, "w").write("\n".join(lines))
print(out_md)
