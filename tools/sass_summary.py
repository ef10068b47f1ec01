#!/usr/bin/env python
"""SASS mnemonic census of the kernels in qaqarot_b200/libb200sv.so (cuobjdump, no GPU needed), written as markdown.

    python tools/sass_summary.py [profiles/r02_k_tile_sass.md]

What to look for (B200_PROFILING.md): UTMALDG / UTMASTG = TMA tensor copies, SYNCS = mbarrier operations,
LDGSTS = the sm_80 cp.async path (must be 0 in k_tile), LDL / STL = register spills, DFMA / DMUL = the FP64 work.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qaqarot_b200", "libb200sv.so")
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "LDL", "STL", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG",
         "BAR", "BRA", "BSSY", "BSYNC", "IMAD", "LOP3", "SHF", "LDC", "LDCU", "R2UR"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else ""
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
        elif fn and "REG:" in line:
            usage[fn] = line.strip()
            fn = None
    kernels = collections.OrderedDict()
    detail = {}                       # per kernel: full mnemonics (BRA.U ...) and FP64 instructions with a UR operand
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            det = detail.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)\s*([^;]*);", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            if m.group(1) == "BRA" and re.search(r"\.U(\.|$)", m.group(2)):
                det["BRA.U"] += 1
            if m.group(1) in ("DFMA", "DMUL"):
                det["fp64"] += 1
                if re.search(r"\bUR\d+", m.group(3)):
                    det["fp64_ur"] += 1
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS census of libb200sv.so (sm_100a)", "",
             "`python tools/sass_summary.py` = `cuobjdump -sass` + `cuobjdump -res-usage` of the shipped library, static instruction",
             "counts per kernel.  `UTMALDG` / `UTMASTG` are the TMA tensor loads / stores, `SYNCS` the mbarrier operations; `LDGSTS`",
             "(the sm_80 `cp.async` of round 1) is gone from `k_tile`; `LDL` / `STL` are spills.", "",
             "| kernel | instructions | " + " | ".join(WATCH) + " | resources |", "|---|---|" + "---|" * (len(WATCH) + 1)]
    for (mangled, cnt), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("b200::", "")
        lines.append(f"| `{short}` | {sum(cnt.values())} | " + " | ".join(str(cnt.get(w, 0)) for w in WATCH) + f" | {usage.get(mangled, '')} |")
    lines += ["", "## k_tile: uniform datapath and TMA in SASS", "",
              "| instantiation | DFMA+DMUL | ... with a UR operand | LDCU | BRA.U | BRX | BSSY | UTMALDG | UTMASTG | SYNCS | LDGSTS |",
              "|---|---|---|---|---|---|---|---|---|---|---|"]
    for (mangled, cnt), name in zip(kernels.items(), demangle):
        if "k_tile<" not in name:
            continue
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("b200::", "")
        det = detail[mangled]
        lines.append(f"| `{short}` | {det['fp64']} | {det['fp64_ur']} | {cnt.get('LDCU', 0)} | {det['BRA.U']} | {cnt.get('BRX', 0)} | "
                     f"{cnt.get('BSSY', 0)} | {cnt.get('UTMALDG', 0)} | {cnt.get('UTMASTG', 0)} | {cnt.get('SYNCS', 0)} | "
                     f"{cnt.get('LDGSTS', 0)} |")
    lines += ["", "The remaining `LDL`/`STL` (column above) hold per-tile and per-round values (tile number, buffer index, cursor of the",
              "per-thread stream) outside the record loop; executed they are 1.8 % of a layer pass's warp instructions",
              "(`profiles/r02e_layers30_ncu_full.md`)."]
    text = "\n".join(lines) + "\n"
    if out:
        with open(out, "w") as f:
            f.write(text)
    print(text)


if __name__ == "__main__":
    main()
