#!/usr/bin/env python
"""Joins an ncu SASS source page with nvdisasm line info: executed warp instructions and stall
samples per CUDA source line (inlined callee lines are attributed to the innermost file:line).
Usage: python tools/ncu_lines.py report.ncu-rep kernel_substring [top_n]   (runs without a GPU)"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "multi-uav-pursuit-evasion_b200", "libhs_b200.so")


def line_table(kernel_sub):
    tmp = "/tmp/_ncu_lines"
    os.makedirs(tmp, exist_ok=True)
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    table, chain, on = {}, [], False
    last_chain = [("?", 0)]
    for l in sass.splitlines():
        if l.startswith(".text."):
            on = kernel_sub in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m:
            if chain:
                last_chain, chain = chain, []
            # innermost frame first, kernel-body frame last
            table[int(m.group(1), 16)] = (last_chain[0], last_chain[-1], m.group(2))
    return table


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tab = line_table(ksub)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = [r for r in rows[h + 1:] if len(r) > isamp and r[ia].startswith("0x")]
    base = int(body[0][ia], 16)
    inner, outer = collections.Counter(), collections.Counter()
    s_inner, s_outer = collections.Counter(), collections.Counter()
    tot_e = tot_s = 0
    for r in body:
        off = int(r[ia], 16) - base
        cur, stack, _ = tab.get(off, (("?", 0), ("?", 0), ""))
        e, s = int(r[ie] or 0), int(r[isamp] or 0)
        inner[cur] += e; outer[stack] += e
        s_inner[cur] += s; s_outer[stack] += s
        tot_e += e; tot_s += s
    print(f"# {rep}: {tot_e} warp instructions, {tot_s} stall samples, {len(body)} SASS instructions")
    print("\n## by outermost (kernel body) line: executed %, samples %")
    for k, v in sorted(s_outer.items(), key=lambda x: -x[1])[:top]:
        print(f"{k[0]}:{k[1]:<5d} exec {100*outer[k]/tot_e:5.1f}%  samples {100*v/tot_s:5.1f}%")
    print("\n## by innermost line")
    for k, v in sorted(s_inner.items(), key=lambda x: -x[1])[:top]:
        print(f"{k[0]}:{k[1]:<5d} exec {100*inner[k]/tot_e:5.1f}%  samples {100*v/tot_s:5.1f}%")


if __name__ == "__main__":
    main()
