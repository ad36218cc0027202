#!/usr/bin/env python
"""Per-source-line stall samples / executed instructions from `ncu --page source --csv --print-source cuda,sass`
(works for any kernel in the report; needs -lineinfo + --import-source on).  Usage: python tools/ncu_srclines.py rep [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fpath, hdr = "?", None
    agg = collections.defaultdict(lambda: [0, 0, ""])
    sass_top = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_e = hdr.index("# Samples") - len(hdr), hdr.index("Instructions Executed") - len(hdr)   # from the end: the source text may contain separators
            continue
        if hdr is None:
            continue
        def num(x):
            try:
                return int(x)
            except ValueError:
                return 0
        if len(r) < -i_s:
            continue
        if r[0] != "":                                    # a CUDA source line (aggregated over its SASS)
            key = (fpath, int(r[0]))
            agg[key][0] += num(r[i_s])
            agg[key][1] += num(r[i_e])
            agg[key][2] = r[1].strip()[:110]
        else:
            sass_top.append((num(r[i_s]), r[3].strip()[:70], fpath))
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_e = sum(v[1] for v in agg.values()) or 1
    print(f"# {rep}: {tot_s} stall samples, {tot_e} warp instructions")
    for (f, ln), (s, e, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / tot_s:5.1f}% samples {100 * e / tot_e:5.1f}% exec  {f}:{ln}  {src}")


if __name__ == "__main__":
    main()
