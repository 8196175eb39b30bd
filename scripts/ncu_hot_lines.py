#!/usr/bin/env python
"""Warp-stall samples of an ``ncu --set full --import-source on`` report aggregated per CUDA source line.
usage: ncu_hot_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                samples = int(d.get("Warp Stall Sampling (All Samples)", "0") or 0)
                inst = int(d.get("Instructions Executed", "0") or 0)
            except ValueError:
                continue
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
            lines.append((samples, inst, cur_file, int(r[0]), r[1].strip()[:110], stalls))
    total = sum(l[0] for l in lines) or 1
    tot_inst = sum(l[1] for l in lines) or 1
    print(f"total samples {total}, instructions {tot_inst}")
    for s, inst, f, ln, src, st in sorted(lines, key=lambda l: -l[0])[:top]:
        top_st = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100 * s / total:5.1f}% smp {100 * inst / tot_inst:5.1f}% ins  {f}:{ln:<4d} {src}\n        [{top_st}]")


if __name__ == "__main__":
    main()
