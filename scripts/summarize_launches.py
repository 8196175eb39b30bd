#!/usr/bin/env python
"""Summarise an ncu launch list (``--metrics gpu__time_duration.sum --csv``) into a per-kernel table.

usage: summarize_launches.py gpurun_out/<tag>_launches.csv profiles/<name>.md "<title>" [steps_in_capture]
"""
import collections
import csv
import re
import sys


def main():
    src, dst, title = sys.argv[1:4]
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}[r["Metric Unit"]]
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")[:90]
        agg[name][0] += 1
        agg[name][1] += v
        total += v
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n")
        f.write(f"Source: `{src}` ({len(rows)} launches over {steps} step(s); ncu serialises launches and runs them "
                "cold-cache, so compare SHARES, not absolute times).\n\n")
        f.write(f"Total kernel time {total:.0f} us = {total / steps:.0f} us/step.\n\n")
        f.write("| kernel | launches/step | avg us | us/step | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {c / steps:.1f} | {t / c:.1f} | {t / steps:.1f} | {100 * t / total:.1f}% |\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
