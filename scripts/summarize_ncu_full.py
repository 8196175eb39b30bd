#!/usr/bin/env python
"""Summarise ``ncu --set full`` reports into a markdown table (one row per captured launch).

usage: summarize_ncu_full.py profiles/<name>.md "<title>" gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        yield d


def fmt(v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return f"{x / 1e6:.2f} MB"
    if u in ("ns", "us", "ms"):
        x *= {"ns": 1e-3, "us": 1, "ms": 1e3}[u]
        return f"{x:.1f} us"
    return f"{x:.1f}" if x != int(x) else str(int(x))


def main():
    dst, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`, one row per captured launch "
                "(each launch replayed ~40x, cold caches).\n\n")
        f.write("| kernel | " + " | ".join(n for _, n in METRICS) + " |\n|---|" + "---:|" * len(METRICS) + "\n")
        for rep in reps:
            for d in rows_of(rep):
                name = d["Kernel Name"][0].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
                cells = [fmt(*d[m]) if m in d else "-" for m, _ in METRICS]
                f.write(f"| `{name}` | " + " | ".join(cells) + " |\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
