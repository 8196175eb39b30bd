#!/usr/bin/env python
"""profiles/<tag>_kernel_shares.json for bench.py: the share of the summed kernel time of one step per dominant kernel
(from the ncu launch list) and the DRAM traffic of one captured launch of each (from the `ncu --set full` reports).

usage: kernel_shares.py gpurun_out/<tag>_launches.csv profiles/<tag>_kernel_shares.json name=report.ncu-rep ...
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

NAMES = {"k_gat_fwd_tiled<1": "k_gat_fwd_tiled<AFFINE1>", "k_gat_bwd_dst_tiled<1": "k_gat_bwd_dst_tiled<AFFINE1>",
         "k_gat_bwd_src_tiled": "k_gat_bwd_src_tiled", "k_tc_proj3r": "k_tc_proj3r", "k_tc_dw3": "k_tc_dw3"}


def main():
    src, dst = sys.argv[1:3]
    rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
    agg, total = collections.defaultdict(float), 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}[r["Metric Unit"]]
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        total += v
        for pre, key in NAMES.items():
            if name.startswith(pre):
                agg[key] += v
    out = {"share": {k: round(v / total, 4) for k, v in agg.items()}, "traffic": {}, "source_launches": src}
    for spec in sys.argv[3:]:
        key, rep = spec.split("=", 1)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rs = list(csv.reader(io.StringIO(txt)))
        if len(rs) < 3:
            continue
        hdr, units = rs[0], rs[1]
        best = None
        for r in rs[2:]:       # the largest captured launch (bond graph)
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))

            def val(m):
                x = float(d[m].replace(",", ""))
                return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[m], 1)
            t = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            if best is None or t > best:
                best = t
        out["traffic"][key] = int(best)
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
