#!/usr/bin/env python
"""Human-readable digest of a bench.py JSON line.  usage: show_bench.py file.json"""
import json
import sys


def main():
    d = json.load(open(sys.argv[1]))
    print("ms/step", d["ms_per_step"], "value", d["value"], "| tf32", d.get("value_tf32"), d.get("ms_per_step_tf32"),
          "| launches", d.get("gpu_launches"))
    if d.get("e2e"):
        print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "h2d", d["e2e"]["h2d_bytes_per_step"],
              "| arena", (d.get("e2e_arena") or {}).get("ms_per_step"), (d.get("e2e_arena") or {}).get("value"))
    print("clocks", d.get("clocks"))
    r = d.get("roofline")
    if r:
        print("step", r["step"])
        for k in r["kernels"]:
            print("  ", k["kernel"], k["us_per_launch"], "us", k["achieved"], "GB/s", k["frac"], "share",
                  k.get("share_of_kernel_time"), "traffic", k.get("traffic"))
        print("  ", {k: v for k, v in r.items() if k not in ("kernels", "step")})
    print("cpu", d.get("cpu_baseline"))
    for k, v in (d.get("extra") or {}).items():
        print(k, json.dumps(v)[:700])


if __name__ == "__main__":
    main()
