#!/usr/bin/env python
"""Does it help to run the step on a HIGH-PRIORITY stream?  The bond chain (the critical path) stays on the caller's
stream; the library's side streams have the default (lowest) priority, so a high-priority caller stream makes the
block scheduler place the critical chain's CTAs first whenever both have CTAs pending.  Times the bench step both ways."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timed(step, batches, n, stream):
    with torch.cuda.stream(stream):
        for i in range(10):
            step(batches[i % len(batches)])
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(n):
            step(batches[i % len(batches)])
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    step, batches = bench.make_step(batch=1024)
    torch.cuda.synchronize()
    lo, hi = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
    for rep in range(3):
        a = timed(step, batches, n, torch.cuda.current_stream())
        b = timed(step, batches, n, lo)
        c = timed(step, batches, n, hi)
        print(f"rep {rep}: default stream {a:.4f} ms  side stream prio 0 {b:.4f} ms  high-priority stream {c:.4f} ms", flush=True)


if __name__ == "__main__":
    main()
