#!/usr/bin/env python
"""Where do the occasional slow end-to-end runs come from?  Runs the bench's host-fed loop, records the host clock after
every iteration and the device time of every 10 steps, and prints the largest gaps."""
import gc
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from fragnet_b200.dataset.prefetch import DevicePrefetcher
    from fragnet_b200.train.fused import LaggedScalars
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    step, dev_batches, host_batches = bench.make_step(batch=1024, return_host=True)
    fs = step.__self__
    dev = dev_batches[0]["x_atoms"].device
    reader = LaggedScalars(lag=1)
    if os.environ.get("NO_GC"):
        gc.disable()
    for rep in range(3):
        staged = DevicePrefetcher((host_batches[i % 4] for i in range(n + 30)), dev, depth=2, hot_path_only=True)
        feed = iter(staged)
        b = next(feed)
        for i in range(30):
            loss = step(b)
            b = next(feed)
            fs.prefetch_plan(b, staged.last_event)
            reader.push(loss)
        torch.cuda.synchronize()
        stamps, evs = [time.perf_counter()], []
        e0 = torch.cuda.Event(True); e0.record()
        for i in range(n):
            loss = step(b)
            b = next(feed, None)
            if b is not None:
                fs.prefetch_plan(b, staged.last_event)
            reader.push(loss)
            stamps.append(time.perf_counter())
            if i % 10 == 9:
                e = torch.cuda.Event(True); e.record(); evs.append(e)
        reader.drain()
        torch.cuda.synchronize()
        gaps = sorted(((stamps[i + 1] - stamps[i]) * 1e3, i) for i in range(n))[-5:]
        dts = [(([e0] + evs)[k].elapsed_time(evs[k]) / 10, k * 10) for k in range(len(evs))]
        print(f"rep {rep}: {1e3 * (stamps[-1] - stamps[0]) / n:.3f} ms/step host loop; device {e0.elapsed_time(evs[-1]) / n:.3f} ms/step; "
              f"largest host gaps (ms, step) {[(round(g, 2), i) for g, i in gaps]}; slowest 10-step windows (ms/step, step) "
              f"{[(round(d, 2), i) for d, i in sorted(dts)[-3:]]}; gc counts {gc.get_count()}", flush=True)


if __name__ == "__main__":
    main()
