#!/usr/bin/env python
"""Host time to ENQUEUE one training step (GPU box): wall clock of 4 step calls issued into an empty queue, so that the
host never waits for the device.  Splits the library call, the optimizer launch, the next batch's collate and staging."""
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from fragnet_b200.dataset.prefetch import DevicePrefetcher
    step, dev_batches, host_batches = bench.make_step(batch=1024, return_host=True)
    fs = step.__self__
    for i in range(10):
        step(dev_batches[i % 4], next_batch=dev_batches[(i + 1) % 4])
    torch.cuda.synchronize()

    def measure(fn, reps=12, n=4):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(n):
                fn(i)
            ts.append((time.perf_counter() - t0) / n * 1e3)
        return statistics.median(ts), min(ts)
    print("step only (library call + Adam launch)      : %.3f ms median, %.3f min" % measure(lambda i: fs._step(dev_batches[i % 4])))
    print("step + prefetch_plan of the next batch      : %.3f ms median, %.3f min" %
          measure(lambda i: step(dev_batches[i % 4], next_batch=dev_batches[(i + 1) % 4])))
    print("prefetch_plan alone                         : %.3f ms median, %.3f min" % measure(lambda i: fs.prefetch_plan(dev_batches[i % 4])))
    staged = DevicePrefetcher((host_batches[i % 4] for i in range(10 ** 6)), dev_batches[0]["x_atoms"].device, depth=2,
                              hot_path_only=True)
    feed = iter(staged)
    next(feed)
    print("DevicePrefetcher next() (stage one batch)   : %.3f ms median, %.3f min" % measure(lambda i: next(feed)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(100):
        step(dev_batches[i % 4], next_batch=dev_batches[(i + 1) % 4])
    e1.record()
    e1.synchronize()
    print("device time per step (resident)             : %.3f ms" % (e0.elapsed_time(e1) / 100))


if __name__ == "__main__":
    main()
