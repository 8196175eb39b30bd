#!/usr/bin/env python
"""In-situ DEVICE time of one training step: kineto (torch.profiler) kernel records of the bench step loop, warm and
back to back -- unlike the ncu launch list, whose per-launch times are cold-cache and serialised.  Prints the summed
kernel time per step by kernel name, the device-busy total, and the span from first launch to last completion."""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    shape = sys.argv[4] if len(sys.argv) > 4 else "unimol"
    step, batches = bench.make_step(batch=batch, shape=shape, autograd_path=len(sys.argv) > 3 and sys.argv[3] == "autograd")
    fused = hasattr(getattr(step, "__self__", None), "prefetch_plan")      # the bench loop: next batch's collate underneath

    def run(i):
        if fused:
            step(batches[i % len(batches)], next_batch=batches[(i + 1) % len(batches)])
        else:
            step(batches[i % len(batches)])
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(steps):
            run(i)
        torch.cuda.synchronize()
    by = collections.defaultdict(lambda: [0, 0.0])
    t_first, t_last, busy = None, None, 0.0
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CUDA:
            continue
        dur = e.device_time if hasattr(e, "device_time") else e.cuda_time
        name = e.name
        if "Memcpy" in name or "Memset" in name:
            name = name.split("(")[0].strip()
        by[name][0] += 1
        by[name][1] += dur
        busy += dur
        t0 = e.time_range.start
        t1 = e.time_range.end
        t_first = t0 if t_first is None else min(t_first, t0)
        t_last = t1 if t_last is None else max(t_last, t1)
    if os.environ.get("FNB_STREAMS_SEQ"):   # last step, in start order, with the CUDA stream of every kernel (kineto)
        try:
            kev = [e for e in prof.profiler.kineto_results.events()
                   if "cuda" in str(e.device_type()).lower() and e.duration_ns() > 0]
            kev.sort(key=lambda e: e.start_ns())
            per = len(kev) // steps
            last = kev[-per:]
            t0 = last[0].start_ns()
            ids = {}
            print("# start_us  dur_us  stream  kernel   (last step)")
            for e in last:
                sid = ids.setdefault(e.device_resource_id(), len(ids))
                nm = e.name().replace("(anonymous namespace)::", "").replace("void ", "")
                print(f"{(e.start_ns() - t0) / 1e3:9.1f} {e.duration_ns() / 1e3:7.1f}   s{sid}   {nm[:70]}")
        except Exception as ex:      # kineto internals differ between torch versions: the summary below still prints
            print("# stream timeline unavailable:", ex)
    if os.environ.get("FNB_SEQ"):      # launch sequence of the last profiled step, in start order
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                     key=lambda e: e.time_range.start)
        per = len(evs) // steps
        last = evs[-per:]
        t0 = last[0].time_range.start
        print("# start_us  dur_us  gap_us  kernel   (last step)")
        prev_end = t0
        for e in last:
            print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f} "
                  f"{e.time_range.start - prev_end:6.1f}  {e.name[:100]}")
            prev_end = e.time_range.end
    rows = sorted(by.items(), key=lambda kv: -kv[1][1])
    print(f"batch {batch}: device busy {busy / steps:.1f} us/step over {sum(v[0] for v in by.values()) / steps:.1f} "
          f"launches/step; span {(t_last - t_first) / steps:.1f} us/step")
    print(f"{'kernel':90s} {'n/step':>7s} {'avg us':>8s} {'us/step':>9s} {'share':>6s}")
    for name, (n, t) in rows[:70]:
        print(f"{name[:90]:90s} {n / steps:7.1f} {t / n:8.2f} {t / steps:9.1f} {100 * t / busy:5.1f}%")


if __name__ == "__main__":
    main()
