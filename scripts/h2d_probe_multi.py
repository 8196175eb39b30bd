#!/usr/bin/env python
"""All ranks copy a pinned 11.5 MB batch host -> device at the same time (torchrun): per-rank bandwidth and host time of
staging one batch, to tell PCIe / host-memory contention from host-CPU contention at 8 GPUs per box."""
import os
import statistics
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    src = torch.empty(11_518_208, dtype=torch.uint8).pin_memory()
    dst = torch.empty_like(src, device=dev)
    cs = torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    dist.barrier()
    ts, hs = [], []
    for i in range(300):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.perf_counter()
        with torch.cuda.stream(cs):
            e0.record(cs)
            for k in range(17):               # 17 tensors per batch in the real loop
                a, b = k * 677_000, (k + 1) * 677_000
                dst[a:b].copy_(src[a:b], non_blocking=True)
            e1.record(cs)
        hs.append((time.perf_counter() - t0) * 1e3)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
        time.sleep(0.0008)
    gbs = 17 * 677_000 / statistics.median(ts) / 1e6
    out = torch.tensor([gbs, statistics.median(hs), max(hs), os.cpu_count() or 0], device=dev)
    allv = [torch.zeros_like(out) for _ in range(dist.get_world_size())]
    dist.all_gather(allv, out)
    if rank == 0:
        for r, v in enumerate(allv):
            print(f"rank {r}: {v[0].item():6.1f} GB/s median device-side, host enqueue of 17 copies {v[1].item():.3f} ms median "
                  f"({v[2].item():.2f} max), cpus {int(v[3].item())}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
