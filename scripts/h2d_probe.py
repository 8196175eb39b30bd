#!/usr/bin/env python
"""Host -> device bandwidth of the bench's pinned host batches (11.5 MB per batch), idle and underneath the training
step, with the NUMA placement of the GPU and of this process: the host-fed bench leg reads 10 % slow on some runs."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def numa_info(dev_index):
    out = {}
    try:
        bdf = torch.cuda.get_device_properties(dev_index).pci_bus_id if hasattr(torch.cuda.get_device_properties(dev_index), "pci_bus_id") else None
    except Exception:
        bdf = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev_index)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bdf, bytes):
            bdf = bdf.decode()
    except Exception as e:
        out["nvml"] = repr(e)
    out["bdf"] = bdf
    if bdf:
        short = bdf.lower()[-12:]
        for cand in (short, "0000:" + short[-7:]):
            p = f"/sys/bus/pci/devices/{cand}/numa_node"
            if os.path.exists(p):
                out["gpu_numa_node"] = open(p).read().strip()
                break
    out["affinity"] = sorted(os.sched_getaffinity(0))
    nodes = {}
    base = "/sys/devices/system/node"
    if os.path.isdir(base):
        for n in sorted(os.listdir(base)):
            if n.startswith("node") and os.path.exists(f"{base}/{n}/cpulist"):
                nodes[n] = open(f"{base}/{n}/cpulist").read().strip()
    out["nodes"] = nodes
    return out


def main():
    print(numa_info(0), flush=True)
    step, dev_batches, host_batches = bench.make_step(batch=1024, return_host=True)
    dev = dev_batches[0]["x_atoms"].device
    cs = torch.cuda.Stream(dev)
    tensors = [v for k, v in host_batches[0].items() if isinstance(v, torch.Tensor) and k not in ("edge_attr", "cnx_attr", "x_frags")]
    dst = [torch.empty_like(t, device=dev) for t in tensors]
    nbytes = sum(t.numel() * t.element_size() for t in tensors)
    big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    big_d = torch.empty_like(big, device=dev)

    def h2d_times(n, busy):
        ts = []
        for i in range(n):
            if busy:
                step(dev_batches[i % 4])
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            with torch.cuda.stream(cs):
                e0.record(cs)
                for s, d in zip(tensors, dst):
                    d.copy_(s, non_blocking=True)
                e1.record(cs)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        torch.cuda.synchronize()
        return ts
    for rep in range(3):
        for busy in (False, True):
            ts = h2d_times(100, busy)
            gbs = [nbytes / t / 1e6 for t in ts]
            print(f"rep {rep} {'under the step' if busy else 'idle          '}: {nbytes / 1e6:.1f} MB in {statistics.median(ts):.3f} ms median "
                  f"({statistics.median(gbs):.1f} GB/s; min {min(gbs):.1f}, max {max(gbs):.1f})", flush=True)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); big_d.copy_(big, non_blocking=True); e1.record(); e1.synchronize()
        print(f"rep {rep} 256 MiB pinned copy: {big.numel() / e0.elapsed_time(e1) / 1e6:.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
