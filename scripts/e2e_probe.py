#!/usr/bin/env python
"""End-to-end feed variants of the fused step (GPU box): blocking .to(), prefetcher depth 1/2, with per-step loss read."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from fragnet_b200.dataset.prefetch import DevicePrefetcher

step, dev_batches, host_batches = bench.make_step(return_host=True)
dev = dev_batches[0]["x_atoms"].device
N = 40
def run(name, feed):
    for _ in range(5):
        step(next(feed)).item()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        step(next(feed)).item()
    torch.cuda.synchronize()
    print(f"{name:40s} {1e3 * (time.perf_counter() - t0) / N:.3f} ms/step", flush=True)
def blocking():
    i = 0
    while True:
        yield {k: v.to(dev, non_blocking=True) for k, v in host_batches[i % 4].items()}
        i += 1
def resident():
    i = 0
    while True:
        yield dev_batches[i % 4]
        i += 1
run("device-resident batches", resident())
run("same-stream .to(non_blocking)", blocking())
for d in (1, 2, 3):
    run(f"DevicePrefetcher depth {d}", iter(DevicePrefetcher((host_batches[i % 4] for i in range(10 ** 6)), dev, depth=d)))

def run_pipelined(name, feed):
    b = next(feed)
    for _ in range(5):
        loss = step(b); b = next(feed); loss.item()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        loss = step(b); b = next(feed); loss.item()
    torch.cuda.synchronize()
    print(f"{name:40s} {1e3 * (time.perf_counter() - t0) / N:.3f} ms/step", flush=True)
for d in (1, 2):
    run_pipelined(f"DevicePrefetcher depth {d}, stage after launch", iter(DevicePrefetcher((host_batches[i % 4] for i in range(10 ** 6)), dev, depth=d)))
