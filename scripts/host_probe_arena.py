#!/usr/bin/env python
"""Host-side cost of the arena-fed step: time to ENQUEUE ``arena.batch`` and ``step`` (the GPU is drained in between),
and the free-running loop with a lagged loss read.  usage (GPU box): python scripts/host_probe_arena.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fragnet_b200 import synth  # noqa: E402
from fragnet_b200.dataset.arena import MoleculeArena  # noqa: E402
from fragnet_b200.train.fused import LaggedScalars  # noqa: E402


def main():
    step, dev_batches = bench.make_step(batch=1024)
    dev = dev_batches[0]["x_atoms"].device
    pool = synth.make_dataset("unimol", 512, seed=100)
    arena = MoleculeArena(pool, dev)
    rng = np.random.default_rng(0)
    ids = [rng.integers(0, len(pool), size=1024) for _ in range(4)]
    for i in range(5):
        step(arena.batch(ids[i % 4]))
    torch.cuda.synchronize()
    tb = ts = 0.0
    n = 20
    for i in range(n):
        t0 = time.perf_counter()
        b = arena.batch(ids[i % 4])
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        step(b)
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        tb += t1 - t0
        ts += t3 - t2
    print(f"host enqueue: arena.batch {1e6 * tb / n:.0f} us, step {1e6 * ts / n:.0f} us")
    for lag in (0, 1):
        reader = LaggedScalars(lag=lag)
        b = arena.batch(ids[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(60):
            loss = step(b)
            b = arena.batch(ids[(i + 1) % 4])
            reader.push(loss)
        reader.drain()
        torch.cuda.synchronize()
        print(f"lag {lag}: {1e3 * (time.perf_counter() - t0) / 60:.3f} ms per step (arena-fed)")
    for lag in (0, 1):
        reader = LaggedScalars(lag=lag)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(60):
            reader.push(step(dev_batches[i % 4]))
        reader.drain()
        torch.cuda.synchronize()
        print(f"lag {lag}: {1e3 * (time.perf_counter() - t0) / 60:.3f} ms per step (resident batches)")


if __name__ == "__main__":
    main()
