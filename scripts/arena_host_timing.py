#!/usr/bin/env python
"""Where the host time of ``MoleculeArena.batch`` goes (per phase, GPU drained between calls)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fragnet_b200 import synth  # noqa: E402
from fragnet_b200.dataset import arena as A  # noqa: E402


def main():
    pool = synth.make_dataset("unimol", 512, seed=100)
    ar = A.MoleculeArena(pool, "cuda")
    rng = np.random.default_rng(0)
    ids = [rng.integers(0, len(pool), size=1024) for _ in range(4)]
    for i in range(8):
        ar.batch(ids[i % 4])
    torch.cuda.synchronize()
    T = {}

    def timed(name, fn):
        t = time.perf_counter()
        r = fn()
        T[name] = T.get(name, 0.0) + time.perf_counter() - t
        return r

    n = 20
    lib_call = ar._lib.fnb_arena_assemble
    for i in range(n):
        x = ids[i % 4]
        timed("stage_ids", lambda: ar._stage_ids(x))
        timed("sums", lambda: ar._count_matrix[:, x].sum(axis=1))
        timed("empty_f32", lambda: torch.empty(20_000_000 + i * 1000, dtype=torch.float32, device="cuda"))
        timed("empty_i64", lambda: torch.empty(1_200_000 + i * 1000, dtype=torch.int64, device="cuda"))
        timed("whole_batch", lambda: ar.batch(x))
        torch.cuda.synchronize()
    for k, v in T.items():
        print(f"{k:12s} {1e6 * v / n:8.0f} us")
    print("allocator:", {k: v for k, v in torch.cuda.memory_stats().items() if k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "reserved_bytes.all.current")})


if __name__ == "__main__":
    main()
