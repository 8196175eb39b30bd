#!/usr/bin/env python
"""BASELINE configs[2]: inference-only batched screening -- UniMol-shaped synthetic molecules, batch 4096, eval mode,
the last layer's attention weights returned (the arrangement of fragnet/vizualize/model.py:72-142) and copied back to
the host.  Prints molecules/s for (a) device-resident batches and (b) host batches through DevicePrefetcher with the
predictions + four attention tensors read back every batch.

usage (GPU box): python scripts/infer_bench.py [--batch 4096] [--batches 12] [--shape unimol]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--batches", type=int, default=12)
    ap.add_argument("--shape", default="unimol")
    ap.add_argument("--precision", default="tf32")
    args = ap.parse_args()
    import bench
    from fragnet.model.gat.gat2 import FragNetFineTune
    from fragnet_b200 import config, ops
    from fragnet_b200.dataset.prefetch import DevicePrefetcher
    from fragnet_b200.model.gat.gat2 import graph_readout
    config.set_precision(args.precision)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = FragNetFineTune(num_layer=4, drop_ratio=0.1, n_classes=1).to(dev).eval()
    host = bench.make_batches(args.shape, args.batch, 4, 512, seed=7)
    for b in host:
        for k in b:
            b[k] = b[k].pin_memory()
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]

    @torch.no_grad()
    def infer(b):
        ops.clear_plan_cache()
        enc = model.pretrain.forward_with_attention(b)           # atoms, frags, bonds, fbonds, then 4 attention sums
        pred = model.fthead(graph_readout(enc[0], enc[1], b))
        return (pred,) + tuple(enc[4:])

    for i in range(3):
        infer(devb[i % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.batches):
        infer(devb[i % 4])
    torch.cuda.synchronize()
    t_res = (time.perf_counter() - t0) / args.batches
    # end to end: pinned host batches in, predictions + attention weights out (pinned host buffers)
    feed = iter(DevicePrefetcher((host[i % 4] for i in range(args.batches + 3)), dev, depth=2, hot_path_only=True))
    outs_host = None
    d2h = 0
    t0 = None
    b = next(feed)
    for i in range(args.batches + 3):
        if i == 3:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        outs = infer(b)
        b = next(feed, None)
        if outs_host is None or any(o.shape != h.shape for o, h in zip(outs, outs_host)):
            outs_host = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
        for o, h in zip(outs, outs_host):
            h.copy_(o, non_blocking=True)
        torch.cuda.synchronize()
        d2h = sum(o.numel() * 4 for o in outs)
    t_e2e = (time.perf_counter() - t0) / args.batches
    h2d = sum(v.numel() * v.element_size() for k, v in host[0].items() if k not in ("edge_attr", "cnx_attr", "x_frags"))
    # end to end through the screening pipeline (fragnet_b200.screen): dataset resident in HBM (MoleculeArena), per
    # batch: molecule ids in, on-device assembly, forward, predictions + attention weights copied to pinned host
    # buffers on a side stream (one batch behind)
    import numpy as np
    from fragnet.vizualize.model import FragNetFineTuneViz
    from fragnet_b200 import synth
    from fragnet_b200.dataset.arena import MoleculeArena
    from fragnet_b200.screen import screen
    viz = FragNetFineTuneViz(num_layer=4, drop_ratio=0.1, n_classes=1, edge_features=17).to(dev).eval()
    viz.load_state_dict(model.state_dict(), strict=True)
    pool = synth.make_dataset(args.shape, 512, seed=7, with_pretrain_targets=False)
    arena = MoleculeArena(pool, dev, pretrain=False)
    rng = np.random.default_rng(7)
    n_warm = 3
    ids = rng.integers(0, len(pool), size=args.batch * (args.batches + n_warm))
    t0, n_out = None, 0
    for i, (bid, outs) in enumerate(screen(viz, arena, batch_size=args.batch, ids=ids)):
        if i == n_warm - 1:
            t0 = time.perf_counter()
        n_out += int(outs[0].shape[0])
    t_arena = (time.perf_counter() - t0) / args.batches
    assert n_out == ids.shape[0]
    print(json.dumps({"workload": f"inference screening, {args.shape}-shaped, batch {args.batch}, eval, attention returned",
                      "precision": args.precision, "molecules_per_s_resident": round(args.batch / t_res, 1),
                      "ms_per_batch_resident": round(1e3 * t_res, 3), "molecules_per_s_e2e": round(args.batch / t_e2e, 1),
                      "ms_per_batch_e2e": round(1e3 * t_e2e, 3), "h2d_bytes_per_batch": h2d, "d2h_bytes_per_batch": d2h,
                      "seconds_per_1M_molecules_e2e": round(1e6 / (args.batch / t_e2e), 2),
                      "molecules_per_s_e2e_screen": round(args.batch / t_arena, 1),
                      "ms_per_batch_e2e_screen": round(1e3 * t_arena, 3), "h2d_bytes_per_batch_screen": args.batch * 8,
                      "seconds_per_1M_molecules_e2e_screen": round(1e6 / (args.batch / t_arena), 2)}))


if __name__ == "__main__":
    main()
