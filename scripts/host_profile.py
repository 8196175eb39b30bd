#!/usr/bin/env python
"""Where does the HOST time of one training step go?  cProfile over the bench step loop (GPU box)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from fragnet.model.gat.gat2_pretrain import FragNetPreTrain
    from fragnet_b200 import config, ops
    from fragnet_b200.dist import FlatGradSync
    from fragnet_b200.train.optim import FlatAdam
    from fragnet_b200.train.pretrain_utils import pretrain_loss
    dev = torch.device("cuda", 0)
    config.set_precision("tf32")
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(1234)
    model = FragNetPreTrain(**bench.PT_KW).to(dev).train()
    loss_fn = torch.nn.MSELoss()
    batches = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_batches("unimol", 1024, 4, 512, seed=100)]
    sync = FlatGradSync(model.parameters())
    opt = None

    def step(batch):
        nonlocal opt
        ops.clear_plan_cache()
        sync.zero()
        loss = pretrain_loss(loss_fn, model(batch), batch)
        loss.backward()
        sync.sync()
        if opt is None:
            opt = FlatAdam(sync.live_parameters(), lr=1e-4)
        opt.step()
        return loss

    for i in range(5):
        step(batches[i % 4])
    torch.cuda.synchronize()
    # host time to ENQUEUE 20 steps (no sync inside) vs device time
    t0 = time.perf_counter()
    for i in range(20):
        step(batches[i % 4])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3 * (t1 - t0) / 20:.3f} ms/step, wall incl. drain {1e3 * (t2 - t0) / 20:.3f} ms/step")
    # phases, each synchronised (device time of the phase when the host is not the bottleneck)
    def timed(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        th = time.perf_counter()
        torch.cuda.synchronize()
        return r, 1e3 * (th - t), 1e3 * (time.perf_counter() - t)
    acc = {}
    for i in range(10):
        b = batches[i % 4]
        ops.clear_plan_cache()
        sync.zero()
        enc, h1, w1 = timed(lambda: model.pretrain(b))
        preds, h2, w2 = timed(lambda: model.head(enc[0], enc[1], enc[2], b))
        loss, h3, w3 = timed(lambda: pretrain_loss(loss_fn, preds, b))
        _, h4, w4 = timed(lambda: loss.backward())
        _, h5, w5 = timed(lambda: (sync.sync(), opt.step()))
        for k, v in (("encoder fwd", (h1, w1)), ("heads fwd", (h2, w2)), ("loss", (h3, w3)), ("backward", (h4, w4)),
                     ("optimizer", (h5, w5))):
            a = acc.setdefault(k, [0.0, 0.0])
            a[0] += v[0] / 10
            a[1] += v[1] / 10
    for k, (h, w) in acc.items():
        print(f"{k:12s} host {h:7.3f} ms   host+device {w:7.3f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for i in range(20):
        step(batches[i % 4])
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)


if __name__ == "__main__":
    main()
