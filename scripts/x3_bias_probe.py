#!/usr/bin/env python
"""Signed error of the tensor-core projections against float64: is there a systematic (round-toward-zero) bias in the
accumulation?  Positive operands make every partial sum positive, so a truncating accumulator shows up as a negative
mean relative error that grows with K."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from fragnet_b200 import ops
    g = torch.Generator().manual_seed(0)
    for K in (32, 128, 256):
        for kind in ("positive", "signed"):
            x = torch.rand(4096, K, generator=g) if kind == "positive" else torch.randn(4096, K, generator=g)
            W = torch.rand(128, K, generator=g) if kind == "positive" else torch.randn(128, K, generator=g)
            ref = x.double() @ W.double().t()
            for prec, name in ((0, "ffma"), (1, "tf32"), (2, "3xtf32")):
                h, _ = ops.proj_fwd(x.cuda(), W.cuda(), None, None, want_S=False, precision=prec)
                e = (h.cpu().double() - ref) / ref.abs().clamp_min(1e-30)
                big = ref.abs() > 0.1 * ref.abs().max()
                print(f"K={K:4d} {kind:8s} {name:7s} mean rel err {float(e[big].mean()):+.3e}  rms {float(e[big].pow(2).mean().sqrt()):.3e}"
                      f"  max-norm {float((h.cpu().double() - ref).abs().max() / ref.abs().max()):.3e}")


if __name__ == "__main__":
    main()
