#!/usr/bin/env python
"""Diagnostic: where do the fused heads' gradients differ from the separate-linears reference in the fp32 (3xTF32)
mode -- a few rows whose ReLU mask flipped, or everywhere?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_heads import _batch, _head, _inputs  # noqa: E402


def main():
    from fragnet_b200 import config
    for prec in ("fp32_simt", "fp32"):
        config.set_precision(prec)
        b = _batch("unimol", 300, 11)
        head = _head(5)
        xa, xf, xe = _inputs(b, 7)
        outs = head(xa, xf, xe, b)
        gen = torch.Generator().manual_seed(3)
        ws = [torch.randn(o.shape, generator=gen).cuda() for o in outs]
        sum((o * w).sum() for o, w in zip(outs[1:], ws[1:])).backward()
        got = xe.grad.clone()
        xe.grad = None
        xa.grad = None
        xf.grad = None
        head.zero_grad(set_to_none=True)
        ref = head._forward_linears(xa, xf, xe, b)
        sum((o * w).sum() for o, w in zip(ref[1:], ws[1:])).backward()
        d = (got - xe.grad).abs().amax(1)
        scale = float(xe.grad.abs().max())
        bad = (d > 1e-5 * scale).nonzero().flatten()
        print(prec, "rows", got.shape[0], "rows off by > 1e-5:", bad.numel(), "max", float(d.max()) / scale,
              "median row err", float(d.median()) / scale)
        # pre-activations of the dihedral head's first layer for the bad rows
        h0 = torch.nn.functional.linear(xe.detach().double(), head.da_layers[0].weight.double(), head.da_layers[0].bias.double())
        for r in bad[:5].tolist():
            print("   row", r, "min |pre-activation| layer 0:", float(h0[r].abs().min()))


if __name__ == "__main__":
    main()
