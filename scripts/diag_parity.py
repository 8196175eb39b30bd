#!/usr/bin/env python
"""Parity margins of the pretraining model on the golden batch, per precision mode: rel_err of every prediction tensor,
the loss, and the worst gradient (tests/test_gpu_model.py::test_pretrain_step_matches_oracle_and_golden)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from conftest import GOLDEN, grad_errs, rel_err  # noqa: E402
from test_gpu_model import _rebuild, _to  # noqa: E402


def main():
    from make_golden import golden_batch
    from fragnet_b200 import config
    from fragnet_b200.train.pretrain_utils import pretrain_loss
    golden = torch.load(GOLDEN)
    gb = golden_batch()
    for prec in ("fp32_simt", "fp32", "tf32"):
        config.set_precision(prec)
        m = _rebuild(golden, "pt").cuda()
        bc = _to(gb, "cuda")
        preds = m(bc)
        names = ("bond_length", "bond_angle", "dihedral", "energy")
        errs = {n: rel_err(a, r) for n, a, r in zip(names, preds, golden["pt_preds"])}
        loss = pretrain_loss(torch.nn.MSELoss(), preds, bc)
        loss.backward()
        named = dict(m.named_parameters())
        ge = grad_errs([(k, named[k].grad, g) for k, g in golden["pt_grads"].items()])
        worst = max(ge, key=ge.get)
        print(prec, {k: f"{v:.2e}" for k, v in errs.items()}, "loss", f"{rel_err(loss, golden['pt_loss']):.2e}",
              "worst grad", worst, f"{ge[worst]:.2e}")
        mf = _rebuild(golden, "ft").cuda()
        pred = mf(bc)
        with torch.no_grad():
            enc = mf.pretrain.forward_with_attention(bc)
        print("   ft_pred", f"{rel_err(pred, golden['ft_pred']):.2e}",
              {n: f"{rel_err(t, golden['encoder'][n]):.2e}" for n, t in zip(golden["encoder"], enc)})
    config.set_precision("fp32")


if __name__ == "__main__":
    main()
