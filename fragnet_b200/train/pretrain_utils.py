"""Pretraining step loop with the reference's semantics (fragnet/train/pretrain/pretrain_utils.py:4-57).

Kept behaviour: every batch tensor is moved with ``.to(device)``; the loss is
``2*MSE(dihedral) + MSE(angle) + MSE(energy)`` because the reference overwrites ``loss_lngth`` with the
dihedral term before summing (pretrain_utils.py:22-26); the epoch loss is divided by the dataset size.
"""
from __future__ import annotations

import torch


def pretrain_loss(loss_fn, preds, batch):
    """``loss_fn(dihedral) + loss_fn(angle) + loss_fn(dihedral) + loss_fn(energy)`` (pretrain_utils.py:22-26).  With
    the reference's ``nn.MSELoss()`` on CUDA tensors the four terms, their sum and the gradients of the predictions
    come from one ``fnb_mse_sum_loss`` launch."""
    _, angle, dihedral, energy = preds
    t_dh, t_ba, y = batch["dh_angl"], batch["bnd_angl"], batch["y"]
    if (isinstance(loss_fn, torch.nn.MSELoss) and loss_fn.reduction == "mean" and dihedral.is_cuda
            and all(t.is_cuda and t.dtype == torch.float32 for t in (t_dh, t_ba, y))
            and dihedral.shape == t_dh.shape and angle.shape == t_ba.shape and energy.numel() == y.numel()
            and min(dihedral.numel(), angle.numel(), energy.numel()) > 0):
        from ..autograd import MseSumLossFn
        return MseSumLossFn.apply((2.0, 1.0, 1.0), dihedral, t_dh, angle, t_ba, energy.view(-1), y)
    l_dh = loss_fn(dihedral, t_dh)
    return l_dh + loss_fn(angle, t_ba) + l_dh + loss_fn(energy.view(-1), y)


class Trainer:
    def __init__(self, loss_fn=None):
        self.loss_fn = loss_fn

    def train(self, model, loader, optimizer, device):
        model.train()
        total = 0.0
        for batch in loader:
            for k in batch:
                batch[k] = batch[k].to(device)
            optimizer.zero_grad()
            loss = pretrain_loss(self.loss_fn, model(batch), batch)
            loss.backward()
            total += loss.item()
            optimizer.step()
        return total / len(loader.dataset)

    def validate(self, loader, model, device):
        model.eval()
        total = 0.0
        with torch.no_grad():
            for batch in loader:
                for k in batch:
                    batch[k] = batch[k].to(device)
                total += pretrain_loss(self.loss_fn, model(batch), batch).item()
        return total / len(loader.dataset)
