"""Pretraining step loop with the reference's semantics (fragnet/train/pretrain/pretrain_utils.py:4-57).

Kept behaviour: every batch tensor is moved with ``.to(device)``; the loss is
``2*MSE(dihedral) + MSE(angle) + MSE(energy)`` because the reference overwrites ``loss_lngth`` with the
dihedral term before summing (pretrain_utils.py:22-26); the epoch loss is divided by the dataset size.
"""
from __future__ import annotations

import torch


def pretrain_loss(loss_fn, preds, batch):
    _, angle, dihedral, energy = preds
    l_dh = loss_fn(dihedral, batch["dh_angl"])
    return l_dh + loss_fn(angle, batch["bnd_angl"]) + l_dh + loss_fn(energy.view(-1), batch["y"])


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
