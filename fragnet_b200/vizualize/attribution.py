"""Batched mask attribution.

The reference's interpretability code measures the contribution of every atom, bond and fragment link of ONE molecule
by deleting it and predicting again: ``FragNetVizApp.calc_atom_contributions / calc_bond_contributions /
calc_fbond_contributions`` (fragnet/vizualize/viz.py:901-1107) loop over the items and, per item, deep-copy the model
(``_mask_prediction*``, viz.py:960-984, 1026-1050, 1145-1169), set ``atom_mask_individual`` / ``bond_mask`` /
``frag_bond_mask`` on every layer and run a batch-1 forward: ``Na + Ea/2 + Ef/2 + 3`` forwards and model copies per
molecule.

Molecules of a batch are independent (block-diagonal graphs, no cross-molecule op on the gat2 path), so here the
molecule is replicated once per mask in ONE batch, replica ``r`` carries mask ``r`` as row lists
(``fnb_layer_params.atom_mask_list / bond_mask_rows / fbond_mask_rows``), and a single forward returns every masked
prediction.  The values equal the reference's loop (same arithmetic per replica; checked against per-mask forwards and
the CPU restatement in tests/test_gpu_attribution.py).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from ..dataset.data import collate_fn


def mask_predictions(model, data_item, collate: Callable = collate_fn, atoms: bool = True, bonds: bool = True,
                     fbonds: bool = True, max_replicas: Optional[int] = 4096) -> Dict[str, torch.Tensor]:
    """Predictions of ``model`` (a ``FragNetFineTune``-like module with ``.pretrain.layers``) for ``data_item`` with no
    mask and with each single mask applied in every layer.

    Returns ``{"pred_no_mask": [n_classes], "atom": [Na, n_classes], "bond": [Ea/2, n_classes] (bond pair starting at
    row 2j), "fbond": [Ef/2, n_classes] (fragment link k; empty for a single-fragment molecule, viz.py:1079-1082)}``;
    attribution = ``pred_no_mask - masked`` as in viz.py:926, 1004, 1097."""
    dev = next(model.parameters()).device
    na = int(data_item.x_atoms.shape[0])
    ea = int(data_item.edge_index.shape[1])
    nfb = int(data_item.node_feautures_fbondg.shape[0])
    single_frag = int(data_item.n_frags.item()) == 1
    jobs = [("none", 0)]
    if atoms:
        jobs += [("atom", i) for i in range(na)]
    if bonds:
        jobs += [("bond", i) for i in range(0, ea, 2)]
    if fbonds and not single_frag:
        jobs += [("fbond", i // 2) for i in range(0, nfb, 2)]
    preds = []
    layers = list(model.pretrain.layers)
    saved = [(l.atom_mask_individual, l.bond_mask, l.frag_bond_mask) for l in layers]
    was_training = model.training
    model.eval()
    step = len(jobs) if not max_replicas else max(1, int(max_replicas))
    try:
        for c0 in range(0, len(jobs), step):
            chunk = jobs[c0:c0 + step]
            batch = collate([data_item] * len(chunk))
            batch = {k: v.to(dev) for k, v in batch.items()}
            # replica r owns atoms [r*Na, (r+1)*Na), bond rows [r*Ea, ...), fragment-link pairs [r*Nfb/2, ...)
            am = [r * na + i for r, (kind, i) in enumerate(chunk) if kind == "atom"]
            bm = [r * ea + i for r, (kind, i) in enumerate(chunk) if kind == "bond"]
            fm = [r * (nfb // 2) + i for r, (kind, i) in enumerate(chunk) if kind == "fbond"]
            for l in layers:
                l.atom_mask_individual = torch.tensor(am, dtype=torch.int64, device=dev) if am else None
                l.bond_mask = torch.tensor(bm, dtype=torch.int64, device=dev) if bm else None
                l.frag_bond_mask = torch.tensor(fm, dtype=torch.int64, device=dev) if fm else None
            with torch.no_grad():
                out = model(batch)
            preds.append(out[0] if isinstance(out, (tuple, list)) else out)
    finally:
        for l, (a, b, f) in zip(layers, saved):
            l.atom_mask_individual, l.bond_mask, l.frag_bond_mask = a, b, f
        model.train(was_training)
    pred = torch.cat(preds, dim=0)
    kinds = [k for k, _ in jobs]
    pick = lambda name: pred[[i for i, k in enumerate(kinds) if k == name]]
    return {"pred_no_mask": pred[0], "atom": pick("atom"), "bond": pick("bond"), "fbond": pick("fbond")}


def attributions(model, data_item, **kw) -> Dict[str, torch.Tensor]:
    """``pred_no_mask - pred_mask`` per atom / bond / fragment link (the ``attr`` column of viz.py:926-936)."""
    p = mask_predictions(model, data_item, **kw)
    return {k: p["pred_no_mask"].unsqueeze(0) - p[k] for k in ("atom", "bond", "fbond")}
