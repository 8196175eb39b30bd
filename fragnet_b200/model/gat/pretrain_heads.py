"""Pretraining heads and the ``FragNetPreTrain`` wrapper (reference fragnet/model/gat/pretrain_heads.py).

The four small MLP heads stay ``nn.Linear`` (they are dense library GEMMs); the
graph readout in front of the energy head runs through the segment-sum kernel.  Parameter names,
shapes and registration order equal the reference's (``head.bl_reduce_layer``, ``head.bl_layers.N``,
``head.ba_layers.N``, ``head.da_layers.N``, ``head.FC_layers.N``), so checkpoints interchange.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import config, ops
from ...autograd import PretrainHeadsFn
from .gat2 import FragNet, graph_readout


def _halving_stack(width: int, depth: int, dim_out: int) -> nn.ModuleList:
    """``depth`` linears halving the width each time, then one to ``dim_out`` (pretrain_heads.py:27-57)."""
    dims = [width // 2 ** l for l in range(depth + 1)] + [dim_out]
    return nn.ModuleList([nn.Linear(a, b, bias=True) for a, b in zip(dims[:-1], dims[1:])])


class PretrainTask(nn.Module):
    """Bond-length, bond-angle, dihedral-angle and energy heads (pretrain_heads.py:8-102)."""

    def __init__(self, dim_in=128, dim_out=1, L=2):
        super().__init__()
        self.bl_reduce_layer = nn.Linear(dim_in * 3, dim_in)
        self.bl_layers = _halving_stack(dim_in, L, dim_out)
        self.ba_layers = _halving_stack(dim_in, L, dim_out)
        self.da_layers = _halving_stack(dim_in, L, dim_out)
        self.FC_layers = _halving_stack(dim_in * 2, L, dim_out)
        self.L = L
        self.activation = nn.ReLU()

    def _apply_index(self, batch):
        return batch.bond_length, batch.distance

    def _tail(self, layers, x):
        for lin in list(layers)[:-1]:
            x = self.activation(lin(x))
        return layers[-1](x)

    def _parameters_in_library_order(self):
        """Wr, br, then (W0, b0, W1, b1, W2, b2) of the bond-length, bond-angle, dihedral and energy stacks
        (``fnb_pretrain_head_params``)."""
        out = [self.bl_reduce_layer.weight, self.bl_reduce_layer.bias]
        for stack in (self.bl_layers, self.ba_layers, self.da_layers, self.FC_layers):
            for lin in stack:
                out += [lin.weight, lin.bias]
        return out

    def _library_shapes(self) -> bool:
        """The fused programs cover the geometry the reference ships (dim_in 128, L 2, dim_out 1)."""
        return (self.L == 2 and tuple(self.bl_reduce_layer.weight.shape) == (128, 384)
                and tuple(self.FC_layers[0].weight.shape) == (128, 256) and self.FC_layers[2].weight.shape[0] == 1)

    def _forward_linears(self, x_atoms, x_frags, edge_attr, batch):
        """The same heads as separate library GEMMs (``nn.Linear``): used for geometries the fused programs do not
        cover and when a gradient is requested for the bond-length output."""
        ei = batch["edge_index"].to(x_atoms.device)
        # [x_begin | x_end | bond features] per directed bond (pretrain_heads.py:67-70)
        pair = torch.cat((x_atoms.index_select(0, ei[0]), x_atoms.index_select(0, ei[1]), edge_attr), dim=1)
        bond_length = self.bl_reduce_layer(pair)
        for lin in self.bl_layers:                       # activation BEFORE each linear (:72-74)
            bond_length = lin(self.activation(bond_length))
        bond_angle = self._tail(self.ba_layers, x_atoms)
        dihedral = self._tail(self.da_layers, edge_attr)
        energy = self._tail(self.FC_layers, graph_readout(x_atoms, x_frags, batch))
        return bond_length, bond_angle, dihedral, energy

    def forward(self, x_atoms, x_frags, edge_attr, batch, bond_length_grad: bool = False):
        """One ``fnb_pretrain_heads_forward`` call (and one backward call under autograd).  ``bond_length_grad=True``
        keeps the bond-length output differentiable (separate GEMMs); the reference's training loss never
        differentiates it (pretrain_utils.py:24 overwrites ``loss_lngth``)."""
        if bond_length_grad or not self._library_shapes():
            return self._forward_linears(x_atoms, x_frags, edge_attr, batch)
        home = x_atoms.device
        dev = ops.require_cuda(home if home.type == "cuda" else self.bl_reduce_layer.weight.device)
        on = lambda t: t if t.device == dev else t.to(dev)
        rp = ops.readout_plan_for(batch["batch"], batch["frag_batch"], dev)
        outs = PretrainHeadsFn.apply(rp, on(batch["edge_index"]), config.precision_id(), torch.is_grad_enabled(),
                                     on(x_atoms), on(x_frags), on(edge_attr),
                                     *[on(p) for p in self._parameters_in_library_order()])
        bond_length, bond_angle, dihedral, energy = outs
        bond_length = bond_length.detach()
        if home.type != "cuda":
            return tuple(t.to(home) for t in (bond_length, bond_angle, dihedral, energy))
        return bond_length, bond_angle, dihedral, energy


class FragNetPreTrain(nn.Module):
    """Encoder + pretraining heads (pretrain_heads.py:105-141, gat2_pretrain.py:7-27)."""

    def __init__(self, num_layer=4, drop_ratio=0.15, num_heads=4, emb_dim=128, atom_features=167,
                 frag_features=167, edge_features=16, fedge_in=6, fbond_edge_in=6):
        super().__init__()
        self.pretrain = FragNet(num_layer=num_layer, drop_ratio=drop_ratio, num_heads=num_heads, emb_dim=emb_dim,
                                atom_features=atom_features, frag_features=frag_features,
                                edge_features=edge_features, fedge_in=fedge_in, fbond_edge_in=fbond_edge_in)
        self.head = PretrainTask(128, 1)

    def forward(self, batch):
        x_atoms, x_frags, e_edge, _ = self.pretrain(batch)
        return self.head(x_atoms, x_frags, e_edge, batch)
