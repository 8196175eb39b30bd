"""Drop-in ``gat2_lite``: the two-graph variant (bond graph + atom graph + atom->fragment pooling).

Mirrors the reference's ``fragnet/model/gat/gat2_lite.py`` -- ``FragNetLayerA`` (:13-150: the bond-graph block, the
atom-graph block and the pooling of ``gat2.py`` without the fragment-connection and fragment blocks; 8-argument
``forward``), ``FragNet`` (:153-216), ``FragNetFineTune`` (:467-509) -- selected by ``model_version: gat2_lite``
(train/finetune/finetune_gat2.py:141-160).  Parameter names, shapes and registration order are those of the reference
(identical to ``gat2``: the fragment-side parameters exist in the ``state_dict`` and never receive a gradient).

The arithmetic is the same sm_100a encoder program as ``gat2`` run over a batch plan whose
fragment-connection and fragment graphs are empty; pooling and the inter-layer ``ReLU(Dropout(.))`` of the last layer
use the library's segment-sum / dropout kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import config, ops
from ...autograd import DropoutReluFn, EncoderConfig, EncoderFn, PoolFn
from . import gat2 as _g
from .gat2 import FTHead1, FTHead2, FTHead3, FTHead4, FTHead5, graph_readout  # noqa: F401  (same classes as gat2)


def _empty_side(dev, k_fbond):
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)
    return z(0, k_fbond), z(2, 0, dt=torch.long), z(0, 6)      # fbond nodes, an empty [2,0] index, fbond edge attributes


class FragNetLayerA(_g.FragNetLayerA):
    """gat2_lite.py:13-150.  ``forward`` takes the 8 tensors of the reference and returns
    ``(x_atoms_new, x_frags, new_bond_features, None)`` (+ ``attn_atoms, None, attn_bonds, None``)."""

    def __init__(self, atom_in=128, atom_out=128, frag_in=128, frag_out=128, edge_in=128, edge_out=128, fedge_in=128,
                 num_heads=2, bond_edge_in=1, fbond_edge_in=8, return_attentions=False, add_frag_self_loops=False):
        super().__init__(atom_in, atom_out, frag_in, frag_out, edge_in, edge_out, fedge_in, num_heads, bond_edge_in,
                         fbond_edge_in, return_attentions, add_frag_self_loops)

    def _check_geometry(self):
        atom_out, edge_out, heads, bond_edge_in, _ = self._geometry     # the fragment-connection width is unused here
        if (atom_out, edge_out, heads, bond_edge_in) != (ops.D, ops.D, ops.H, 1):
            raise NotImplementedError("fragnet_b200 kernels are specialised for emb_dim=128, num_heads=4, "
                                      f"bond_edge_in=1; got {self._geometry}")

    def _lite_parameters(self, dev):
        """The 14 tensors the encoder program takes; the fragment-side ones are detached (gat2_lite never reads them)."""
        on = lambda t: t if t.device == dev else t.to(dev)
        live = {id(p) for p in (self.projection_b.weight, self.projection_b.bias, self.edge_attr_bond_embed.weight,
                                self.edge_attr_bond_embed.bias, self.projection_a.weight, self.projection_a.bias,
                                self.a_b, self.a)}
        out = []
        for p in self._live_parameters():
            if id(p) in live:
                out.append(on(p))
            elif p is self.edge_attr_fbond_embed.weight and p.shape[1] != 6:
                out.append(torch.zeros(p.shape[0], 6, device=dev))      # any width: the empty graph never reads it
            else:
                out.append(on(p.detach()))
        return out

    def forward(self, x_atoms, edge_index, edge_attr, x_frags, atom_to_frag_ids, node_feautures_bond_graph,
                edge_index_bonds_graph, edge_attr_bond_graph):
        self._check_geometry()
        home = x_atoms.device
        dev = ops.require_cuda(home if home.type == "cuda" else self.a.device)
        on = lambda t: t if t.device == dev else t.to(dev)
        x_fb, e_idx, e_attr = _empty_side(dev, self.projection_fb.in_features)
        plan = _g._plan_for(dev, x_atoms.size(0), x_frags.size(0), node_feautures_bond_graph.size(0), 0, edge_index,
                            e_idx, atom_to_frag_ids, edge_index_bonds_graph, edge_attr_bond_graph, e_idx, e_attr)
        sw = self._switches(False, self.return_attentions)
        cfg = EncoderConfig([sw], post_act=False, precision=config.precision_id(), grad_enabled=torch.is_grad_enabled())
        outs = EncoderFn.apply(plan, cfg, on(x_atoms), on(node_feautures_bond_graph), x_fb, *self._lite_parameters(dev))
        x_atoms_new, new_bond = outs[0], outs[2]
        pooled = PoolFn.apply(plan, x_atoms_new)                    # gat2_lite.py:140
        res = (x_atoms_new, pooled, new_bond, None)
        if self.return_attentions:
            res += (outs[4], None, outs[6], None)
        if home.type != "cuda":
            res = tuple(None if t is None else t.to(home) for t in res)
        return res


class FragNet(nn.Module):
    """gat2_lite.py:153-216 (note ``edge_features=16`` as the default there)."""

    def __init__(self, num_layer, drop_ratio=0.2, emb_dim=128, atom_features=167, frag_features=167, edge_features=16,
                 fedge_in=6, fbond_edge_in=6, num_heads=4):
        super().__init__()
        self.num_layer = num_layer
        self.dropout = nn.Dropout(p=drop_ratio)
        self.act = nn.ReLU()
        self.layers = nn.ModuleList()
        self.layers.append(FragNetLayerA(atom_in=atom_features, atom_out=emb_dim, frag_in=frag_features,
                                         frag_out=emb_dim, edge_in=edge_features, fedge_in=fedge_in,
                                         fbond_edge_in=fbond_edge_in, edge_out=emb_dim, num_heads=num_heads))
        for _ in range(num_layer - 1):
            self.layers.append(FragNetLayerA(atom_in=emb_dim, atom_out=emb_dim, frag_in=emb_dim, frag_out=emb_dim,
                                             edge_in=emb_dim, edge_out=emb_dim, fedge_in=emb_dim,
                                             fbond_edge_in=fbond_edge_in, num_heads=num_heads))

    def forward(self, batch):
        x_atoms = batch["x_atoms"]
        home = x_atoms.device
        dev = ops.require_cuda(home if home.type == "cuda" else self.layers[0].a.device)
        on = lambda t: t if t.device == dev else t.to(dev)
        for layer in self.layers:
            layer._check_geometry()
        bond_nodes = batch["node_features_bonds"]
        x_fb, e_idx, e_attr = _empty_side(dev, self.layers[0].projection_fb.in_features)
        plan = _g._plan_for(dev, x_atoms.size(0), batch["x_frags"].size(0), bond_nodes.size(0), 0, batch["edge_index"],
                            e_idx, batch["atom_to_frag_ids"], batch["edge_index_bonds_graph"], batch["edge_attr_bonds"],
                            e_idx, e_attr, batch.get("batch"), batch.get("frag_batch"))
        p, training, grad = float(self.dropout.p), self.training, torch.is_grad_enabled()
        xa, xb = on(x_atoms), on(bond_nodes)
        body, last = list(self.layers[:-1]), self.layers[-1]
        if body:      # every layer but the last: one program call, ReLU(Dropout(.)) fused (gat2_lite.py:193-212)
            cfg = EncoderConfig([l._switches(False, False) for l in body], post_act=True, drop_p=p, training=training,
                                precision=config.precision_id(), grad_enabled=grad)
            params = [t for l in body for t in l._lite_parameters(dev)]
            outs = EncoderFn.apply(plan, cfg, xa, xb, x_fb, *params)
            xa, xb = outs[0], outs[2]
            x_fb = torch.zeros(0, ops.D, device=dev)
        elif training and p > 0:
            xa = DropoutReluFn.apply(xa, p, True, False)             # input dropout (gat2_lite.py:188)
        # last layer: pre-activation outputs, pooled fragments, then ReLU(Dropout(.)) on all three
        cfg = EncoderConfig([last._switches(False, False)], post_act=False, precision=config.precision_id(),
                            grad_enabled=grad)
        outs = EncoderFn.apply(plan, cfg, xa, xb, x_fb, *last._lite_parameters(dev))
        pooled = PoolFn.apply(plan, outs[0])
        post = lambda t: DropoutReluFn.apply(t, p, training, True)
        result = (post(outs[0]), post(pooled), post(outs[2]), None)
        if home.type != "cuda":
            result = tuple(None if t is None else t.to(home) for t in result)
        return result


class FragNetFineTune(nn.Module):
    """gat2_lite.py:467-509: encoder, per-molecule sums of atoms and (pooled) fragments, regression head."""

    def __init__(self, n_classes=1, atom_features=167, frag_features=167, edge_features=16, num_layer=4, num_heads=4,
                 drop_ratio=0.15, h1=256, h2=256, h3=256, h4=256, act="celu", emb_dim=128, fthead="FTHead3"):
        super().__init__()
        self.pretrain = FragNet(num_layer=num_layer, drop_ratio=drop_ratio, num_heads=num_heads, emb_dim=emb_dim,
                                atom_features=atom_features, frag_features=frag_features, edge_features=edge_features)
        if fthead == "FTHead1":
            self.fthead = FTHead1(n_classes=n_classes)
        elif fthead == "FTHead2":
            self.fthead = FTHead2(n_classes=n_classes)
        elif fthead == "FTHead3":
            self.fthead = FTHead3(n_classes=n_classes, h1=h1, h2=h2, h3=h3, h4=h4, drop_ratio=drop_ratio, act=act)
        elif fthead == "FTHead4":
            self.fthead = FTHead4(n_classes=n_classes, h1=h1, drop_ratio=drop_ratio, act=act)

    def forward(self, batch):
        x_atoms, x_frags, _, _ = self.pretrain(batch)
        return self.fthead(graph_readout(x_atoms, x_frags, batch))
