"""Drop-in ``gat2_edge``: the three-graph variant (bond graph, atom graph, pooling, fragment graph whose edge term comes
from the connection attributes).

Mirrors the reference's ``fragnet/model/gat/gat2_edge.py`` -- ``FragNetLayerA`` (:13-176: 10-argument ``forward``, 3- or
6-tuple return), ``FragNet`` (:180-239), ``FragNetFineTune`` (:520-561) -- selected by ``model_version: gat2_edge``
(train/finetune/finetune_gat2.py:165-166).  Parameter names, shapes and registration order are the reference's
(``cnx_attr_transform`` = ``Linear(8, 128)``, gat2_edge.py:46; no ``projection_fb`` / ``f_a_b`` / fragment-connection
embedding).  The reference's current featuriser emits 6-wide connection attributes; this variant needs the 8-wide ones
of the data version it was written for, exactly like upstream.

The arithmetic is the same sm_100a encoder program as ``gat2``.  The fragment block reads
``<cnx_attr_transform(cnx_attr[e]), f_e[h]>`` per edge and head (gat2_edge.py:152-158: every head sees the full 128-wide
transformed attribute); by App. A.5 of the survey that is the affine map ``cnx_attr[e] @ (W8^T f_e[h]) + <b8, f_e[h]>`` of
the 8 raw attributes -- an ``[Ef, 8] x [8, 4]`` product formed here with two tiny torch ops (autograd carries its
gradient to ``cnx_attr_transform`` and ``f``) and handed to the program as the fragment graph's edge table
(``fnb_encoder_io.frag_table``); the program returns the table's gradient.  The fragment-connection graph of ``gat2``
does not exist in this variant: its chain runs over zero-filled operands with an empty edge list.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import config, ops
from ...autograd import EncoderConfig, EncoderFn, LayerSwitches
from . import gat2 as _g
from .gat2 import FTHead1, FTHead2, FTHead3, FTHead4, FTHead5, graph_readout  # noqa: F401  (same classes as gat2)

CNX_IN = 8     # gat2_edge.py:46


class FragNetLayerA(nn.Module):
    """gat2_edge.py:13-176."""

    def __init__(self, atom_in=128, atom_out=128, frag_in=128, frag_out=128, edge_in=128, edge_out=128, num_heads=2,
                 bond_edge_in=1, return_attentions=False, add_frag_self_loops=False):
        super().__init__()
        self.add_frag_self_loops = add_frag_self_loops
        self.return_attentions = return_attentions
        self.edge_out = edge_out
        # registration order = the reference's (gat2_edge.py:22-53): it fixes state_dict order and RNG consumption
        self.atom_embed = nn.Linear(atom_in, atom_out, bias=True)
        self.frag_embed = nn.Linear(frag_in, frag_out)
        self.edge_embed = nn.Linear(edge_in, edge_out)
        self.bond_edge_embed = nn.Linear(edge_in, edge_out)
        self.frag_message_mlp = nn.Linear(atom_out * 2, atom_out)
        for name in ("atom_mlp", "frag_mlp"):
            setattr(self, name, nn.Sequential(nn.Linear(atom_out, 2 * atom_out), nn.ReLU(),
                                              nn.Linear(2 * atom_out, atom_out)))
        self.bias = nn.Parameter(torch.zeros(atom_out))       # uninitialised memory upstream (gat2_edge.py:35), never read
        self.leakyrelu = nn.LeakyReLU(0.2)
        self.num_heads = num_heads
        self.edge_attr_bond_embed2 = nn.Linear(edge_out, edge_out)
        d_edge = edge_out // num_heads
        self.projection_b = nn.Linear(edge_in, d_edge * num_heads, bias=True)
        self.edge_attr_bond_embed = nn.Linear(bond_edge_in, d_edge)
        self.cnx_attr_transform = nn.Linear(CNX_IN, edge_out)
        d_atom = atom_out // num_heads
        self.projection_a = nn.Linear(atom_in, d_atom * num_heads)
        self.a_b = nn.Parameter(torch.empty(num_heads, 2 * d_edge + d_edge))
        self.a = nn.Parameter(torch.empty(num_heads, 2 * d_atom + d_edge * num_heads))
        self.f = nn.Parameter(torch.empty(num_heads, 2 * d_atom + d_edge * num_heads))
        for t in (self.projection_b.weight, self.a_b, self.a, self.f):          # gat2_edge.py:55-59
            nn.init.xavier_uniform_(t.data, gain=1.414)
        self._geometry = (atom_out, edge_out, num_heads, bond_edge_in)

    def _check_geometry(self):
        if self._geometry != (ops.D, ops.D, ops.H, 1):
            raise NotImplementedError("fragnet_b200 kernels are specialised for emb_dim=128, num_heads=4, "
                                      f"bond_edge_in=1; got {self._geometry}")
        if self.add_frag_self_loops:
            raise NotImplementedError("fragnet_b200: add_frag_self_loops=True is not supported (no shipped caller sets "
                                      "it: gat2_edge.py:190-196)")

    def _program_parameters(self, dev, k_fbond=6):
        """The 14 tensors of the encoder program in ``fnb_layer_params`` order; the fragment-connection side of ``gat2``
        does not exist here and is zero-filled (``k_fbond``: width of that dummy chain's input, 6 for a first layer and
        128 behind another layer)."""
        on = lambda t: t if t.device == dev else t.to(dev)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        return [on(self.projection_b.weight), on(self.projection_b.bias), z(ops.D, k_fbond), z(ops.D),
                on(self.edge_attr_bond_embed.weight), on(self.edge_attr_bond_embed.bias), z(32, 6), z(32),
                on(self.projection_a.weight), on(self.projection_a.bias), on(self.a_b), on(self.a), on(self.f), z(ops.H, 96)]

    def _frag_edge_table(self, cnx_attr, dev):
        """``S_e[e, h] = <cnx_attr_transform(cnx_attr[e]), f[h, 32:160]>`` (gat2_edge.py:152-158) as the affine map of
        the raw attributes: ``cnx_attr @ (W8^T f_e^T) + b8 f_e^T``  ->  [Ef, 4]."""
        f_e = self.f[:, 32:32 + ops.D]                                        # [4, 128]: the edge slice of every head
        coef = self.cnx_attr_transform.weight.t() @ f_e.t()                   # [8, 4]
        bias = self.cnx_attr_transform.bias @ f_e.t()                         # [4]
        cnx = cnx_attr.to(device=dev, dtype=torch.float32)
        return cnx @ coef.to(dev) + bias.to(dev)

    def forward(self, x_atoms, edge_index, edge_attr, frag_index, x_frags, atom_to_frag_ids, node_feautures_bond_graph,
                edge_index_bonds_graph, edge_attr_bond_graph, cnx_attr):
        """Same positional signature and return value as the reference (gat2_edge.py:62-71, 172-175)."""
        self._check_geometry()
        home = x_atoms.device
        dev = ops.require_cuda(home if home.type == "cuda" else self.a.device)
        on = lambda t: t if t.device == dev else t.to(dev)
        plan, x_fb = _edge_plan(dev, x_atoms.size(0), x_frags.size(0), node_feautures_bond_graph.size(0), edge_index,
                                frag_index, atom_to_frag_ids, edge_index_bonds_graph, edge_attr_bond_graph)
        sw = LayerSwitches(True, self.return_attentions)
        cfg = EncoderConfig([sw], post_act=False, precision=config.precision_id(), grad_enabled=torch.is_grad_enabled(),
                            frag_table=True)
        outs = EncoderFn.apply(plan, cfg, on(x_atoms), on(node_feautures_bond_graph), x_fb,
                               *self._program_parameters(dev), self._frag_edge_table(cnx_attr, dev))
        res = (outs[0], outs[1], outs[2])
        if self.return_attentions:
            res += (outs[4], outs[5], outs[6])
        if home.type != "cuda":
            res = tuple(t.to(home) for t in res)
        return res


_EMPTY = {}


def _edge_plan(dev, n_atoms, n_frags, n_bond_nodes, edge_index, frag_index, atom_to_frag_ids, edge_index_bonds_graph,
               edge_attr_bonds, batch_vec=None, frag_batch_vec=None):
    """Batch plan of the three-graph variant: the fragment-connection graph keeps its node set (the program ties the
    fragment graph's edges to it) but has no edges; its features are zeros."""
    key = str(dev)
    if key not in _EMPTY:      # persistent empties: the plan cache is keyed on tensor identity
        _EMPTY[key] = (torch.zeros((2, 0), dtype=torch.long, device=dev), torch.zeros((0, 6), device=dev))
    e_idx, e_attr = _EMPTY[key]
    n_fb = frag_index.shape[1]
    plan = _g._plan_for(dev, n_atoms, n_frags, n_bond_nodes, n_fb, edge_index, frag_index, atom_to_frag_ids,
                        edge_index_bonds_graph, edge_attr_bonds, e_idx, e_attr, batch_vec, frag_batch_vec)
    return plan, torch.zeros((n_fb, 6), device=dev)


class FragNet(nn.Module):
    """gat2_edge.py:180-239 (note ``edge_features=16`` as the default there)."""

    def __init__(self, num_layer, drop_ratio=0.2, emb_dim=128, atom_features=167, frag_features=167, edge_features=16,
                 num_heads=4):
        super().__init__()
        self.num_layer = num_layer
        self.dropout = nn.Dropout(p=drop_ratio)
        self.act = nn.ReLU()
        self.layers = nn.ModuleList()
        self.layers.append(FragNetLayerA(atom_in=atom_features, atom_out=emb_dim, frag_in=frag_features,
                                         frag_out=emb_dim, edge_in=edge_features, edge_out=emb_dim, num_heads=num_heads))
        for _ in range(num_layer - 1):
            self.layers.append(FragNetLayerA(atom_in=emb_dim, atom_out=emb_dim, frag_in=emb_dim, frag_out=emb_dim,
                                             edge_in=emb_dim, edge_out=emb_dim, num_heads=num_heads))

    def forward(self, batch):
        """All layers in one ``fnb_encoder_forward`` call; returns ``(x_atoms, x_frags, edge_features)``."""
        x_atoms = batch["x_atoms"]
        home = x_atoms.device
        dev = ops.require_cuda(home if home.type == "cuda" else self.layers[0].a.device)
        on = lambda t: t if t.device == dev else t.to(dev)
        for layer in self.layers:
            layer._check_geometry()
        bond_nodes = batch["node_features_bonds"]
        plan, x_fb = _edge_plan(dev, x_atoms.size(0), batch["x_frags"].size(0), bond_nodes.size(0), batch["edge_index"],
                                batch["frag_index"], batch["atom_to_frag_ids"], batch["edge_index_bonds_graph"],
                                batch["edge_attr_bonds"], batch.get("batch"), batch.get("frag_batch"))
        last = len(self.layers) - 1
        # as in gat2, the fragment block of every layer but the last is overwritten unread (gat2_edge.py:129): skipped
        switches = [LayerSwitches(li == last, False) for li in range(len(self.layers))]
        cfg = EncoderConfig(switches, post_act=True, drop_p=float(self.dropout.p), training=self.training,
                            precision=config.precision_id(), grad_enabled=torch.is_grad_enabled(), frag_table=True)
        params = [t for li, layer in enumerate(self.layers) for t in layer._program_parameters(dev, 6 if li == 0 else ops.D)]
        table = self.layers[last]._frag_edge_table(batch["cnx_attr"], dev)
        outs = EncoderFn.apply(plan, cfg, on(x_atoms), on(bond_nodes), x_fb, *params, table)
        result = (outs[0], outs[1], outs[2])
        if home.type != "cuda":
            result = tuple(t.to(home) for t in result)
        return result


class FragNetFineTune(nn.Module):
    """gat2_edge.py:520-561: encoder, per-molecule sums of atoms and fragments, regression head."""

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
        x_atoms, x_frags, _ = self.pretrain(batch)
        return self.fthead(graph_readout(x_atoms, x_frags, batch))
