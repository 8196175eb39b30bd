"""Drop-in ``FragNetLayerA`` / ``FragNet`` / ``FragNetFineTune`` / ``FTHead*`` backed by sm_100a kernels.

Mirrors the public surface of the reference's ``fragnet/model/gat/gat2.py`` for the GAT2 hot path:
same class names, constructor keywords, ``forward`` signatures (including the reference's spelling
``node_feautures_*``), returned tuple arity and order, parameter names / shapes / registration order
(so ``state_dict`` round-trips strictly in both directions), and the externally poked attributes
``bond_mask`` / ``frag_bond_mask`` / ``atom_mask_individual`` / ``return_attentions``.

What differs, deliberately:
* the arithmetic runs in hand-written CUDA kernels through ``libfragnet_b200.so`` (no torch_scatter,
  no torch_geometric, no CPU fallback);  CPU inputs are staged to the current CUDA device and the
  results are handed back on the inputs' device;
* the layer does not ``print`` on every forward (reference gat2.py:172,174,230,232);
* the never-used ``bias`` parameter is zero-filled instead of uninitialised memory (gat2.py:81);
* inside ``FragNet`` the fragment-graph block of every layer but the last is skipped: its output is
  overwritten unread by the next layer (gat2.py:234), so nothing observable changes.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import config, ops
from ...autograd import EncoderConfig, EncoderFn, LayerSwitches, ReadoutFn

_ACTIVATIONS = {
    "relu": nn.ReLU, "silu": nn.SiLU, "gelu": nn.GELU, "celu": nn.CELU, "selu": nn.SELU,
    "rrelu": nn.RReLU, "relu6": nn.ReLU6, "prelu": nn.PReLU, "leakyrelu": nn.LeakyReLU,
}


def _as_int_or_none(v):
    if v is None or isinstance(v, int):
        return v
    if hasattr(v, "ndim") and v.ndim == 0:
        return int(v)
    return v


def _plan_for(dev, n_atoms, n_frags, n_bond_nodes, n_fbond_nodes, edge_index, frag_index, atom_to_frag_ids,
              edge_index_bonds_graph, edge_attr_bonds, edge_index_fbonds, edge_attr_fbonds, batch_vec=None,
              frag_batch_vec=None):
    """CSR plan of a batch (built on the device once, cached on the identity of the caller's index tensors)."""
    index_tensors = (edge_index, frag_index, atom_to_frag_ids, edge_index_bonds_graph, edge_attr_bonds,
                     edge_index_fbonds, edge_attr_fbonds)
    return ops.layer_plan_for(index_tensors, (n_atoms, n_frags, n_bond_nodes, n_fbond_nodes), dev, batch_vec,
                              frag_batch_vec)


class FragNetLayerA(nn.Module):
    """One GAT2 layer over the four coupled graphs (reference gat2.py:40-330)."""

    def __init__(self, atom_in=128, atom_out=128, frag_in=128, frag_out=128, edge_in=128, edge_out=128,
                 fedge_in=128, num_heads=2, bond_edge_in=1, fbond_edge_in=8, return_attentions=False,
                 add_frag_self_loops=False, bond_mask=None, frag_bond_mask=None, atom_mask_individual=None):
        super().__init__()
        self.add_frag_self_loops = add_frag_self_loops
        self.return_attentions = return_attentions
        self.edge_out = edge_out
        self.num_heads = num_heads
        # Registration order below is the reference's (gat2.py:64-109): it fixes state_dict order and
        # the order in which the global RNG is consumed at construction.
        # -- parameters the reference constructs and checkpoints but never reads in forward
        self.atom_embed = nn.Linear(atom_in, atom_out)
        self.frag_embed = nn.Linear(frag_in, frag_out)
        self.edge_embed = nn.Linear(edge_in, edge_out)
        self.bond_edge_embed = nn.Linear(edge_in, edge_out)
        self.frag_message_mlp = nn.Linear(2 * atom_out, atom_out)
        for name in ("atom_mlp", "frag_mlp"):
            setattr(self, name, nn.Sequential(nn.Linear(atom_out, 2 * atom_out), nn.ReLU(),
                                              nn.Linear(2 * atom_out, atom_out)))
        self.bias = nn.Parameter(torch.zeros(atom_out))
        self.leakyrelu = nn.LeakyReLU(0.2)
        self.edge_attr_bond_embed2 = nn.Linear(edge_out, edge_out)
        # -- live parameters
        d_edge, d_atom = edge_out // num_heads, atom_out // num_heads
        self.projection_b = nn.Linear(edge_in, d_edge * num_heads)
        self.projection_fb = nn.Linear(fedge_in, d_edge * num_heads)
        self.edge_attr_bond_embed = nn.Linear(bond_edge_in, d_edge)
        self.edge_attr_fbond_embed = nn.Linear(fbond_edge_in, d_edge)
        self.projection_a = nn.Linear(atom_in, d_atom * num_heads)
        self.a_b = nn.Parameter(torch.empty(num_heads, 3 * d_edge))
        self.a = nn.Parameter(torch.empty(num_heads, 2 * d_atom + d_edge * num_heads))
        self.f = nn.Parameter(torch.empty(num_heads, 2 * d_atom + d_edge * num_heads))
        self.f_a_b = nn.Parameter(torch.empty(num_heads, 3 * d_edge))
        for t in (self.projection_b.weight, self.a_b, self.a, self.f, self.f_a_b):   # gat2.py:111-115
            nn.init.xavier_uniform_(t.data, gain=1.414)
        self.bond_mask = bond_mask
        self.frag_bond_mask = frag_bond_mask
        self.atom_mask_individual = atom_mask_individual
        self._geometry = (atom_out, edge_out, num_heads, bond_edge_in, fbond_edge_in)

    def _check_geometry(self):
        atom_out, edge_out, heads, bond_edge_in, fbond_edge_in = self._geometry
        if (atom_out, edge_out, heads, bond_edge_in, fbond_edge_in) != (ops.D, ops.D, ops.H, 1, 6):
            raise NotImplementedError(
                "fragnet_b200 kernels are specialised for emb_dim=128, num_heads=4, bond_edge_in=1, "
                f"fbond_edge_in=6 (every shipped gat2 config); got {self._geometry}")

    def _live_parameters(self):
        return (self.projection_b.weight, self.projection_b.bias, self.projection_fb.weight, self.projection_fb.bias,
                self.edge_attr_bond_embed.weight, self.edge_attr_bond_embed.bias,
                self.edge_attr_fbond_embed.weight, self.edge_attr_fbond_embed.bias,
                self.projection_a.weight, self.projection_a.bias, self.a_b, self.a, self.f, self.f_a_b)

    def _switches(self, run_frag_block=True, want_attention=False) -> LayerSwitches:
        return LayerSwitches(run_frag_block, want_attention, _as_int_or_none(self.bond_mask),
                             _as_int_or_none(self.frag_bond_mask), _as_int_or_none(self.atom_mask_individual))

    def _run(self, x_atoms, edge_index, frag_index, n_frags, atom_to_frag_ids, x_bond_nodes, edge_index_bonds_graph,
             edge_attr_bond_graph, x_fbond_nodes, edge_index_fbond_graph, edge_attr_fbond_graph, want_attention=False):
        """The bare layer: one ``fnb_encoder_forward`` call in pre-activation mode."""
        self._check_geometry()
        dev = ops.require_cuda(x_atoms.device if x_atoms.is_cuda else self.a.device)
        plan = _plan_for(dev, x_atoms.size(0), int(n_frags), x_bond_nodes.size(0), x_fbond_nodes.size(0), edge_index,
                         frag_index, atom_to_frag_ids, edge_index_bonds_graph, edge_attr_bond_graph,
                         edge_index_fbond_graph, edge_attr_fbond_graph)
        cfg = EncoderConfig([self._switches(True, want_attention)], post_act=False, precision=config.precision_id(),
                            grad_enabled=torch.is_grad_enabled())
        on = lambda t: t if t.device == dev else t.to(dev)
        params = [on(p) for p in self._live_parameters()]
        return EncoderFn.apply(plan, cfg, on(x_atoms), on(x_bond_nodes), on(x_fbond_nodes), *params)

    def forward(self, x_atoms, edge_index, edge_attr, frag_index, x_frags, atom_to_frag_ids,
                node_feautures_bond_graph, edge_index_bonds_graph, edge_attr_bond_graph,
                node_feautures_fbond_graph, edge_index_fbond_graph, edge_attr_fbond_graph):
        """Same positional signature and return value as the reference (gat2.py:121-135, 318-330).
        ``edge_attr`` and the values of ``x_frags`` are ignored there too (only ``x_frags.size(0)``
        matters: gat2.py:184, :234)."""
        home = x_atoms.device
        outs = self._run(x_atoms, edge_index, frag_index, x_frags.size(0), atom_to_frag_ids,
                         node_feautures_bond_graph, edge_index_bonds_graph, edge_attr_bond_graph,
                         node_feautures_fbond_graph, edge_index_fbond_graph, edge_attr_fbond_graph,
                         want_attention=self.return_attentions)
        if home.type != "cuda":
            outs = tuple(o.to(home) for o in outs)
        return outs


class FragNet(nn.Module):
    """The GAT2 encoder (reference gat2.py:333-442)."""

    def __init__(self, num_layer, drop_ratio=0.2, emb_dim=128, atom_features=167, frag_features=167,
                 edge_features=17, fedge_in=6, fbond_edge_in=6, num_heads=4):
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

    def _encode(self, batch, attention_from_last: bool):
        """All layers in one ``fnb_encoder_forward`` call (and one ``fnb_encoder_backward`` call under autograd)."""
        x_atoms = batch["x_atoms"]
        home = x_atoms.device
        dev = ops.require_cuda(home if home.type == "cuda" else self.layers[0].a.device)
        for layer in self.layers:
            layer._check_geometry()
        bond_nodes, fbond_nodes = batch["node_features_bonds"], batch["node_features_fbonds"]
        plan = _plan_for(dev, x_atoms.size(0), batch["x_frags"].size(0), bond_nodes.size(0), fbond_nodes.size(0),
                         batch["edge_index"], batch["frag_index"], batch["atom_to_frag_ids"],
                         batch["edge_index_bonds_graph"], batch["edge_attr_bonds"], batch["edge_index_fbonds"],
                         batch["edge_attr_fbonds"], batch.get("batch"), batch.get("frag_batch"))
        last = len(self.layers) - 1
        # the fragment-graph block of every layer but the last is skipped: its output is overwritten unread by the
        # next layer (gat2.py:234), so nothing observable changes
        switches = [layer._switches(run_frag_block=li == last,
                                    want_attention=(attention_from_last and li == last) or
                                                   (layer.return_attentions and li == last))
                    for li, layer in enumerate(self.layers)]
        cfg = EncoderConfig(switches, post_act=True, drop_p=float(self.dropout.p), training=self.training,
                            precision=config.precision_id(), grad_enabled=torch.is_grad_enabled())
        on = lambda t: t if t.device == dev else t.to(dev)
        params = [on(p) for layer in self.layers for p in layer._live_parameters()]
        outs = EncoderFn.apply(plan, cfg, on(x_atoms), on(bond_nodes), on(fbond_nodes), *params)
        result = tuple(outs[:4]) + (tuple(outs[4:]) if attention_from_last else ())
        if home.type != "cuda":
            result = tuple(t.to(home) for t in result)
        return result

    def forward(self, batch):
        return self._encode(batch, attention_from_last=False)

    def forward_with_attention(self, batch):
        """Encoder outputs plus the last layer's (atoms, frags, bonds, fbonds) attention sums: the
        arrangement of the reference's ``vizualize/model.py:72-142`` in one call."""
        return self._encode(batch, attention_from_last=True)


def graph_readout(x_atoms, x_frags, batch):
    """``cat(scatter_add(x_atoms, batch), scatter_add(x_frags, frag_batch))`` (gat2.py:820-823)."""
    home = x_atoms.device
    dev = ops.require_cuda(home)
    rp = ops.readout_plan_for(batch["batch"], batch["frag_batch"], dev)
    out = ReadoutFn.apply(rp, x_atoms.to(dev), x_frags.to(dev))
    return out if home.type == "cuda" else out.to(home)


class _MLPHead(nn.Sequential):
    """``act(dropout(linear(x)))`` per hidden layer, bare last linear (gat2.py:631-637, 719-725)."""

    def _build(self, dims, drop_ratio, act):
        self.dropout = nn.Dropout(p=drop_ratio)
        if act in _ACTIVATIONS:
            self.activation = _ACTIVATIONS[act]()
        self.predictor = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])

    def forward(self, enc):
        for lin in self.predictor[:-1]:
            enc = self.activation(self.dropout(lin(enc)))
        return self.predictor[-1](enc)


class FTHead3(_MLPHead):
    def __init__(self, input_dim=128, h1=128, h2=1024, h3=1024, h4=512, drop_ratio=0.2, n_classes=1, act="relu"):
        super().__init__()
        self.hidden_dims = [h1, h2, h3, h4]
        self._build([input_dim * 2] + self.hidden_dims + [n_classes], drop_ratio, act)


class FTHead5(_MLPHead):
    def __init__(self, input_dim=128, h1=128, h2=1024, h4=512, drop_ratio=0.2, n_classes=1, act="relu"):
        super().__init__()
        self.hidden_dims = [h1, h2]
        self._build([input_dim * 2] + self.hidden_dims + [n_classes], drop_ratio, act)


class FTHead2(nn.Sequential):
    """gat2.py:728-751: registers an unused ``lin1`` / ``out`` pair, then a fixed 1024-1024-512 MLP with p=0.1."""

    def __init__(self, input_dim=128, h1=128, drop_ratio=0.2, n_classes=1):
        super().__init__()
        self.lin1 = nn.Linear(input_dim * 2, h1)
        self.out = nn.Linear(h1, n_classes)
        self.dropout = nn.Dropout(p=drop_ratio)
        self.activation = nn.ReLU()
        self.hidden_dims = [1024, 1024, 512]
        dims = [input_dim * 2] + self.hidden_dims + [n_classes]
        self.predictor = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.dropout = nn.Dropout(p=0.1)

    def forward(self, enc):
        for lin in self.predictor[:-1]:
            enc = torch.relu(self.dropout(lin(enc)))
        return self.predictor[-1](enc)


class FTHead1(nn.Sequential):
    """gat2.py:569-588: dropout, linear, ReLU, dropout, linear."""

    def __init__(self, emb_dim=128, h1=128, drop_ratio=0.2, n_classes=1):
        super().__init__()
        self.lin1 = nn.Linear(emb_dim * 2, h1)
        self.out = nn.Linear(h1, n_classes)
        self.dropout = nn.Dropout(p=drop_ratio)
        self.activation = nn.ReLU()

    def forward(self, enc):
        return self.out(self.dropout(self.activation(self.lin1(self.dropout(enc)))))


class FTHead4(nn.Module):
    """gat2.py:640-675: dropout, dense, act, dropout, out_proj."""

    def __init__(self, input_dim=128, h1=128, act="relu", n_classes=1, drop_ratio=0.2):
        super().__init__()
        if act in _ACTIVATIONS:
            self.activation = _ACTIVATIONS[act]()
        self.dense = nn.Linear(input_dim * 2, h1)
        self.dropout = nn.Dropout(p=drop_ratio)
        self.out_proj = nn.Linear(h1, n_classes)

    def forward(self, x):
        return self.out_proj(self.dropout(self.activation(self.dense(self.dropout(x)))))


def do_nothing(a):
    pass


class FragNetFineTune(nn.Module):
    """Encoder + readout + regression head (reference gat2.py:758-826)."""

    def __init__(self, n_classes=1, atom_features=167, frag_features=167, edge_features=17, num_layer=4,
                 num_heads=4, drop_ratio=0.15, h1=256, h2=256, h3=256, h4=256, act="celu", emb_dim=128,
                 fthead="FTHead3"):
        super().__init__()
        self.pretrain = FragNet(num_layer=num_layer, drop_ratio=drop_ratio, num_heads=num_heads, emb_dim=emb_dim,
                                atom_features=atom_features, frag_features=frag_features,
                                edge_features=edge_features)
        if fthead == "FTHead1":
            self.fthead = FTHead1(n_classes=n_classes)
        elif fthead == "FTHead2":
            self.fthead = FTHead2(n_classes=n_classes)
        elif fthead == "FTHead3":
            self.fthead = FTHead3(n_classes=n_classes, input_dim=emb_dim, h1=h1, h2=h2, h3=h3, h4=h4,
                                  drop_ratio=drop_ratio, act=act)
        elif fthead == "FTHead4":
            self.fthead = FTHead4(n_classes=n_classes, h1=h1, drop_ratio=drop_ratio, act=act)

    def forward(self, batch):
        x_atoms, x_frags, _, _ = self.pretrain(batch)
        return self.fthead(graph_readout(x_atoms, x_frags, batch))
