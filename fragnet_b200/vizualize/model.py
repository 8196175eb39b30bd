"""Attention-returning task models: drop-ins for the classes of the reference's ``fragnet/vizualize/model.py``.

``FragNetViz`` (:45-142) is the GAT2 encoder whose LAST layer is built with ``return_attentions=True`` and whose
``forward`` returns the four embeddings followed by that layer's by-source attention sums
``(atoms, frags, bonds, fbonds)``; ``FragNetFineTuneViz`` (:146-201), ``FragNetFineTuneBaseViz`` (:205-248) and
``FragNetPreTrainViz`` (:256-280) put the readout + a head on top and hand the attention tensors through.  Parameter
names and registration order equal those of ``FragNetFineTune`` / ``FragNetPreTrain``, so a trained checkpoint loads
strictly (that is how viz.py:560-575 uses them).  The reference module imports RDKit / matplotlib at file scope and
cannot be imported where those are absent; this one has no such dependency.

One ``fnb_encoder_forward`` call produces embeddings and attention sums of the whole batch
(``FragNet.forward_with_attention``).
"""
from __future__ import annotations

import torch.nn as nn

from ..model.gat.gat2 import FragNet, FTHead1, FTHead2, FTHead3, FTHead4, graph_readout
from ..model.gat.pretrain_heads import PretrainTask


class FragNetViz(FragNet):
    """vizualize/model.py:45-142."""

    def __init__(self, num_layer, drop_ratio=0.2, emb_dim=128, atom_features=167, frag_features=167, edge_features=17,
                 fedge_in=6, fbond_edge_in=6, num_heads=4):
        super().__init__(num_layer, drop_ratio, emb_dim, atom_features, frag_features, edge_features, fedge_in,
                         fbond_edge_in, num_heads)
        self.layers[-1].return_attentions = True          # model.py:65-68

    def forward(self, batch):
        return self.forward_with_attention(batch)


def _head(fthead, n_classes, h1, h2, h3, h4, drop_ratio, act):
    if fthead == "FTHead1":
        return FTHead1(n_classes=n_classes)
    if fthead == "FTHead2":
        return FTHead2(n_classes=n_classes)
    if fthead == "FTHead3":
        return FTHead3(n_classes=n_classes, h1=h1, h2=h2, h3=h3, h4=h4, drop_ratio=drop_ratio, act=act)
    if fthead == "FTHead4":
        return FTHead4(n_classes=n_classes, h1=h1, drop_ratio=drop_ratio, act=act)
    return None


class FragNetFineTuneViz(nn.Module):
    """vizualize/model.py:146-201: ``forward`` returns ``(prediction, attn_atoms, attn_frags, attn_bonds, attn_fbonds)``."""

    def __init__(self, n_classes=1, atom_features=167, frag_features=167, edge_features=16, num_layer=4, num_heads=4,
                 drop_ratio=0.15, h1=256, h2=256, h3=256, h4=256, act="celu", emb_dim=128, fthead="FTHead3"):
        super().__init__()
        self.pretrain = FragNetViz(num_layer=num_layer, drop_ratio=drop_ratio, num_heads=num_heads, emb_dim=emb_dim,
                                   atom_features=atom_features, frag_features=frag_features,
                                   edge_features=edge_features)
        head = _head(fthead, n_classes, h1, h2, h3, h4, drop_ratio, act)
        if head is not None:
            self.fthead = head

    def forward(self, batch):
        x_atoms, x_frags, _, _, a_atoms, a_frags, a_bonds, a_fbonds = self.pretrain(batch)
        return self.fthead(graph_readout(x_atoms, x_frags, batch)), a_atoms, a_frags, a_bonds, a_fbonds


class FragNetFineTuneBaseViz(FragNetFineTuneViz):
    """vizualize/model.py:205-248: the graph representation ``cat(atoms pooled, frags pooled)`` itself."""

    def __init__(self, n_classes=1, atom_features=167, frag_features=167, edge_features=17, **kw):
        super().__init__(n_classes, atom_features, frag_features, edge_features, **kw)

    def forward(self, batch):
        x_atoms, x_frags = self.pretrain(batch)[:2]
        return graph_readout(x_atoms, x_frags, batch)


class FragNetPreTrainViz(nn.Module):
    """vizualize/model.py:256-280: ``forward`` returns ``(energy / graph_rep, attn_atoms, attn_frags, attn_bonds,
    attn_fbonds)``."""

    def __init__(self, num_layer=4, drop_ratio=0.15, num_heads=4, emb_dim=128, atom_features=167, frag_features=167,
                 edge_features=16):
        super().__init__()
        self.pretrain = FragNetViz(num_layer=num_layer, drop_ratio=drop_ratio, num_heads=num_heads, emb_dim=emb_dim,
                                   atom_features=atom_features, frag_features=frag_features,
                                   edge_features=edge_features)
        self.head = PretrainTask(128, 1)

    def forward(self, batch):
        x_atoms, x_frags, x_edge, _, a_atoms, a_frags, a_bonds, a_fbonds = self.pretrain(batch)
        graph_rep = self.head(x_atoms, x_frags, x_edge, batch)[3]
        return graph_rep, a_atoms, a_frags, a_bonds, a_fbonds
