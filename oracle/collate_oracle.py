"""Restatement of the reference's batch assembly.  TEST INFRASTRUCTURE ONLY.

Follows ``fragnet/dataset/data.py``: ``get_incr_*`` (:11-113) build, for each of the five index
tensors, a float32 vector holding -- for every column of molecule ``i`` -- the number of nodes in
molecules ``0..i-1``; ``collate_fn`` (:877-948) / ``collate_fn_pt`` (:951-1032) concatenate the
per-molecule tensors, add those vectors and cast back to int64.  The float32 detour is kept on
purpose (it is what the reference does; exact below 2**24 nodes per index space), except for the
fragment-connection graph whose offsets the reference casts to int64 before adding (:922-924).
"""
from __future__ import annotations

import torch


def _running_offsets(data_list, width_of, count_of) -> torch.Tensor:
    """float32 [sum widths]: molecule i's columns carry sum_{k<i} count_of(k)  (data.py:11-113)."""
    pieces, total = [], 0
    for i, d in enumerate(data_list):
        if i > 0:
            total += count_of(data_list[i - 1])
        pieces.append(torch.zeros(width_of(d)) + total)
    return torch.cat(pieces)


def collate(data_list, pretrain: bool = False):
    cat = torch.cat
    edge_index = cat([d.edge_index for d in data_list], dim=1) + _running_offsets(
        data_list, lambda d: d.edge_index.shape[1], lambda d: d.x_atoms.size(0))
    frag_index = cat([d.frag_index for d in data_list], dim=1) + _running_offsets(
        data_list, lambda d: d.frag_index.shape[1], lambda d: d.n_frags.item())
    a2f = cat([d.atom_id_frag_id for d in data_list], dim=0) + _running_offsets(
        data_list, lambda d: d.atom_id_frag_id.shape[0], lambda d: d.n_frags.item())
    ei_bonds = cat([d.edge_index_bonds for d in data_list], dim=1) + _running_offsets(
        data_list, lambda d: d.edge_index_bonds.shape[1], lambda d: d.node_features_bonds.size(0))
    ei_fbonds = cat([d.edge_index_fbondg for d in data_list], dim=1) + _running_offsets(
        data_list, lambda d: d.edge_index_fbondg.shape[1], lambda d: d.node_feautures_fbondg.size(0)).to(torch.long)
    batch = cat([torch.zeros(d.x_atoms.shape[0]) + i for i, d in enumerate(data_list)])
    frag_batch = cat([torch.zeros(d.n_frags.item()) + i for i, d in enumerate(data_list)])
    out = {
        "x_atoms": cat([d.x_atoms for d in data_list], dim=0),
        "edge_index": edge_index.type(torch.long),
        "frag_index": frag_index.type(torch.long),
        "x_frags": cat([d.x_frags for d in data_list], dim=0),
        "edge_attr": cat([d.edge_attr for d in data_list], dim=0),
        "cnx_attr": cat([d.cnx_attr for d in data_list], dim=0),
        "batch": batch.type(torch.long),
        "frag_batch": frag_batch.type(torch.long),
        "atom_to_frag_ids": a2f.type(torch.long),
        "node_features_bonds": cat([d.node_features_bonds for d in data_list], dim=0),
        "edge_index_bonds_graph": ei_bonds.type(torch.long),
        "edge_attr_bonds": cat([d.edge_attr_bonds for d in data_list], dim=0),
        "node_features_fbonds": cat([d.node_feautures_fbondg for d in data_list], dim=0),
        "edge_index_fbonds": ei_fbonds,
        "edge_attr_fbonds": cat([d.edge_attr_fbondg for d in data_list], dim=0),
    }
    if pretrain:
        out["bnd_lngth"] = cat([d.bnd_lngth for d in data_list], dim=0)
        out["bnd_angl"] = cat([d.bnd_angl for d in data_list], dim=0)
        out["dh_angl"] = cat([d.dh_angl for d in data_list], dim=0)
    out["y"] = cat([d.y for d in data_list], dim=0).type(torch.float)
    return out
