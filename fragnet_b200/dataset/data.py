"""Batch assembly for the GAT2 hot path: ``collate_fn`` / ``collate_fn_pt``.

Drop-in for the reference's ``fragnet/dataset/data.py:877-948`` (``collate_fn``)
and ``:951-1032`` (``collate_fn_pt``): same input (a list of per-molecule
records carrying the attribute names of ``data.py:437-482``), same output keys,
dtypes and values.  The reference concatenates per-molecule tensors and adds
per-molecule node-count prefixes to the five index tensors with Python loops
over the molecule list (``get_incr_*``, ``data.py:11-113``), routing the offsets
through float32 (exact below 2**24 nodes per index space).  Here the offsets are
one exclusive prefix sum per index space and one ``repeat_interleave`` per index
tensor, kept in int64 throughout, so the result is the same integers with no
2**24 ceiling.

RDKit-dependent graph *construction* (``CreateData``) is preprocessing and out
of scope for this package.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


def _prefix(counts: torch.Tensor) -> torch.Tensor:
    """Exclusive prefix sum of per-molecule counts (int64)."""
    out = torch.zeros_like(counts)
    if counts.numel() > 1:
        torch.cumsum(counts[:-1], 0, out=out[1:])
    return out


def _offset_columns(parts: List[torch.Tensor], node_counts: torch.Tensor, dim: int) -> torch.Tensor:
    """Concatenate per-molecule index tensors along ``dim`` and shift molecule ``i``'s entries by
    the number of nodes in molecules ``0..i-1`` (what ``get_incr_*`` + add does, data.py:11-113)."""
    cat = torch.cat([p.to(torch.long) for p in parts], dim=dim)
    widths = torch.tensor([p.shape[dim] for p in parts], dtype=torch.long)
    shift = torch.repeat_interleave(_prefix(node_counts), widths)
    return cat + shift


def _collate(data_list, pretrain: bool) -> Dict[str, torch.Tensor]:
    n_atoms = torch.tensor([d.x_atoms.shape[0] for d in data_list], dtype=torch.long)
    n_frags = torch.tensor([int(d.n_frags.item()) for d in data_list], dtype=torch.long)
    n_bnodes = torch.tensor([d.node_features_bonds.shape[0] for d in data_list], dtype=torch.long)
    n_fbnodes = torch.tensor([d.node_feautures_fbondg.shape[0] for d in data_list], dtype=torch.long)
    mol_ids = torch.arange(len(data_list), dtype=torch.long)
    out = {
        "x_atoms": torch.cat([d.x_atoms for d in data_list], dim=0),
        "edge_index": _offset_columns([d.edge_index for d in data_list], n_atoms, 1),
        "frag_index": _offset_columns([d.frag_index for d in data_list], n_frags, 1),
        "x_frags": torch.cat([d.x_frags for d in data_list], dim=0),
        "edge_attr": torch.cat([d.edge_attr for d in data_list], dim=0),
        "cnx_attr": torch.cat([d.cnx_attr for d in data_list], dim=0),
        "batch": torch.repeat_interleave(mol_ids, n_atoms),
        "frag_batch": torch.repeat_interleave(mol_ids, n_frags),
        "atom_to_frag_ids": _offset_columns([d.atom_id_frag_id for d in data_list], n_frags, 0),
        "node_features_bonds": torch.cat([d.node_features_bonds for d in data_list], dim=0),
        "edge_index_bonds_graph": _offset_columns([d.edge_index_bonds for d in data_list], n_bnodes, 1),
        "edge_attr_bonds": torch.cat([d.edge_attr_bonds for d in data_list], dim=0),
        "node_features_fbonds": torch.cat([d.node_feautures_fbondg for d in data_list], dim=0),
        "edge_index_fbonds": _offset_columns([d.edge_index_fbondg for d in data_list], n_fbnodes, 1),
        "edge_attr_fbonds": torch.cat([d.edge_attr_fbondg for d in data_list], dim=0),
    }
    if pretrain:
        out["bnd_lngth"] = torch.cat([d.bnd_lngth for d in data_list], dim=0)
        out["bnd_angl"] = torch.cat([d.bnd_angl for d in data_list], dim=0)
        out["dh_angl"] = torch.cat([d.dh_angl for d in data_list], dim=0)
    out["y"] = torch.cat([d.y for d in data_list], dim=0).type(torch.float)
    return out


def collate_fn(data_list):
    """Finetune batches: 16 keys (reference ``data.py:931-948``)."""
    return _collate(data_list, pretrain=False)


def collate_fn_pt(data_list):
    """Pretraining batches: adds ``bnd_lngth``, ``bnd_angl``, ``dh_angl`` (reference ``data.py:1012-1032``)."""
    return _collate(data_list, pretrain=True)


# ---- compact wire format ---------------------------------------------------------------------------------------------
# collate_fn's output is 32 MB per 1 024 UniMol-shaped molecules (39.6 MB with the tensors FragNet.forward never reads):
# one-hot / small-integer feature matrices as fp32 (x_atoms alone is 17 MB) and every index as int64.  The information
# is ~11 MB.  ``compact_batch`` narrows what can be narrowed EXACTLY -- feature matrices whose entries are integers in
# [0, 255] to uint8, index tensors whose entries fit to int32 -- and leaves everything else alone; ``DevicePrefetcher``
# copies the narrow tensors and widens them on the device in one launch (``fnb_widen_batch``), so the consumer sees the
# dtypes and values of ``collate_fn`` bit for bit.  It is meant to run where the collate runs (the DataLoader workers).
INDEX_KEYS = ("edge_index", "frag_index", "batch", "frag_batch", "atom_to_frag_ids", "edge_index_bonds_graph",
              "edge_index_fbonds")
ONE_HOT_KEYS = ("x_atoms", "x_frags", "edge_attr", "cnx_attr", "node_features_bonds", "node_features_fbonds",
                "edge_attr_fbonds")


def compact_batch(batch: Dict[str, torch.Tensor], pin: bool = False) -> Dict[str, torch.Tensor]:
    """The same batch dict with exactly-narrowable tensors narrowed (float32 -> uint8, int64 -> int32)."""
    out = {}
    for k, v in batch.items():
        t = v
        if isinstance(v, torch.Tensor) and v.device.type == "cpu" and v.numel() > 0:
            if k in ONE_HOT_KEYS and v.dtype == torch.float32:
                n = v.to(torch.uint8)
                if torch.equal(n.to(torch.float32), v):
                    t = n
            elif k in INDEX_KEYS and v.dtype == torch.int64:
                if int(v.min()) >= -2 ** 31 and int(v.max()) < 2 ** 31:
                    t = v.to(torch.int32)
        out[k] = t.pin_memory() if (pin and isinstance(t, torch.Tensor) and t.device.type == "cpu") else t
    return out


# ---- packed batches: one host buffer, one copy ----------------------------------------------------------------------
# A batch dict is ~20 tensors, i.e. ~20 cudaMemcpyAsync calls per step on the host thread that also launches the
# training step (0.27 ms of host time per 1 024-molecule batch, gpurun_out host_enqueue_probe).  ``pack_batch`` lays the
# tensors the GAT2 path reads out in ONE flat byte buffer (64-byte aligned segments) and hands back the same dict whose
# entries are views into it; ``DevicePrefetcher`` then moves the batch with a single copy.  ``PackedBatch.pin_memory``
# pins the buffer once and re-creates the views, which is what ``DataLoader(pin_memory=True)`` calls on a custom batch
# type.  The tensors ``FragNet.forward`` never reads stay ordinary entries (copied one by one only if asked for).
PACK_SKIP = ("edge_attr", "cnx_attr", "x_frags")


class PackedBatch(dict):
    """A batch dict whose hot-path tensors are views into ``blob`` (uint8, one allocation); ``layout`` lists
    ``(key, dtype, shape, byte offset, bytes)``."""
    blob: Optional[torch.Tensor] = None
    layout: tuple = ()

    def pin_memory(self):
        if self.blob is None or self.blob.is_pinned():
            return self
        return _packed_views(self, self.blob.pin_memory(), self.layout)


def _packed_views(src: dict, blob: torch.Tensor, layout) -> "PackedBatch":
    out = PackedBatch(src)
    out.blob, out.layout = blob, tuple(layout)
    for key, dtype, shape, off, nbytes in layout:
        out[key] = blob[off:off + nbytes].view(dtype).view(shape)
    return out


def pack_batch(batch: Dict[str, torch.Tensor], pin: bool = False) -> PackedBatch:
    """The same batch with the CPU tensors of the hot path re-homed as views of one flat buffer (see above)."""
    layout, off = [], 0
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and v.device.type == "cpu" and k not in PACK_SKIP:
            nbytes = v.numel() * v.element_size()
            layout.append((k, v.dtype, tuple(v.shape), off, nbytes))
            off += (nbytes + 63) // 64 * 64
    blob = torch.empty(max(off, 64), dtype=torch.uint8, pin_memory=pin)
    out = _packed_views(batch, blob, layout)
    for k, *_ in layout:
        out[k].copy_(batch[k])
    if pin:
        for k in PACK_SKIP:
            if isinstance(out.get(k), torch.Tensor) and out[k].device.type == "cpu":
                out[k] = out[k].pin_memory()
    return out


def collate_fn_pt_packed(data_list):
    """``collate_fn_pt`` in the compact wire format, packed into one buffer (one host -> device copy per batch)."""
    return pack_batch(compact_batch(_collate(data_list, pretrain=True)))


def collate_fn_packed(data_list):
    """``collate_fn`` in the compact wire format, packed into one buffer."""
    return pack_batch(compact_batch(_collate(data_list, pretrain=False)))


def collate_fn_compact(data_list):
    """``collate_fn`` in the compact wire format (widened back by ``DevicePrefetcher``)."""
    return compact_batch(_collate(data_list, pretrain=False))


def collate_fn_pt_compact(data_list):
    """``collate_fn_pt`` in the compact wire format (widened back by ``DevicePrefetcher``)."""
    return compact_batch(_collate(data_list, pretrain=True))
