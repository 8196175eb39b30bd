"""Packed per-dataset arena + on-device batch assembly.

The reference assembles every batch on the host: ``collate_fn`` / ``collate_fn_pt`` (fragnet/dataset/data.py:877-948,
:951-1032) ``torch.cat`` the per-molecule tensors and add node-count offsets built by Python loops over the molecule
list (``get_incr_*``, data.py:11-113), then the training loop moves all 16 / 19 tensors with ``batch[k].to(device)``
(train/pretrain/pretrain_utils.py:13-14).  At B200 step rates that host work and the ~40 MB of PCIe traffic per
1 024-molecule batch are the end-to-end bound.

``MoleculeArena`` uploads the dataset ONCE (all molecules' tensors concatenated per key, index tensors kept
molecule-local as int32) and ``arena.batch(ids)`` produces, with two kernel launches (``fnb_arena_assemble``), exactly the
dict ``collate_fn_pt([data_list[i] for i in ids])`` would give after ``.to(device)``: same keys, shapes, dtypes and
values.  Per step the host only ships the molecule ids and sums a few per-molecule counts (to size the outputs).
``ArenaLoader`` is the ``DataLoader(dataset, collate_fn=collate_fn_pt, ...)`` replacement for the training loops.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch

from .. import _abi

# batch key -> (record attribute, concatenation dim).  Order = key order of the reference's dict (data.py:931-948).
_FEATURES = [
    ("x_atoms", "x_atoms"), ("x_frags", "x_frags"), ("edge_attr", "edge_attr"), ("cnx_attr", "cnx_attr"),
    ("node_features_bonds", "node_features_bonds"), ("edge_attr_bonds", "edge_attr_bonds"),
    ("node_features_fbonds", "node_feautures_fbondg"), ("edge_attr_fbonds", "edge_attr_fbondg"),
]
_PRETRAIN = [("bnd_lngth", "bnd_lngth"), ("bnd_angl", "bnd_angl"), ("dh_angl", "dh_angl")]
# index key -> (record attribute, node-count kind whose batch prefix is added)   (get_incr_*, data.py:11-113)
_INDEX2 = [
    ("edge_index", "edge_index", "n_atoms"), ("frag_index", "frag_index", "n_frags"),
    ("edge_index_bonds_graph", "edge_index_bonds", "n_bnodes"), ("edge_index_fbonds", "edge_index_fbondg", "n_fbnodes"),
]
_KEY_ORDER = ["x_atoms", "edge_index", "frag_index", "x_frags", "edge_attr", "cnx_attr", "batch", "frag_batch",
              "atom_to_frag_ids", "node_features_bonds", "edge_index_bonds_graph", "edge_attr_bonds",
              "node_features_fbonds", "edge_index_fbonds", "edge_attr_fbonds", "bnd_lngth", "bnd_angl", "dh_angl", "y"]


class _Kind:
    def __init__(self, counts: np.ndarray, device):
        self.counts_host = counts.astype(np.int64)
        prefix = np.zeros(len(counts), dtype=np.int64)
        if len(counts) > 1:
            np.cumsum(self.counts_host[:-1], out=prefix[1:])
        if len(counts) and int(self.counts_host.max()) >= 2 ** 31:
            raise ValueError("a molecule with >= 2^31 rows does not fit the arena")
        self.counts = torch.from_numpy(self.counts_host.astype(np.int32)).to(device)
        self.prefix = torch.from_numpy(prefix).to(device)


class MoleculeArena:
    """Device-resident packed dataset.  ``data_list``: the per-molecule records the reference's collate takes
    (attributes of ``CreateData.create_data_point``, data.py:437-482)."""

    def __init__(self, data_list: Sequence, device, pretrain: Optional[bool] = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("MoleculeArena lives in CUDA device memory (there is no CPU path)")
        self.n_mols = len(data_list)
        if pretrain is None:
            pretrain = self.n_mols > 0 and hasattr(data_list[0], "bnd_lngth")
        self.pretrain = bool(pretrain)
        self._lib = _abi.load()
        dev = self.device
        self._kinds: Dict[str, _Kind] = {}
        self._kind_ids: Dict[str, int] = {}
        self._store: Dict[str, torch.Tensor] = {}
        self._jobs: List[tuple] = []     # (key, row, src tensor, kind id, offset kind id, width, mode)
        self._shape: Dict[str, tuple] = {}   # key -> (layout, dtype, trailing shape, kind name)

        def kind(name: str, counts) -> int:
            if name not in self._kinds:
                if len(self._kinds) >= _abi.ARENA_MAX_KINDS:
                    raise ValueError("too many row-count kinds")
                self._kinds[name] = _Kind(np.asarray(counts, dtype=np.int64), dev)
                self._kind_ids[name] = len(self._kind_ids)
            return self._kind_ids[name]

        def cat(parts, dim=0):
            return torch.cat(list(parts), dim=dim) if self.n_mols else torch.zeros(0)

        # node-count kinds (the offsets of the index tensors)
        kind("n_atoms", [d.x_atoms.shape[0] for d in data_list])
        kind("n_frags", [int(d.n_frags.item()) for d in data_list])
        kind("n_bnodes", [d.node_features_bonds.shape[0] for d in data_list])
        kind("n_fbnodes", [d.node_feautures_fbondg.shape[0] for d in data_list])

        feats = list(_FEATURES) + (list(_PRETRAIN) if self.pretrain else [])
        for key, attr in feats:
            parts = [getattr(d, attr) for d in data_list]
            if any(p.dtype != torch.float32 for p in parts):
                raise TypeError(f"{key}: the arena stores float32 feature tensors")
            k = kind("rows:" + key, [p.shape[0] for p in parts])
            trailing = tuple(parts[0].shape[1:]) if parts else ()
            width = int(np.prod(trailing)) if trailing else 1
            self._store[key] = cat(parts).contiguous().to(dev)
            self._shape[key] = ("rows", torch.float32, trailing, "rows:" + key)
            self._jobs.append((key, 0, self._store[key], k, 0, width, _abi.ARENA_COPY32))
        # y: torch.cat(...).type(torch.float)  (data.py:946)
        ys = [d.y.to(torch.float32) for d in data_list]
        k = kind("rows:y", [p.shape[0] for p in ys])
        trailing = tuple(ys[0].shape[1:]) if ys else ()
        self._store["y"] = cat(ys).contiguous().to(dev)
        self._shape["y"] = ("rows", torch.float32, trailing, "rows:y")
        self._jobs.append(("y", 0, self._store["y"], k, 0, int(np.prod(trailing)) if trailing else 1, _abi.ARENA_COPY32))
        # [2, E] index tensors: one flat molecule-local int32 array per row
        for key, attr, node_kind in _INDEX2:
            parts = [getattr(d, attr) for d in data_list]
            k = kind("cols:" + key, [p.shape[1] for p in parts])
            for r in (0, 1):
                flat = cat([p[r].to(torch.int64) for p in parts])
                if flat.numel() and (int(flat.max()) >= 2 ** 31 or int(flat.min()) < 0):
                    raise ValueError(f"{key}: molecule-local indices must fit int32")
                self._store[f"{key}/{r}"] = flat.to(torch.int32).contiguous().to(dev)
                self._jobs.append((key, r, self._store[f"{key}/{r}"], k, self._kind_ids[node_kind], 1, _abi.ARENA_INDEX))
            self._shape[key] = ("index2", torch.int64, (), "cols:" + key)
        # atom -> fragment membership: 1-D, offset by the fragments before the molecule (data.py:59-75)
        parts = [d.atom_id_frag_id.to(torch.int64) for d in data_list]
        k = kind("rows:atom_to_frag_ids", [p.shape[0] for p in parts])
        self._store["atom_to_frag_ids"] = cat(parts).to(torch.int32).contiguous().to(dev)
        self._shape["atom_to_frag_ids"] = ("rows", torch.int64, (), "rows:atom_to_frag_ids")
        self._jobs.append(("atom_to_frag_ids", 0, self._store["atom_to_frag_ids"], k, self._kind_ids["n_frags"], 1,
                           _abi.ARENA_INDEX))
        # molecule position per atom / fragment (data.py:905-913)
        self._shape["batch"] = ("rows", torch.int64, (), "n_atoms")
        self._jobs.append(("batch", 0, None, self._kind_ids["n_atoms"], 0, 1, _abi.ARENA_FILL))
        self._shape["frag_batch"] = ("rows", torch.int64, (), "n_frags")
        self._jobs.append(("frag_batch", 0, None, self._kind_ids["n_frags"], 0, 1, _abi.ARENA_FILL))
        if len(self._jobs) > _abi.ARENA_MAX_JOBS:
            raise ValueError("too many arena jobs")

        self._ckinds = (_abi.CArenaKind * len(self._kinds))()
        for name, i in self._kind_ids.items():
            self._ckinds[i].counts = self._kinds[name].counts.data_ptr()
            self._ckinds[i].prefix = self._kinds[name].prefix.data_ptr()
        self._keys = [k for k in _KEY_ORDER if k in self._shape]
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._ws: Optional[torch.Tensor] = None
        self.nbytes = sum(t.numel() * t.element_size() for t in self._store.values())

    def __len__(self) -> int:
        return self.n_mols

    _ID_SLOTS = 8
    SLOTS = 3        # batches alive at a time (see ``batch``)

    def _slot_buffers(self, sizes):
        """Flat output buffers of the next slot (grow-only, reused round-robin).  Going through the caching allocator
        for ~80 MB per batch costs a ``cudaMalloc`` -- milliseconds, device-synchronising -- whenever the training
        step's own workspaces have fragmented the pool (measured: 0.24 ms -> 2.9 ms per call inside the step loop)."""
        ring = getattr(self, "_slots", None)
        if ring is None:
            ring = self._slots = [dict() for _ in range(self.SLOTS)]
            self._slot_next = 0
        slot = ring[self._slot_next]
        self._slot_next = (self._slot_next + 1) % self.SLOTS
        for dt, n in sizes.items():
            if dt not in slot or slot[dt].numel() < max(n, 1):
                slot[dt] = torch.empty(int(max(n, 1) * 1.25) + 64, dtype=dt, device=self.device)
        return slot

    def _stage_ids(self, ids: np.ndarray) -> torch.Tensor:
        """Molecule ids -> device through a ring of persistent pinned buffers (a fresh ``pin_memory()`` per batch
        ends in ``cudaHostAlloc`` -- a device-synchronising call -- as soon as the host runs ahead of the GPU)."""
        g = int(ids.shape[0])
        ring = getattr(self, "_id_ring", None)
        if ring is None or ring[0][0].numel() < g:
            cap = max(g, 1024)
            ring = [(torch.empty(cap, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(self._ID_SLOTS)]
            self._id_ring, self._id_next, self._id_used = ring, 0, [False] * self._ID_SLOTS
        k = self._id_next
        self._id_next = (k + 1) % self._ID_SLOTS
        host, ev = ring[k]
        if self._id_used[k]:
            ev.synchronize()                     # the copy that last read this slot (8 batches ago) has completed
        host[:g].numpy()[:] = ids
        out = torch.empty(g, dtype=torch.int64, device=self.device)
        out.copy_(host[:g], non_blocking=True)
        ev.record(torch.cuda.current_stream(self.device))
        self._id_used[k] = True
        return out

    def batch_nbytes(self, ids) -> int:
        """Bytes of the batch dict the reference would move host -> device for these molecules."""
        ids = np.asarray(ids, dtype=np.int64)
        total = 0
        for key in self._keys:
            layout, dtype, trailing, kname = self._shape[key]
            rows = int(self._kinds[kname].counts_host[ids].sum())
            width = int(np.prod(trailing)) if trailing else 1
            total += rows * width * (2 if layout == "index2" else 1) * (8 if dtype == torch.int64 else 4)
        return total

    def batch_overlapped(self, mol_ids) -> Dict[str, torch.Tensor]:
        """``batch`` on a side stream: the assembly of the next batch runs underneath the step that is in flight instead
        of between two steps (~45 us per 1 024-molecule batch).  The current stream waits for it (so whatever the caller
        launches next sees the batch), ``last_event`` marks its completion for consumers on other streams (the collate
        of ``FusedPretrainStep.prefetch_plan``).  The slot this batch overwrites must have been consumed by work queued
        on the current stream before the PREVIOUS request (true for a loop that asks for batch n+1 after launching
        step n, and for ``DevicePrefetcher`` with depth <= ``SLOTS`` - 1): the assembly waits for exactly that work."""
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(dev)
            self._req_events = [torch.cuda.Event(), torch.cuda.Event()]
            self._n_req = 0
        n = self._n_req
        self._n_req += 1
        if n >= 1:
            self._side.wait_event(self._req_events[(n - 1) % 2])    # everything queued before the previous request
        self._req_events[n % 2].record(cur)
        with torch.cuda.stream(self._side):
            out = self.batch(mol_ids)
        done = torch.cuda.Event()
        done.record(self._side)
        cur.wait_event(done)
        self.last_event = done
        out = DeviceBatch(out)
        out.ready_event = done
        return out

    def batch(self, mol_ids, ids_device: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """The batch dict of ``collate_fn_pt([data_list[i] for i in mol_ids])`` on the device, assembled on the device.

        ``mol_ids``: host sequence / numpy array / CPU tensor of molecule indices (they size the outputs);
        ``ids_device``: the same ids already on the device (otherwise they are copied from pinned memory on the
        current stream).  No host synchronisation, no allocation in the steady state: the returned tensors are views
        into one of ``SLOTS`` persistent device slots and stay valid until ``SLOTS`` further batches have been
        assembled (the contract of ``DevicePrefetcher``); work that reads them must be ordered before that assembly
        on the same stream (``.clone()`` what has to live longer)."""
        if isinstance(mol_ids, torch.Tensor):
            ids = mol_ids.detach().cpu().numpy().astype(np.int64, copy=False)
        else:
            ids = np.asarray(mol_ids, dtype=np.int64)
        g = int(ids.shape[0])
        if g and (int(ids.min()) < 0 or int(ids.max()) >= self.n_mols):
            raise IndexError("molecule id out of range")
        dev = self.device
        if ids_device is None:
            ids_device = self._stage_ids(ids)
        out: Dict[str, torch.Tensor] = {}
        if getattr(self, "_count_matrix", None) is None:      # [n_kinds, n_mols]: one fancy-index + one sum per batch
            self._count_names = list(self._kinds)
            self._count_matrix = np.stack([self._kinds[n].counts_host for n in self._count_names]) \
                if self.n_mols else np.zeros((len(self._count_names), 0), dtype=np.int64)
        sums = self._count_matrix[:, ids].sum(axis=1)
        totals = {name: int(v) for name, v in zip(self._count_names, sums)}
        # the tensors of a batch are 64-element aligned views into one flat buffer per dtype of a persistent slot
        shapes, sizes = {}, {torch.float32: 0, torch.int64: 0}
        for key in self._keys:
            layout, dtype, trailing, kname = self._shape[key]
            n = totals[kname]
            shape = (2, n) if layout == "index2" else (n,) + tuple(trailing)
            numel = int(np.prod(shape))
            shapes[key] = (shape, dtype, sizes[dtype], numel)
            sizes[dtype] += (numel + 63) // 64 * 64
        flat = self._slot_buffers(sizes)
        for key, (shape, dtype, off, numel) in shapes.items():
            out[key] = flat[dtype][off:off + numel].view(shape)
        if g == 0:
            return out
        if getattr(self, "_cjobs", None) is None:     # static fields once; only the destinations change per batch
            self._cjobs = (_abi.CArenaJob * len(self._jobs))()
            for i, (key, row, src, kind, okind, width, mode) in enumerate(self._jobs):
                self._cjobs[i].src = src.data_ptr() if src is not None else None
                self._cjobs[i].kind, self._cjobs[i].offset_kind = kind, okind
                self._cjobs[i].width, self._cjobs[i].mode = width, mode
        cjobs = self._cjobs
        for i, (key, row, _src, _k, _o, _w, _m) in enumerate(self._jobs):
            dst = out[key]
            cjobs[i].dst = dst.data_ptr() + (row * dst.shape[1] * 8 if row else 0)
        need = self._lib.fnb_arena_workspace_bytes(g, len(self._kinds))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(int(need * 1.5) + 256, dtype=torch.uint8, device=dev)
        rc = self._lib.fnb_arena_assemble(ids_device.data_ptr(), g, self.n_mols, self._ckinds, len(self._kinds), cjobs,
                                          len(self._jobs), self._ws.data_ptr(), self._ws.numel(),
                                          self._status.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"fnb_arena_assemble: {self._lib.fnb_error_string(rc).decode()}")
        return out


def epoch_batches(n: int, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                  generator: Optional[torch.Generator] = None, rank: int = 0, world: int = 1):
    """Molecule ids of every batch of one epoch for this rank (host-side, no device work).

    ``world == 1``: the batches of ``DataLoader(dataset, batch_size, shuffle, drop_last)``.  ``world > 1`` (SURVEY.md
    section 8(e): molecules are independent, rank ``r`` takes a contiguous share of every step): the epoch is cut into
    global batches of ``batch_size * world`` molecules (``batch_size`` = per-GPU batch, BASELINE configs[3]) and rank
    ``r`` gets slice ``[r * batch_size, (r + 1) * batch_size)`` of each; every rank must pass a generator in the same
    state so that the shuffles agree.  All ranks get the same NUMBER of batches (a ragged tail is split evenly; with
    ``drop_last`` the incomplete global batch is dropped), so the gradient all-reduce never waits for a missing step."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank / world")
    order = torch.randperm(n, generator=generator).numpy() if shuffle else np.arange(n, dtype=np.int64)
    gb = batch_size * world
    out = []
    for start in range(0, n, gb):
        chunk = order[start:start + gb]
        if len(chunk) < gb:
            if drop_last or len(chunk) < world:
                break
            per = len(chunk) // world             # ragged tail: equal shares, the remainder (< world) is dropped
            out.append(chunk[rank * per:(rank + 1) * per])
        else:
            out.append(chunk[rank * batch_size:(rank + 1) * batch_size])
    return out


class DeviceBatch(dict):
    """A device-resident batch dict with the CUDA event after which its tensors are complete (``ready_event``): consumers
    on the stream that requested it need nothing (that stream already waits), others (``DevicePrefetcher.last_event`` ->
    ``FusedPretrainStep.prefetch_plan``) wait for the event instead of for the requesting stream."""
    ready_event = None


class ArenaLoader:
    """``DataLoader(dataset, batch_size, shuffle, drop_last, collate_fn=collate_fn_pt)`` over a ``MoleculeArena``:
    iterates device-resident batch dicts (reference loops: pretrain_gat2.py:140-147, pretrain_utils.py:12-14).
    ``rank`` / ``world`` shard every step over the data-parallel ranks (``epoch_batches``)."""

    def __init__(self, arena: MoleculeArena, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                 generator: Optional[torch.Generator] = None, rank: int = 0, world: int = 1, overlap: bool = True):
        """``overlap``: assemble every batch on a side stream underneath the step in flight
        (``MoleculeArena.batch_overlapped``; its contract holds for loops that launch step n before they ask for batch
        n+1 and for ``DevicePrefetcher`` up to depth 2)."""
        self.arena, self.batch_size, self.shuffle, self.drop_last = arena, int(batch_size), shuffle, drop_last
        self.generator, self.rank, self.world = generator, int(rank), int(world)
        self.overlap = bool(overlap)
        self.dataset = arena         # the loops divide by len(loader.dataset) (pretrain_utils.py:31, utils.py:351)

    def __len__(self) -> int:
        return len(epoch_batches(len(self.arena), self.batch_size, False, self.drop_last, None, self.rank, self.world))

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        for ids in epoch_batches(len(self.arena), self.batch_size, self.shuffle, self.drop_last, self.generator,
                                 self.rank, self.world):
            yield self.arena.batch_overlapped(ids) if self.overlap else self.arena.batch(ids)
