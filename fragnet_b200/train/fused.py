"""One pretraining step per library call: ``FusedPretrainStep``.

The reference's loop body (fragnet/train/pretrain/pretrain_utils.py:12-30) is::

    optimizer.zero_grad(); preds = model(batch); loss = 2*MSE(dihedral)+MSE(angle)+MSE(energy)
    loss.backward(); optimizer.step()            # torch.optim.Adam(model.parameters(), lr)  (pretrain_gat2.py:165)

Through ``nn.Module`` / autograd this works unchanged on the drop-in modules; it costs ~2.4 ms of host time per step
at batch 1024, more than the device needs.  ``FusedPretrainStep`` runs the same arithmetic for a ``FragNetPreTrain``
model as ``fnb_pretrain_step`` (collate, forward, loss, backward: one call) + optional gradient all-reduce +
``fnb_adam_step`` (one launch over a flat parameter buffer).  The model's parameters stay ordinary ``nn.Parameter``
objects with their names and shapes (``state_dict`` round-trips); those that receive gradients are re-homed as views
into one flat buffer and their ``.grad`` are views into a second one.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .. import _abi, config, ops
from ..autograd import N_PARAMS, _F_INDEX

from .optim import flat_views as _flat_views


class FusedPretrainStep:
    """``step(batch) -> loss`` (0-dim CUDA tensor) for a ``FragNetPreTrain`` model; Adam hyper-parameters as
    ``torch.optim.Adam``.  ``group``: a ``torch.distributed`` process group for data-parallel training (mean of the
    per-rank gradients, the DDP semantics of finetune_gat2_pl.py:230); ``None`` uses the default group if one is
    initialised."""

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 group=None):
        enc, head = model.pretrain, model.head
        if not head._library_shapes():
            raise NotImplementedError("FusedPretrainStep needs PretrainTask(128, 1, L=2)")
        self.model, self.group = model, group
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.dev = dev = ops.require_cuda(enc.layers[0].a.device)
        for layer in enc.layers:
            layer._check_geometry()
        n_layers = len(enc.layers)
        layer_params = [list(layer._live_parameters()) for layer in enc.layers]
        head_params = head._parameters_in_library_order()
        # parameters that receive a gradient, in flat-buffer order
        live: List[torch.nn.Parameter] = []
        for li, ps in enumerate(layer_params):
            live += [p for j, p in enumerate(ps) if j != _F_INDEX or li == n_layers - 1]
        live += head_params[8:]                       # ba, da, FC stacks (the bond-length head is never trained)
        if any(p.dtype != torch.float32 or p.device != dev for p in live):
            raise ValueError("FusedPretrainStep needs fp32 parameters on one CUDA device")
        self.live = live
        self.flat_p, p_views = _flat_views(live, dev)
        self._peers = None
        g_store = self._symmetric_gradient_buffer(live, dev, group)
        self.flat_g, g_views = _flat_views(live, dev, storage=g_store)
        with torch.no_grad():
            torch._foreach_copy_(p_views, [p.data for p in live])
            for p, v, g in zip(live, p_views, g_views):
                p.data = v
                p.grad = g
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.t = 0
        self._p_views = p_views
        if self.world_size > 1:
            # DDP / Fabric broadcast rank 0's parameters at wrap time (finetune_gat2_pl.py:230): differently seeded
            # ranks, or a checkpoint loaded on one rank only, must not diverge silently
            dist.broadcast(self.flat_p, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        grad_of: Dict[int, torch.Tensor] = {id(p): g for p, g in zip(live, g_views)}
        # ---- C structs (built once: parameter storage never moves afterwards)
        self._layers = (_abi.CLayerParams * n_layers)()
        self._layer_grads = (_abi.CLayerGrads * n_layers)()
        for li, ps in enumerate(layer_params):
            L, G = self._layers[li], self._layer_grads[li]
            for name, p in zip(_abi.PARAM_FIELDS, ps):
                setattr(L, name, p.data_ptr())
                g = grad_of.get(id(p))
                setattr(G, name, None if g is None else g.data_ptr())
            L.K_bond, L.K_fbond, L.K_atom = ps[0].shape[1], ps[2].shape[1], ps[8].shape[1]
            L.run_frag_block, L.want_attention = int(li == n_layers - 1), 0
            L.bond_mask = L.frag_bond_mask = L.atom_mask = -1
            L.atom_mask_list, L.n_atom_mask = None, 0
        assert len(layer_params[0]) == N_PARAMS
        self._heads, self._head_grads = _abi.CPretrainHeadParams(), _abi.CPretrainHeadParams()
        self._heads.Wr, self._heads.br = head_params[0].data_ptr(), head_params[1].data_ptr()
        for gi, group_name in enumerate(("bl", "ba", "da", "fc")):
            for fi, fname in enumerate(_abi.MLP3_FIELDS):
                p = head_params[2 + gi * 6 + fi]
                setattr(getattr(self._heads, group_name), fname, p.data_ptr())
                g = grad_of.get(id(p))
                setattr(getattr(self._head_grads, group_name), fname, None if g is None else g.data_ptr())
        self._keep = (layer_params, head_params, p_views, g_views)
        self._ws: Optional[torch.Tensor] = None
        self._loss = torch.zeros(8, dtype=torch.float32, device=dev)      # ring of loss scalars
        self._loss_i = 0
        self.n_layers = n_layers
        # collate of the next batch underneath the running step (prefetch_plan): two alternating plan arenas
        self._plan_arenas: List[Optional[torch.Tensor]] = [None, None]
        self._step_parity: Optional[int] = None     # arena the step launched last reads, if it found its plan prefetched
        self._prefetched = None          # (key of the batch inputs, arena, tensors kept alive)
        self.plan_hits = 0               # steps that found their plan prefetched

    # ------------------------------------------------------------------------------------------
    def _symmetric_gradient_buffer(self, live, dev, group):
        """World > 1: the flat gradient buffer is allocated in symmetric memory (peer-mapped over NVLink at rendezvous)
        so that ``fnb_allreduce_adam_step`` can read every rank's gradients directly -- exchange, mean and Adam are then
        one kernel instead of ncclAllReduce + scale + Adam.  Returns the storage tensor, or None (single rank, no NVLink
        peer access, ``FNB_FUSED_ALLREDUCE=0``): ``step`` then uses ``dist.all_reduce`` + ``fnb_adam_step``."""
        import os
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
            return None
        w = dist.get_world_size(group)
        if os.environ.get("FNB_FUSED_ALLREDUCE", "1") == "0" or w > 8 or dist.get_backend(group) != "nccl":
            return None
        from .optim import ALIGN
        total = sum((p.numel() + ALIGN - 1) // ALIGN * ALIGN for p in live)
        try:
            import torch.distributed._symmetric_memory as symm
            pg = group if group is not None else dist.group.WORLD
            g = symm.empty(total, dtype=torch.float32, device=dev)
            flags = symm.empty(2 * w, dtype=torch.int32, device=dev)
            g.zero_()
            flags.zero_()
            hg, hf = symm.rendezvous(g, pg), symm.rendezvous(flags, pg)
            ps = _abi.CPeerSet()
            for r in range(w):
                ps.grads[r], ps.flags[r] = hg.buffer_ptrs[r], hf.buffer_ptrs[r]
            ps.world, ps.rank = w, hg.rank
            ok = torch.ones(1, device=dev)
        except Exception as ex:          # no P2P between these devices, or an older torch: NCCL path
            import warnings
            warnings.warn(f"fragnet_b200: symmetric-memory gradient exchange unavailable ({type(ex).__name__}: {ex}); "
                          "using ncclAllReduce", RuntimeWarning)
            g, ok = None, torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)       # all ranks take the same path (zeroed flags visible)
        if float(ok) < 1.0 or g is None:
            return None
        self._peers, self._flags, self._epoch = ps, flags, 0
        self._done = torch.zeros(1, dtype=torch.int32, device=dev)
        self._symm_keep = (hg, hf)
        return g

    def bound_to_model(self) -> bool:
        """True while every live parameter still is its view of the flat buffer (``model.to()``, ``.cpu()`` or a
        re-assigned ``p.data`` break the binding; the caller then has to fall back to the generic loop)."""
        return all(p.data_ptr() == v.data_ptr() and p.device == v.device for p, v in zip(self.live, self._p_views))

    def attach_optimizer(self, optimizer) -> None:
        """Keep a caller-owned ``torch.optim.Adam`` truthful: its ``state`` entries of the live parameters become VIEWS
        of this object's moment buffers plus a shared step counter, so ``optimizer.state_dict()`` checkpoints the real
        state; moments already present in ``optimizer.state`` (a resumed run) are adopted first."""
        from .optim import flat_views
        _, m_views = flat_views(self.live, self.dev, storage=self.exp_avg)
        _, v_views = flat_views(self.live, self.dev, storage=self.exp_avg_sq)
        steps = [int(optimizer.state[p]["step"]) for p in self.live if p in optimizer.state and "step" in optimizer.state[p]]
        with torch.no_grad():
            for p, m, v in zip(self.live, m_views, v_views):
                st = optimizer.state.get(p)
                if st and "exp_avg" in st:
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
        if steps:
            self.t = max(steps)
        self._opt_step = torch.tensor(float(self.t))
        for p, m, v in zip(self.live, m_views, v_views):
            optimizer.state[p] = {"step": self._opt_step, "exp_avg": m, "exp_avg_sq": v}

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _args(self, batch, backward: bool, want_preds: bool):
        dev = self.dev

        def idx(k):
            t = batch[k]
            if t.device != dev or t.dtype != torch.int64 or not t.is_contiguous():
                t = t.to(device=dev, dtype=torch.int64).contiguous()
            return t

        def f32(k):
            t = batch[k]
            if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=dev, dtype=torch.float32).contiguous()
            return t

        ei, fi, a2f = idx("edge_index"), idx("frag_index"), idx("atom_to_frag_ids")
        eb, efb = idx("edge_index_bonds_graph"), idx("edge_index_fbonds")
        bv, fbv = idx("batch"), idx("frag_batch")
        cos, a6 = f32("edge_attr_bonds"), f32("edge_attr_fbonds")
        xa, xb, xfb = f32("x_atoms"), f32("node_features_bonds"), f32("node_features_fbonds")
        t_ba, t_dh, y = f32("bnd_angl"), f32("dh_angl"), f32("y")
        na, nf, nb, nfb, g = xa.shape[0], batch["x_frags"].shape[0], xb.shape[0], xfb.shape[0], y.numel()
        if ei.shape[1] != nb or fi.shape[1] != nfb:
            raise ValueError("fragnet_b200: the bond graph needs one node per column of edge_index and the "
                             "fragment-connection graph one node per column of frag_index")
        if t_ba.numel() != na or t_dh.numel() != nb:
            raise ValueError("fragnet_b200: bnd_angl / dh_angl must have one entry per atom / directed bond")
        pt = ops._ptr
        a = _abi.CPretrainStepArgs()
        a.batch = _abi.CBatchInputs(pt(ei), pt(fi), pt(a2f), pt(eb), pt(efb), pt(bv), pt(fbv), pt(cos), pt(a6),
                                    na, nf, nb, eb.shape[1], nfb, efb.shape[1], g)
        a.x_atoms, a.x_bond, a.x_fbond = pt(xa), pt(xb), pt(xfb)
        a.t_bond_angle, a.t_dihedral, a.t_energy = pt(t_ba), pt(t_dh), pt(y)
        a.n_layers = self.n_layers
        a.layers = C.cast(self._layers, C.c_void_p)
        a.layer_grads = C.cast(self._layer_grads, C.c_void_p)
        a.heads = C.cast(C.pointer(self._heads), C.c_void_p)
        a.head_grads = C.cast(C.pointer(self._head_grads), C.c_void_p)
        enc = self.model.pretrain
        training = bool(self.model.training)
        a.drop_p, a.training = float(enc.dropout.p), int(training)
        a.seed, a.offset = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, 0
        a.precision, a.backward = config.precision_id(), int(backward)
        self._loss_i = (self._loss_i + 1) % self._loss.numel()
        loss = self._loss[self._loss_i]
        a.loss = loss.data_ptr()
        preds = None
        if want_preds:
            new = lambda n: torch.empty((n, 1), dtype=torch.float32, device=dev)
            preds = (new(nb), new(na), new(nb), new(g))
            a.bond_length, a.bond_angle, a.dihedral, a.energy = (pt(t) for t in preds)
        keep = (ei, fi, a2f, eb, efb, bv, fbv, cos, a6, xa, xb, xfb, t_ba, t_dh, y)
        pre, self._prefetched = self._prefetched, None       # one prefetch serves one step
        self._step_parity = None
        if pre is not None and pre[0] == self._plan_key(a.batch):
            a.plan_arena = pre[1].data_ptr()       # collated ahead of this step (see prefetch_plan)
            self._step_parity = pre[3]             # ... and this arena is now in use until the step has run
            self.plan_hits += 1
        return a, loss, preds, keep

    # ---- collate of the next batch underneath the running step -------------------------------------------------------
    _PLAN_TENSORS = ("edge_index", "frag_index", "atom_to_frag_ids", "edge_index_bonds_graph", "edge_index_fbonds",
                     "batch", "frag_batch")

    @staticmethod
    def _plan_key(inp: "_abi.CBatchInputs"):
        return tuple(getattr(inp, n) for n, _ in _abi.CBatchInputs._fields_)

    def prefetch_plan(self, next_batch, ready_event: Optional["torch.cuda.Event"] = None) -> bool:
        """Queue the on-device collate (``fnb_batch_plan_build``) of the batch the NEXT ``step`` call will get, on a
        library stream underneath the step that is running -- the plan depends on nothing but the batch's index tensors,
        and without it the first attention kernel of a step waits ~50 us for nine small launches that compete with the
        layer-0 GEMMs for SMs.  ``ready_event``: recorded once the tensors of ``next_batch`` are complete (a
        ``DevicePrefetcher`` slot still being copied); None if they already are.  Only batches that need no dtype /
        device conversion qualify (returns False otherwise: the step then collates by itself, as always).  The tensors
        must not be modified until that step has been launched."""
        self._prefetched = None
        dev = self.dev
        for k in self._PLAN_TENSORS:
            t = next_batch.get(k)
            if not isinstance(t, torch.Tensor) or t.device != dev or t.dtype != torch.int64 or not t.is_contiguous():
                return False
        for k in ("edge_attr_bonds", "edge_attr_fbonds"):
            t = next_batch.get(k)
            if not isinstance(t, torch.Tensor) or t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                return False
        b = next_batch
        na, nf = b["x_atoms"].shape[0], b["x_frags"].shape[0]
        nb, nfb, g = b["edge_index"].shape[1], b["frag_index"].shape[1], b["y"].numel()
        pt = ops._ptr
        inp = _abi.CBatchInputs(pt(b["edge_index"]), pt(b["frag_index"]), pt(b["atom_to_frag_ids"]),
                                pt(b["edge_index_bonds_graph"]), pt(b["edge_index_fbonds"]), pt(b["batch"]),
                                pt(b["frag_batch"]), pt(b["edge_attr_bonds"]), pt(b["edge_attr_fbonds"]),
                                na, nf, nb, b["edge_index_bonds_graph"].shape[1], nfb, b["edge_index_fbonds"].shape[1], g)
        lib = ops._lib()
        need = lib.fnb_batch_plan_bytes(C.byref(inp))
        if need == 0:
            return False
        # never the arena the step launched last reads (its backward still needs the plan); a second prefetch before the
        # next step simply replaces the first
        parity = 0 if self._step_parity is None else 1 - self._step_parity
        arena = self._plan_arenas[parity]
        if arena is None or arena.numel() < need:
            # (re)allocation: the block may have been freed by work still running on the compute stream, and the collate
            # writes it from another stream -- rare (first steps, a larger batch), so simply wait
            torch.cuda.current_stream(dev).synchronize()
            arena = torch.empty(int(need * 1.25) + 256, dtype=torch.uint8, device=dev)
            self._plan_arenas[parity] = arena
        ev = C.c_void_p(ready_event.cuda_event) if ready_event is not None else None
        rc = lib.fnb_pretrain_plan_prefetch(C.byref(inp), ops._ptr(arena), arena.numel(), ev)
        if rc != 0:
            return False                 # single-stream mode: the step collates by itself
        keep = tuple(b[k] for k in self._PLAN_TENSORS) + (b["edge_attr_bonds"], b["edge_attr_fbonds"])
        self._prefetched = (self._plan_key(inp), arena, keep, parity)
        return True

    def _run(self, batch, backward: bool, want_preds: bool = False):
        lib = ops._lib()
        a, loss, preds, keep = self._args(batch, backward, want_preds)
        if a.training and a.drop_p > 0:
            a.offset = ops.reserve_rng(lib.fnb_pretrain_step_rng_span(C.byref(a)))
        need = lib.fnb_pretrain_step_workspace_bytes(C.byref(a))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(int(need * 1.25) + 256, dtype=torch.uint8, device=self.dev)
        _abi.check(lib.fnb_pretrain_step(C.byref(a), ops._ptr(self._ws), self._ws.numel(), ops._ptr(ops.scratch(self.dev)),
                                         ops._stream()), "pretrain_step")
        del keep        # stream-ordered allocator: safe to release once the launches are queued on this stream
        return loss, preds

    def step(self, batch, next_batch=None, next_ready=None) -> torch.Tensor:
        """collate + forward + loss + backward (+ all-reduce) + Adam on one batch dict; returns the loss.
        ``next_batch`` (optional): the batch of the following call, whose collate is then queued underneath this step
        (``prefetch_plan``); ``next_ready``: the CUDA event after which its tensors are complete, if they are not yet."""
        loss = self._step(batch)
        if next_batch is not None:
            self.prefetch_plan(next_batch, next_ready)
        return loss

    def _step(self, batch) -> torch.Tensor:
        loss, _ = self._run(batch, backward=True)
        w = self.world_size
        self.t += 1
        if w > 1 and self._peers is not None:
            # gradient exchange + mean + Adam in one kernel over NVLink peer memory (csrc/dist.cu)
            self._epoch += 1
            _abi.check(ops._lib().fnb_allreduce_adam_step(
                C.byref(self._peers), ops._p(self.flat_p), ops._p(self.exp_avg), ops._p(self.exp_avg_sq),
                self.flat_p.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.t,
                self._epoch & 0xFFFFFFFF, ops._p(self._done), ops._stream()), "allreduce_adam_step")
            if getattr(self, "_opt_step", None) is not None:
                self._opt_step.fill_(float(self.t))
            return loss
        if w > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
            self.flat_g.mul_(1.0 / w)
        _abi.check(ops._lib().fnb_adam_step(ops._p(self.flat_p), ops._p(self.flat_g), ops._p(self.exp_avg),
                                            ops._p(self.exp_avg_sq), self.flat_p.numel(), self.lr, self.betas[0],
                                            self.betas[1], self.eps, self.weight_decay, self.t, ops._stream()), "adam_step")
        if getattr(self, "_opt_step", None) is not None:
            self._opt_step.fill_(float(self.t))
        return loss

    def forward_backward(self, batch) -> torch.Tensor:
        """Gradients only (``p.grad`` of the live parameters are views of ``flat_g``), no parameter update."""
        return self._run(batch, backward=True)[0]

    @torch.no_grad()
    def evaluate(self, batch, return_predictions: bool = False):
        """Forward + loss without gradients (Trainer.validate, pretrain_utils.py:33-57)."""
        loss, preds = self._run(batch, backward=False, want_preds=return_predictions)
        return (loss, preds) if return_predictions else loss


class LaggedScalars:
    """Every step's scalar result (the loss) read back to the host WITHOUT stalling the launch of the next step.

    The reference reads ``loss.item()`` right after ``optimizer.step()`` (pretrain_utils.py:28): the host then waits
    for the whole step and the GPU idles while the next step is being enqueued.  Here the 4-byte device-to-host copy
    of step ``i`` is queued behind step ``i`` into a pinned slot, and collected once step ``i + lag`` has been enqueued
    (``push`` returns the values that have become due, in step order; ``drain`` returns the rest)."""

    def __init__(self, lag: int = 1):
        import collections
        self.lag = max(0, int(lag))
        n = self.lag + 2
        self._host = torch.empty(n, dtype=torch.float32).pin_memory()
        self._events = [torch.cuda.Event() for _ in range(n)]
        self._ready = [torch.cuda.Event() for _ in range(n)]
        self._side: Optional[torch.cuda.Stream] = None
        self._pending = collections.deque()
        self._count = 0

    def _collect(self, keep: int) -> List[float]:
        out = []
        while len(self._pending) > keep:
            s = self._pending.popleft()
            self._events[s].synchronize()
            out.append(float(self._host[s]))
        return out

    def push(self, value: torch.Tensor) -> List[float]:
        s = self._count % len(self._events)
        self._count += 1
        # the copy runs on a side stream behind the step: in the compute stream the 4-byte DMA would sit between this
        # step's last kernel and the next step's first one (~10 us per step, gpurun_out/r5x)
        cur = torch.cuda.current_stream(value.device)
        if self._side is None:
            self._side = torch.cuda.Stream(value.device)
        src = value.detach().reshape(1)
        self._ready[s].record(cur)
        self._side.wait_event(self._ready[s])
        with torch.cuda.stream(self._side):
            self._host[s:s + 1].copy_(src, non_blocking=True)
        src.record_stream(self._side)
        self._events[s].record(self._side)
        self._pending.append(s)
        return self._collect(self.lag)

    def drain(self) -> List[float]:
        return self._collect(0)
