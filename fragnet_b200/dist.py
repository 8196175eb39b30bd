"""Data-parallel gradient exchange for GAT2 training: one process per GPU, one flat all-reduce.

The reference's only multi-GPU code is Lightning Fabric DDP (fragnet/train/finetune/finetune_gat2_pl.py:230,
``fabric.backward``): mean-reduced gradients across ranks.  Molecules are independent (block-diagonal
batches, fragnet/dataset/data.py:877-948) so the forward/backward needs no communication; the only
exchange is the gradient of the LIVE parameters (87 of FragNet's tensors never receive a gradient,
SURVEY.md fact 7, and are skipped exactly as DDP/Adam skip ``grad is None``).

Gradients of live parameters are views into one flat fp32 buffer, so the exchange is a single
``all_reduce`` (NCCL over NVLink/NVSwitch; gloo in the CPU tests) with no pack/unpack copies.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class FlatGradSync:
    """Call ``prepare()`` once after the first backward, then per step: ``zero()``, backward, ``sync()``."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.all_params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.live: List[torch.nn.Parameter] = []
        self.flat = None

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def prepare(self) -> None:
        """Find the parameters that received a gradient and rebind their ``.grad`` into one flat buffer.
        The live set must agree across ranks (it is a property of the model, not the batch)."""
        self.live = [p for p in self.all_params if p.grad is not None]
        if self.world_size > 1:
            n = torch.tensor([len(self.live), sum(p.numel() for p in self.live)], dtype=torch.int64,
                             device=self.live[0].device)
            lo, hi = n.clone(), n.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
            if not torch.equal(lo, hi):
                raise RuntimeError("FlatGradSync: ranks disagree on the set of live parameters")
        total = sum(p.numel() for p in self.live)
        ref = self.live[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.live:
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.grad)
            p.grad = view
            off += p.numel()

    def zero(self) -> None:
        if self.flat is None:
            for p in self.all_params:
                p.grad = None
        else:
            self.flat.zero_()

    def sync(self) -> None:
        """Mean of the per-rank gradients (DDP semantics)."""
        if self.flat is None:
            self.prepare()
        w = self.world_size
        if w == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(w)

    def live_parameters(self) -> List[torch.nn.Parameter]:
        return self.live
