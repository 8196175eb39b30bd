"""Data-parallel gradient exchange for GAT2 training: one process per GPU, one flat all-reduce.

The reference's only multi-GPU code is Lightning Fabric DDP (fragnet/train/finetune/finetune_gat2_pl.py:230,
``fabric.backward``): mean-reduced gradients across ranks.  Molecules are independent (block-diagonal
batches, fragnet/dataset/data.py:877-948) so the forward/backward needs no communication; the only
exchange is the gradient of the LIVE parameters (87 of FragNet's tensors never receive a gradient,
SURVEY.md fact 7, and are skipped exactly as DDP/Adam skip ``grad is None``).

Per step: ``zero()`` drops the gradients (``grad = None``: the next backward then hands its freshly written
gradient tensors straight to the parameters, no accumulate kernels), ``sync()`` packs the live gradients
into one flat fp32 buffer (one multi-tensor copy), all-reduces it once (NCCL over NVLink/NVSwitch; gloo in
the CPU tests), scales by 1/world and rebinds every ``.grad`` to its slice of the buffer.  With one rank
``sync()`` does nothing.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class FlatGradSync:
    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.all_params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.live: List[torch.nn.Parameter] = []
        self.flat = None
        self._views: List[torch.Tensor] = []

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def prepare(self) -> None:
        """Find the parameters that received a gradient (a property of the model, not of the batch; checked to
        agree across ranks) and lay out the flat buffer."""
        self.live = [p for p in self.all_params if p.grad is not None]
        if not self.live:
            raise RuntimeError("FlatGradSync: no parameter has a gradient yet -- call backward() before sync()")
        if self.world_size > 1:
            n = torch.tensor([len(self.live), sum(p.numel() for p in self.live)], dtype=torch.int64,
                             device=self.live[0].device)
            lo, hi = n.clone(), n.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
            if not torch.equal(lo, hi):
                raise RuntimeError("FlatGradSync: ranks disagree on the set of live parameters")
        ref = self.live[0]
        from .train.optim import flat_views
        self.flat, self._views = flat_views(self.live, ref.device, ref.dtype)   # same layout as FlatAdam's buffers

    def zero(self) -> None:
        for p in self.all_params:
            p.grad = None

    def sync(self) -> None:
        """Mean of the per-rank gradients (DDP semantics)."""
        if self.flat is None:
            self.prepare()
        w = self.world_size
        if w == 1:
            return
        torch._foreach_copy_(self._views, [p.grad for p in self.live])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(w)
        for p, v in zip(self.live, self._views):
            p.grad = v

    def live_parameters(self) -> List[torch.nn.Parameter]:
        return self.live
