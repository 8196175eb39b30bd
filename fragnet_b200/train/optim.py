"""Flat-buffer Adam for the GAT2 training step.

The reference trains with ``torch.optim.Adam(model.parameters(), lr)`` (fragnet/train/pretrain/pretrain_gat2.py:165,
finetune_gat2.py:251).  That optimizer still works unchanged on the drop-in modules.  ``FlatAdam`` applies the same
update (torch's Adam, no amsgrad) to the parameters that actually receive gradients, but through ONE flat fp32
buffer: the live parameters become views into it, their gradients are packed with one multi-tensor copy, and the
update is one ``fnb_adam_step`` launch -- instead of per-step Python bookkeeping over ~80 tensors.  Parameters keep
their identity, names and shapes, so ``state_dict()`` / ``load_state_dict()`` are unaffected.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch

from .. import _abi


ALIGN = 64     # floats: every tensor of a flat buffer starts on a 256-byte boundary (TMA operands need 16 bytes)


def flat_views(params, device, dtype=torch.float32, storage=None):
    """One zero-filled flat buffer with an aligned slot per tensor of ``params``; returns (flat, views).  ``storage``:
    an existing flat buffer of that layout to take the views of instead of allocating one."""
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
    if storage is not None and storage.numel() != total:
        raise ValueError("flat_views: storage does not match the layout of params")
    flat = storage if storage is not None else torch.zeros(total, dtype=dtype, device=device)
    return flat, [flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, params)]


class FlatAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        ref = self.params[0]
        if any(p.dtype != torch.float32 or p.device != ref.device or not p.is_cuda for p in self.params):
            raise ValueError("FlatAdam needs fp32 CUDA parameters on one device")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.flat_p, views = flat_views(self.params, ref.device)
        with torch.no_grad():
            torch._foreach_copy_(views, [p.data for p in self.params])
            for p, v in zip(self.params, views):
                p.data = v                     # the parameter now lives in the flat buffer
        self.flat_g, self._g_views = flat_views(self.params, ref.device)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.t = 0

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self, flat_grad: Optional[torch.Tensor] = None) -> None:
        """``flat_grad``: gradients already packed in this optimizer's parameter order (e.g. the all-reduced buffer of
        ``FlatGradSync`` built from the same parameter list); otherwise they are packed here from ``p.grad``."""
        if flat_grad is None:
            torch._foreach_copy_(self._g_views, [p.grad for p in self.params])
            flat_grad = self.flat_g
        elif flat_grad.numel() != self.flat_p.numel():
            raise ValueError("FlatAdam.step: flat_grad does not match the parameter layout")
        self.t += 1
        lib = _abi.load()
        rc = lib.fnb_adam_step(C.c_void_p(self.flat_p.data_ptr()), C.c_void_p(flat_grad.data_ptr()),
                               C.c_void_p(self.exp_avg.data_ptr()), C.c_void_p(self.exp_avg_sq.data_ptr()),
                               self.flat_p.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                               self.t, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _abi.check(rc, "adam_step")
