"""Host -> device staging of batch dicts, overlapped with compute.

The reference's loops move every tensor of a batch with a blocking ``batch[k].to(device)`` from pageable memory
right before the forward (fragnet/train/pretrain/pretrain_utils.py:13-14, train/utils.py:335-336): ~40 MB per
1024-molecule batch (19 tensors, int64 indices, one-hot fp32 features), i.e. ~1.2 ms of PCIe time that the GPU spends
idle.  ``DevicePrefetcher`` wraps any iterable of batch dicts (a ``DataLoader`` with ``collate_fn_pt``) and issues
the copies of the NEXT batches on a side stream from pinned memory while the current one is being consumed.
"""
from __future__ import annotations

import collections
from typing import Dict, Iterable, Iterator

import torch


class DevicePrefetcher:
    """``for batch in DevicePrefetcher(loader, device): ...`` yields device-resident batch dicts.

    ``depth`` batches are in flight (2 = double buffering).  Tensors that are not pinned are pinned first (use
    ``DataLoader(pin_memory=True)`` to do that in the loader's workers)."""

    def __init__(self, batches: Iterable[Dict[str, torch.Tensor]], device, depth: int = 2):
        self.batches, self.device, self.depth = batches, torch.device(device), max(1, int(depth))
        if self.device.type != "cuda":
            raise ValueError("DevicePrefetcher stages batches onto a CUDA device")
        self._copy_stream = torch.cuda.Stream(self.device)

    def _stage(self, host: Dict[str, torch.Tensor]):
        with torch.cuda.stream(self._copy_stream):
            dev = {}
            for k, v in host.items():
                if isinstance(v, torch.Tensor):
                    if not v.is_cuda and not v.is_pinned():
                        v = v.pin_memory()
                    dev[k] = v.to(self.device, non_blocking=True)
                else:
                    dev[k] = v
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return dev, ev, host          # the (pinned) host tensors stay alive until the copy has been waited for

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        it = iter(self.batches)
        queue = collections.deque()
        done = False
        while True:
            while not done and len(queue) < self.depth:
                try:
                    queue.append(self._stage(next(it)))
                except StopIteration:
                    done = True
            if not queue:
                return
            dev, ev, _host = queue.popleft()
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for v in dev.values():
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    v.record_stream(cur)          # allocated on the copy stream, consumed on the compute stream
            yield dev
