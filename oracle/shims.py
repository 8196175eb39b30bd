"""Pure-torch stand-ins for the third-party ops on the reference's hot path.  TEST INFRASTRUCTURE.

Semantics restated from the packages' published behaviour:

* ``torch_scatter.scatter_add(src, index, dim=0)`` (alias of ``scatter_sum``): ``index`` is
  broadcast to ``src``; the output has ``index.max()+1`` rows when ``dim_size`` is not given
  (the reference never passes it: gat2.py:162,165,216,219,234,265,268,309,312,820,821).
* ``torch_scatter.scatter_softmax(src, index, dim=0)``: subtract the per-group max, ``exp``,
  divide by the per-group sum; no epsilon (gat2.py:153,210,257,303).
* ``torch_geometric.utils.add_self_loops(edge_index)`` (2.6.1) with ``num_nodes=None``:
  appends ``arange(edge_index.max()+1)`` as both rows and returns ``(edge_index, None)``
  (gat2.py:179).
"""
from __future__ import annotations

import sys
import types

import torch


def _rows(index: torch.Tensor, dim_size):
    if dim_size is not None:
        return int(dim_size)
    return int(index.max()) + 1 if index.numel() else 0


def _expand(index: torch.Tensor, src: torch.Tensor, dim: int) -> torch.Tensor:
    if index.dim() == src.dim():
        return index
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    dim = dim % src.dim()
    idx = _expand(index, src, dim)
    if out is None:
        size = list(src.shape)
        size[dim] = _rows(index, dim_size)
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, idx, src)


scatter_sum = scatter_add


def scatter_max(src, index, dim=0, dim_size=None):
    dim = dim % src.dim()
    idx = _expand(index, src, dim)
    size = list(src.shape)
    size[dim] = _rows(index, dim_size)
    out = torch.full(size, float("-inf"), dtype=src.dtype, device=src.device)
    out = out.scatter_reduce(dim, idx, src, reduce="amax", include_self=True)
    return out, None


def scatter_softmax(src, index, dim=0, dim_size=None):
    dim = dim % src.dim()
    idx = _expand(index, src, dim)
    group_max, _ = scatter_max(src, index, dim, dim_size)
    shifted = (src - group_max.gather(dim, idx)).exp()
    group_sum = scatter_add(shifted, index, dim, dim_size=dim_size)
    return shifted / group_sum.gather(dim, idx)


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    n = int(num_nodes) if num_nodes is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device).repeat(2, 1)
    return torch.cat([edge_index, loops], dim=1), edge_attr


def degree(index, num_nodes=None, dtype=None):
    n = int(num_nodes) if num_nodes is not None else (int(index.max()) + 1 if index.numel() else 0)
    out = torch.zeros(n, dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.scatter_add_(0, index, torch.ones_like(index, dtype=out.dtype))


class _Unavailable(torch.nn.Module):
    """Placeholder for torch_geometric layers the hot path never constructs."""

    def __init__(self, *a, **k):
        super().__init__()
        raise RuntimeError("torch_geometric layer stub: not part of the GAT2 hot path")


def install() -> None:
    """Register the stand-ins under the third-party module names (idempotent)."""
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")
        ts.scatter_add, ts.scatter_sum = scatter_add, scatter_sum
        ts.scatter_softmax, ts.scatter_max = scatter_softmax, scatter_max
        sys.modules["torch_scatter"] = ts
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgu = types.ModuleType("torch_geometric.utils")
        tgu.add_self_loops, tgu.degree = add_self_loops, degree
        tgn = types.ModuleType("torch_geometric.nn")
        tgn.TransformerConv = _Unavailable
        tgnn = types.ModuleType("torch_geometric.nn.norm")
        tgnn.BatchNorm = _Unavailable
        tgn.norm = tgnn
        tg.utils, tg.nn = tgu, tgn
        sys.modules.update({"torch_geometric": tg, "torch_geometric.utils": tgu,
                            "torch_geometric.nn": tgn, "torch_geometric.nn.norm": tgnn})
