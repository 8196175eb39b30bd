"""world_size-2 gloo test of the data-parallel gradient exchange (host logic only; runs on CPU)."""
import os
import socket

import pytest

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.live_a = torch.nn.Linear(4, 3)
        self.dead = torch.nn.Linear(4, 3)          # never used: grad stays None, like 87 FragNet tensors
        self.live_b = torch.nn.Parameter(torch.ones(3))

    def forward(self, x):
        return (self.live_a(x) * self.live_b).sum()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fragnet_b200.dist import FlatGradSync
    torch.manual_seed(0)
    model = _Toy()
    sync = FlatGradSync(model.parameters())
    xs = torch.arange(2 * 5 * 4, dtype=torch.float32).view(2, 5, 4) / 10
    for step in range(2):
        sync.zero()
        model(xs[rank] + step).backward()
        sync.sync()
    grads = {k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()}
    if rank == 0:
        torch.save(grads, out)
    dist.destroy_process_group()


def test_flat_grad_sync_equals_mean_of_rank_grads(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = _Toy()
    xs = torch.arange(2 * 5 * 4, dtype=torch.float32).view(2, 5, 4) / 10
    want = None
    for r in range(2):
        model.zero_grad()
        model(xs[r] + 1).backward()                 # the second step's inputs
        g = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        want = g if want is None else {k: want[k] + g[k] for k in g}
    for k in want:
        assert torch.allclose(got[k], want[k] / 2, atol=1e-6), k
    assert got["dead.weight"] is None and got["dead.bias"] is None


def test_flat_grad_sync_without_gradients_says_so():
    from fragnet_b200.dist import FlatGradSync
    with pytest.raises(RuntimeError, match="no parameter has a gradient"):
        FlatGradSync(_Toy().parameters()).sync()


def test_epoch_batches_shard_every_step_over_the_ranks():
    """Host-side batch schedule of ArenaLoader: world 1 = DataLoader semantics; world W = disjoint equal slices of
    every global batch, the same number of steps on every rank, identical shuffles from identical generators."""
    import numpy as np
    from fragnet_b200.dataset.arena import epoch_batches
    n, bs = 103, 8
    one = epoch_batches(n, bs)
    assert [len(b) for b in one] == [8] * 12 + [7] and np.array_equal(np.concatenate(one), np.arange(n))
    assert len(epoch_batches(n, bs, drop_last=True)) == 12
    for drop in (False, True):
        per_rank = [epoch_batches(n, bs, shuffle=True, drop_last=drop, generator=torch.Generator().manual_seed(5),
                                  rank=r, world=4) for r in range(4)]
        steps = {len(b) for b in per_rank}
        assert len(steps) == 1                                   # every rank runs the same number of steps
        for step in range(steps.pop()):
            sizes = {len(per_rank[r][step]) for r in range(4)}
            assert len(sizes) == 1                               # ... with equal shares
        seen = np.concatenate([np.concatenate(b) for b in per_rank])
        assert len(np.unique(seen)) == len(seen)                 # disjoint
        assert len(seen) == (96 if drop else 100)                # 3 global batches of 32 (+ 4 x 1 of the tail of 7)
    with pytest.raises(ValueError):
        epoch_batches(10, 2, rank=2, world=2)
