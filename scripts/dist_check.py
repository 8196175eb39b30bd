#!/usr/bin/env python
"""Multi-GPU check of the fused gradient exchange (csrc/dist.cu), to be launched with torchrun on N >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py

Twin models per rank, different batches per rank: one trained through ``fnb_allreduce_adam_step`` (peer-memory reads +
Adam in one kernel), one through ``dist.all_reduce`` + ``fnb_adam_step``.  Asserts (1) the fused path is active, (2)
after every step the parameters of all ranks are BITWISE identical in the fused path, (3) the two paths agree to fp32
rounding, (4) reports the time per step of both."""
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    from fragnet.model.gat.gat2_pretrain import FragNetPreTrain
    from fragnet_b200.train.fused import FusedPretrainStep
    torch.manual_seed(1234 + rank)          # DIFFERENT initial weights per rank: the constructor must broadcast rank 0's
    m1 = FragNetPreTrain(**bench.PT_KW).to(dev).train()
    m2 = copy.deepcopy(m1)
    batches = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_batches("unimol", 256, 3, 128, seed=50 + rank)]
    f1 = FusedPretrainStep(m1, lr=1e-3)
    os.environ["FNB_FUSED_ALLREDUCE"] = "0"
    f2 = FusedPretrainStep(m2, lr=1e-3)
    os.environ["FNB_FUSED_ALLREDUCE"] = "1"
    assert f1._peers is not None, "fused exchange not active"
    assert f2._peers is None
    f2.flat_p.copy_(f1.flat_p)
    torch.manual_seed(7)
    for i in range(4):
        torch.manual_seed(100 + i)           # same dropout stream for the twins
        l1 = f1.step(batches[i % 3])
        torch.manual_seed(100 + i)
        l2 = f2.step(batches[i % 3])
        torch.cuda.synchronize()
        gathered = [torch.empty_like(f1.flat_p) for _ in range(world)]
        dist.all_gather(gathered, f1.flat_p)
        assert all(torch.equal(gathered[0], g) for g in gathered), f"step {i}: ranks diverged in the fused path"
        err = float((f1.flat_p - f2.flat_p).abs().max() / f2.flat_p.abs().max())
        assert err < 2e-5, (i, err)
        assert torch.isfinite(l1) and abs(float(l1) - float(l2)) <= 1e-4 * abs(float(l2)) + 1e-6, (float(l1), float(l2))
    # with dropout: the same seed gives the twins the same masks (the counter restarts with torch.manual_seed)
    # timing
    big = [{k: v.to(dev) for k, v in b.items()} for b in bench.make_batches("unimol", 1024, 2, 256, seed=70 + rank)]
    out = {}
    for name, f in (("fused", f1), ("nccl", f2)):
        for i in range(5):
            f.step(big[i % 2])
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for i in range(100):
            f.step(big[i % 2])
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 100], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[name] = float(ms)
    if rank == 0:
        print(f"dist_check OK: world {world}, rank-identical parameters, fused vs nccl max rel diff < 2e-5; "
              f"ms/step fused {out['fused']:.4f} nccl {out['nccl']:.4f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
