"""Finetuning loops (reference fragnet/train/utils.py): EarlyStopping on the CPU; on the GPU an epoch of
``TrainerFineTune.train_regr`` against the reference's loop written out by hand (blocking copies, per-step
``loss.item()``) on identically initialised models."""
import copy

import pytest
import torch


def test_early_stopping_counts_and_checkpoints(tmp_path, capsys):
    from fragnet.train.utils import EarlyStopping
    m = torch.nn.Linear(2, 1)
    es = EarlyStopping(patience=2, chkpoint_name=str(tmp_path / "best.pt"))
    es(1.0, m)
    assert (tmp_path / "best.pt").exists() and es.val_loss_min == 1.0 and not es.early_stop
    es(1.5, m)
    assert es.counter == 1 and not es.early_stop
    es(0.5, m)
    assert es.counter == 0 and es.val_loss_min == 0.5
    es(0.6, m)
    es(0.7, m)
    assert es.early_stop
    capsys.readouterr()


@pytest.mark.gpu
def test_train_regr_epoch_equals_the_reference_loop():
    from torch.utils.data import DataLoader

    from fragnet.dataset.data import collate_fn
    from fragnet.model.gat.gat2 import FragNetFineTune
    from fragnet.train.utils import TrainerFineTune, test_fn
    from fragnet_b200 import synth
    ds = synth.make_dataset("esol", 40, seed=12, with_pretrain_targets=False)
    loader = DataLoader(ds, batch_size=16, shuffle=False, drop_last=True, collate_fn=collate_fn)
    val = DataLoader(ds[:8], batch_size=8, shuffle=False, collate_fn=collate_fn)
    torch.manual_seed(2)
    m1 = FragNetFineTune(n_classes=1, num_layer=2, drop_ratio=0.0, h1=32, h2=32, h3=32, h4=32, act="relu").cuda()
    m2 = copy.deepcopy(m1)
    o1 = torch.optim.Adam(m1.parameters(), lr=1e-3)
    o2 = torch.optim.Adam(m2.parameters(), lr=1e-3)
    tr = TrainerFineTune(target_type="regr")
    got = tr.train(m1, loader, o1, None, "cuda", val)
    # the reference's loop, utils.py:330-344
    m2.train()
    total = 0.0
    for batch in loader:
        for k in batch:
            batch[k] = batch[k].to("cuda")
        o2.zero_grad()
        loss = torch.nn.MSELoss()(m2(batch).view(-1), batch["y"])
        loss.backward()
        total += loss.item()
        o2.step()
    want = total / len(loader.dataset)
    assert abs(got - want) <= 1e-6 * max(1.0, abs(want))
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    v = tr.validate(m1, val, "cuda")
    mse, t, p = tr.test(m1, val, "cuda")
    assert abs(v * len(val.dataset) - mse) <= 1e-5 * max(1.0, mse) and t.shape == p.shape == (8,)
    assert test_fn(val, m1, "cuda")[0] == mse


def test_error_behaviour_without_a_gpu():
    """The product path has no CPU fallback: host-side entry points fail loudly instead of computing on the CPU."""
    from fragnet.train.utils import TrainerFineTune
    from fragnet_b200 import synth
    from fragnet_b200.dataset.arena import MoleculeArena
    from fragnet_b200.dataset.prefetch import DevicePrefetcher
    with pytest.raises(NotImplementedError):
        TrainerFineTune(target_type="clsf_ms")
    with pytest.raises(ValueError):
        MoleculeArena(synth.make_dataset("esol", 2, seed=0), "cpu")
    with pytest.raises(ValueError):
        DevicePrefetcher([], "cpu")
