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


def test_neg_mean_auc_is_the_reference_quantity():
    """``validate_clsf_bce`` / ``test_clsf_bce`` return MINUS the mean ROC-AUC over the scorable label columns, on raw
    model outputs (reference utils.py:477-486, :532-544); EarlyStopping minimises it (finetune_gat2.py:268-274)."""
    import numpy as np
    from sklearn.metrics import roc_auc_score

    from fragnet.train.utils import TrainerFineTune
    rng = np.random.default_rng(0)
    target = rng.integers(0, 2, size=(40, 4)).astype(np.float32)
    target[rng.random((40, 4)) < 0.2] = -1.0          # missing labels
    target[:, 3] = 1.0                                # a column without negatives is skipped
    pred = rng.normal(size=(40, 4)).astype(np.float32) + target
    want = []
    for i in range(3):
        ok = target[:, i] > -0.5
        want.append(roc_auc_score(target[ok, i], pred[ok, i]))
    got = TrainerFineTune._neg_mean_auc(target, pred)
    assert got < 0 and abs(got + sum(want) / 3) < 1e-12


@pytest.mark.gpu
def test_clsf_bce_train_validate_test_follow_the_reference():
    """One epoch of ``train_clsf_bce`` + ``validate`` + ``test`` against the reference's loops written out by hand
    (utils.py:406-438, :460-487, :514-544): same mean loss, same weights, ``(-roc_auc, targets, raw logits)``."""
    import numpy as np
    from sklearn.metrics import roc_auc_score
    from torch.utils.data import DataLoader

    from fragnet.dataset.data import collate_fn
    from fragnet.model.gat.gat2 import FragNetFineTune
    from fragnet.train.utils import TrainerFineTune
    from fragnet_b200 import synth
    ds = synth.make_dataset("esol", 48, seed=21, with_pretrain_targets=False)
    gen = torch.Generator().manual_seed(4)
    for m in ds:
        y = torch.randint(0, 2, (3,), generator=gen).float()
        y[torch.rand(3, generator=gen) < 0.15] = -1.0
        m.y = y
    loader = DataLoader(ds, batch_size=16, shuffle=False, drop_last=True, collate_fn=collate_fn)
    torch.manual_seed(3)
    m1 = FragNetFineTune(n_classes=3, num_layer=2, drop_ratio=0.0, h1=32, h2=32, h3=32, h4=32, act="relu").cuda()
    m2 = copy.deepcopy(m1)
    o1, o2 = torch.optim.Adam(m1.parameters(), lr=1e-3), torch.optim.Adam(m2.parameters(), lr=1e-3)
    tr = TrainerFineTune(target_type="clsf")
    got = tr.train(m1, loader, o1, None, "cuda", loader)
    bce = torch.nn.BCEWithLogitsLoss(reduction="none")
    m2.train()
    total = 0.0
    for batch in loader:
        batch = {k: v.to("cuda") for k, v in batch.items()}
        o2.zero_grad()
        out = m2(batch)
        labels = batch["y"].view(out.shape)
        valid = labels > -0.5
        loss_mat = torch.where(valid, bce(out, labels), torch.zeros_like(out))
        loss = torch.sum(loss_mat) / torch.sum(valid)
        loss.backward()
        total += loss.item()
        o2.step()
    assert abs(got - total / len(loader.dataset)) <= 1e-6
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    m2.eval()
    with torch.no_grad():
        t, p = [], []
        for batch in loader:
            batch = {k: v.to("cuda") for k, v in batch.items()}
            out = m2(batch)
            t.append(batch["y"].view(out.shape))
            p.append(out)
    t, p = torch.cat(t).cpu().numpy(), torch.cat(p).cpu().numpy()
    rocs = [roc_auc_score(t[t[:, i] > -0.5, i], p[t[:, i] > -0.5, i]) for i in range(3)
            if np.sum(t[:, i] == 1) > 0 and np.sum(t[:, i] == 0) > 0]
    want = -sum(rocs) / len(rocs)
    v = tr.validate(m1, loader, "cuda")
    s_, t1, p1 = tr.test(m1, loader, "cuda")
    assert v == s_ and abs(v - want) < 1e-9 and v < 0
    assert np.array_equal(t1, t) and np.array_equal(p1, p)          # raw logits, not probabilities


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
