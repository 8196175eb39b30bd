"""gat2_lite (bond graph + atom graph + pooling; reference fragnet/model/gat/gat2_lite.py): the CPU restatement pinned
against the unmodified reference module and the committed golden vectors; the drop-in's surface; and -- on the GPU --
the CUDA path against both."""
import os
import sys

import pytest
import torch

from conftest import FP32_REL_TOL, ROOT, grad_errs, rel_err
from oracle import gat2_oracle as O, ref_import

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
GOLDEN = os.path.join(ROOT, "tests", "golden", "gat2_lite_golden.pt")
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")
GRAD_TOL = 5e-5


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN)


@pytest.fixture(scope="module")
def batch():
    from make_golden_lite import lite_batch
    return lite_batch()


def _product(gold, train=False):
    from fragnet.model.gat.gat2_lite import FragNetFineTune
    torch.manual_seed(gold["weight_seed"])
    m = FragNetFineTune(**gold["kwargs"])
    return m.train() if train else m.eval()


def test_lite_surface_and_golden_inputs(gold, batch):
    m = _product(gold)
    sd = m.state_dict()
    assert list(sd) == gold["state_keys"]
    for k, v in gold["state_checksums"].items():
        assert abs(float(sd[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, v), k
    for k, v in gold["batch_checksums"].items():
        assert abs(float(batch[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, v), k


def test_lite_oracle_against_golden_vectors(gold, batch):
    P = O.params_from_module(_product(gold))
    pred = O.lite_finetune_forward(P, batch, num_layer=gold["kwargs"]["num_layer"])
    assert rel_err(pred, gold["pred"]) <= 1e-6
    target = torch.linspace(-1.0, 1.0, pred.numel()).view_as(pred)
    torch.nn.functional.mse_loss(pred, target).backward()
    for k, g in gold["grads"].items():
        assert rel_err(P[k].grad, g) <= 1e-6, k
    for k in gold["grad_none"]:
        assert P[k].grad is None, k
    with torch.no_grad():
        enc = O.lite_fragnet_forward(O.params_from_module(_product(gold), False), batch, gold["kwargs"]["num_layer"])
    for a, b in zip(enc[:3], gold["encoder"]):
        assert rel_err(a, b) <= 1e-6


@needs_ref
def test_lite_oracle_equals_reference_module(batch):
    lite = ref_import.load_lite()
    torch.manual_seed(11)
    with ref_import.quiet():
        m = lite.FragNetFineTune(n_classes=1, num_layer=2, drop_ratio=0.0, edge_features=17, fthead="FTHead4", act="silu")
    for n, p in m.named_parameters():
        if n.endswith(".bias") and n.split(".")[-2].isdigit():
            p.data.zero_()
    m.eval()
    with ref_import.quiet():
        want = m(batch)
        layer = m.pretrain.layers[0]
        layer.return_attentions = True
        lw = layer(batch["x_atoms"], batch["edge_index"], batch["edge_attr"], batch["x_frags"], batch["atom_to_frag_ids"],
                   batch["node_features_bonds"], batch["edge_index_bonds_graph"], batch["edge_attr_bonds"])
    P = O.params_from_module(m, False)
    assert torch.equal(O.lite_finetune_forward(P, batch, num_layer=2, fthead="FTHead4", act="silu"), want)
    lg = O.lite_layer_forward(P, "pretrain.layers.0.", 4, batch["x_atoms"], batch["edge_index"], batch["edge_attr"],
                              batch["x_frags"], batch["atom_to_frag_ids"], batch["node_features_bonds"],
                              batch["edge_index_bonds_graph"], batch["edge_attr_bonds"])
    for a, b in zip(lg, lw):
        assert (a is None and b is None) or torch.equal(a, b)


@pytest.mark.gpu
def test_lite_cuda_path_against_golden_and_oracle(gold, batch):
    m = _product(gold).cuda()
    b = {k: v.cuda() for k, v in batch.items()}
    pred = m(b)
    assert rel_err(pred, gold["pred"]) <= FP32_REL_TOL
    enc = m.pretrain(b)
    assert enc[3] is None
    for a, w in zip(enc[:3], gold["encoder"]):
        assert rel_err(a, w) <= FP32_REL_TOL
    target = torch.linspace(-1.0, 1.0, pred.numel()).view_as(pred).cuda()
    torch.nn.functional.mse_loss(pred, target).backward()
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert sorted(k for k, p in m.named_parameters() if p.grad is None) == gold["grad_none"]
    P = O.params_from_module(_product(gold))
    ref = O.lite_finetune_forward(P, batch, num_layer=gold["kwargs"]["num_layer"])
    torch.nn.functional.mse_loss(ref, target.cpu()).backward()
    bad = {k: v for k, v in grad_errs([(k, g, P[k].grad) for k, g in got.items()]).items() if v > GRAD_TOL}
    assert not bad, bad
    bad = {k: v for k, v in grad_errs([(k, got[k], g) for k, g in gold["grads"].items()]).items() if v > GRAD_TOL}
    assert not bad, bad


@pytest.mark.gpu
def test_lite_layer_signature_attentions_and_training_mode(gold, batch):
    m = _product(gold).cuda()
    b = {k: v.cuda() for k, v in batch.items()}
    layer = m.pretrain.layers[0]
    layer.return_attentions = True
    with torch.no_grad():
        out = layer(b["x_atoms"], b["edge_index"], b["edge_attr"], b["x_frags"], b["atom_to_frag_ids"],
                    b["node_features_bonds"], b["edge_index_bonds_graph"], b["edge_attr_bonds"])
    assert len(out) == 8 and out[3] is None and out[5] is None and out[7] is None
    P = O.params_from_module(_product(gold), False)
    want = O.lite_layer_forward(P, "pretrain.layers.0.", 4, batch["x_atoms"], batch["edge_index"], batch["edge_attr"],
                                batch["x_frags"], batch["atom_to_frag_ids"], batch["node_features_bonds"],
                                batch["edge_index_bonds_graph"], batch["edge_attr_bonds"])
    for a, w in zip(out, want):
        assert (a is None and w is None) or rel_err(a, w) <= FP32_REL_TOL
    # training mode: dropout active, finite, seeded
    mt = _product(gold, train=True).cuda()
    torch.manual_seed(5)          # the library's dropout stream: (torch seed, torch's CUDA generator offset)
    y1 = mt(b)
    torch.manual_seed(5)
    y2 = mt(b)
    assert not torch.equal(y1, mt(b))      # a further call draws fresh masks
    assert torch.isfinite(y1).all() and torch.equal(y1, y2)
    y1.sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in mt.parameters() if p.grad is not None)
