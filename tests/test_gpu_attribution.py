"""Batched mask attribution: one forward over R replicas of a molecule, replica r carrying mask r, against (a) the
reference's procedure -- one batch-1 forward per mask with the scalar mask attributes (viz.py:960-984, 1026-1050,
1145-1169) -- through the same CUDA path and (b) the CPU restatement of the masked reference layer."""
import pytest
import torch

from conftest import FP32_REL_TOL, rel_err

pytestmark = pytest.mark.gpu


def _setup():
    from fragnet.model.gat.gat2 import FragNetFineTune
    from fragnet_b200 import synth
    torch.manual_seed(21)
    m = FragNetFineTune(n_classes=1, num_layer=3, drop_ratio=0.1, h1=64, h2=64, h3=64, h4=64, act="relu").cuda().eval()
    mol = synth.make_dataset("esol", 3, seed=9, with_pretrain_targets=False)[2]
    return m, mol


def test_batched_masks_equal_one_forward_per_mask():
    from fragnet_b200.dataset.data import collate_fn
    from fragnet_b200.vizualize.attribution import attributions, mask_predictions
    m, mol = _setup()
    got = mask_predictions(m, mol)
    b1 = {k: v.cuda() for k, v in collate_fn([mol]).items()}
    na, ea, nfb = mol.x_atoms.shape[0], mol.edge_index.shape[1], mol.node_feautures_fbondg.shape[0]
    assert got["atom"].shape[0] == na and got["bond"].shape[0] == ea // 2 and got["fbond"].shape[0] == nfb // 2

    def single(attr, value):
        for l in m.pretrain.layers:
            setattr(l, attr, value)
        with torch.no_grad():
            y = m(b1)[0]
        for l in m.pretrain.layers:
            setattr(l, attr, None)
        return y

    with torch.no_grad():
        assert rel_err(got["pred_no_mask"], m(b1)[0]) <= 1e-6
    scale = float(got["pred_no_mask"].abs().max())
    for i in range(na):
        assert float((got["atom"][i] - single("atom_mask_individual", i)).abs().max()) <= 1e-6 * max(scale, 1.0), i
    for j in range(ea // 2):
        assert float((got["bond"][j] - single("bond_mask", 2 * j)).abs().max()) <= 1e-6 * max(scale, 1.0), j
    for k in range(nfb // 2):
        assert float((got["fbond"][k] - single("frag_bond_mask", k)).abs().max()) <= 1e-6 * max(scale, 1.0), k
    # masks are removed again, chunking gives the same values, and masking changes the prediction
    assert all(l.bond_mask is None and l.atom_mask_individual is None and l.frag_bond_mask is None
               for l in m.pretrain.layers)
    chunked = mask_predictions(m, mol, max_replicas=7)
    for k in got:
        assert rel_err(chunked[k], got[k]) <= 1e-6, k
    att = attributions(m, mol)
    assert float(att["atom"].abs().max()) > 0 and float(att["bond"].abs().max()) > 0


def test_batched_masks_equal_cpu_restatement():
    from fragnet_b200.dataset.data import collate_fn
    from fragnet_b200.vizualize.attribution import mask_predictions
    from oracle import gat2_oracle as O
    m, mol = _setup()
    got = mask_predictions(m, mol)
    P = O.params_from_module(m, False)
    hb = collate_fn([mol])

    def ref(**masks):
        with torch.no_grad():
            xa, xf, _, _ = O.fragnet_forward(P, hb, 3, masks=masks)
            return O.fthead_forward(P, O.readout(xa, xf, hb))[0]

    assert rel_err(got["pred_no_mask"], ref()) <= FP32_REL_TOL
    for i in (0, mol.x_atoms.shape[0] - 1):
        assert rel_err(got["atom"][i], ref(atom_mask_individual=i)) <= FP32_REL_TOL
    for j in (0, mol.edge_index.shape[1] // 2 - 1):
        assert rel_err(got["bond"][j], ref(bond_mask=2 * j)) <= FP32_REL_TOL
    for k in range(mol.node_feautures_fbondg.shape[0] // 2):
        assert rel_err(got["fbond"][k], ref(frag_bond_mask=k)) <= FP32_REL_TOL
