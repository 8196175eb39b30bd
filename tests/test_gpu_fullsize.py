"""Size-independent properties at BASELINE.json's full sizes (batch 4 096 UniMol-shaped molecules; the stress shape):
the CPU restatement cannot run these sizes in seconds, so parity is checked through invariants of the domain --
CSR sortedness / permutation / inverse properties (bit-exact), softmax row sums of the returned attention weights,
independence of a molecule's outputs from the rest of its batch, arena assembly == host collate, and
launch-to-launch determinism of the training step."""
import numpy as np
import pytest
import torch

from conftest import FP32_REL_TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    from fragnet_b200 import synth
    from fragnet_b200.dataset.data import collate_fn_pt
    pool = synth.make_dataset("unimol", 512, seed=40)
    rng = np.random.default_rng(40)
    ids = rng.integers(0, len(pool), size=4096)
    ids[:16] = np.arange(16)                      # known molecules at both ends of the batch
    ids[-16:] = np.arange(16, 32)
    hb = collate_fn_pt([pool[int(i)] for i in ids])
    return pool, ids, hb, {k: v.cuda() for k, v in hb.items()}


def _plan(b):
    from fragnet_b200 import ops
    return ops.build_layer_plan(b["edge_index"], b["frag_index"], b["atom_to_frag_ids"], b["edge_index_bonds_graph"],
                                b["edge_attr_bonds"], b["edge_index_fbonds"], b["edge_attr_fbonds"],
                                b["x_atoms"].shape[0], b["x_frags"].shape[0], b["node_features_bonds"].shape[0],
                                b["node_features_fbonds"].shape[0], "cuda", b["batch"], b["frag_batch"])


def test_csr_properties_at_full_size(big):
    _, _, _, b = big
    plan = _plan(b)
    assert int(plan.status.item()) == 0
    na = b["x_atoms"].shape[0]
    loops = torch.arange(na, device="cuda")
    graphs = {   # name -> (destination, source) of every input edge, in input order
        "bond": (b["edge_index_bonds_graph"][0], b["edge_index_bonds_graph"][1]),
        "fbond": (b["edge_index_fbonds"][0], b["edge_index_fbonds"][1]),
        "atom": (torch.cat((b["edge_index"][1], loops)), torch.cat((b["edge_index"][0], loops))),
        "frag": (b["frag_index"][1], b["frag_index"][0]),
    }
    for name, (dst, src) in graphs.items():
        g = getattr(plan, name)
        n, e = g.n_nodes, g.n_edges
        assert e == dst.numel()
        rowptr, row, col, eid = g.rowptr.long(), g.row.long(), g.col.long(), g.eid.long()
        assert int(rowptr[0]) == 0 and int(rowptr[-1]) == e and bool((rowptr[1:] >= rowptr[:-1]).all())
        assert torch.equal(rowptr[1:] - rowptr[:-1], torch.bincount(dst, minlength=n))       # segment sizes
        assert bool((row[1:] >= row[:-1]).all())                                             # destination-sorted
        assert torch.equal(torch.sort(eid).values, torch.arange(e, device="cuda"))           # eid is a permutation
        assert torch.equal(dst[eid], row) and torch.equal(src[eid], col)                     # ... of the input edges
        seg_start = rowptr[row]
        same = row[1:] == row[:-1]
        assert bool((eid[1:][same] > eid[:-1][same]).all())                                  # stable inside a segment
        assert torch.equal(g.slot_of_eid.long()[eid], torch.arange(e, device="cuda"))        # inverse permutation
        rrowptr, rslot, rdst = g.rrowptr.long(), g.rslot.long(), g.rdst.long()
        assert torch.equal(rrowptr[1:] - rrowptr[:-1], torch.bincount(src, minlength=n))
        assert torch.equal(torch.sort(rslot).values, torch.arange(e, device="cuda"))
        assert torch.equal(row[rslot], rdst)
        rsrc = col[rslot]
        assert bool((rsrc[1:] >= rsrc[:-1]).all())                                           # source-sorted
        del seg_start
    pool_g = plan.pool
    assert torch.equal((pool_g.rowptr[1:] - pool_g.rowptr[:-1]).long(),
                       torch.bincount(b["atom_to_frag_ids"], minlength=plan.n_frags))
    assert torch.equal(b["atom_to_frag_ids"][pool_g.col.long()],
                       torch.repeat_interleave(torch.arange(plan.n_frags, device="cuda"),
                                               (pool_g.rowptr[1:] - pool_g.rowptr[:-1]).long()))


def test_arena_assembly_equals_collate_at_full_size(big):
    from fragnet_b200.dataset.arena import MoleculeArena
    pool, ids, hb, _ = big
    got = MoleculeArena(pool, "cuda").batch(ids)
    assert list(got) == list(hb)
    for k in hb:
        assert torch.equal(got[k].cpu(), hb[k]), k


def test_attention_row_sums_and_batch_independence_at_full_size(big):
    """Every destination's softmax sums to one, so the by-source attention sums add up to the number of destinations
    with at least one incoming edge (per head); and a molecule's outputs do not depend on its batch mates."""
    from fragnet.vizualize.model import FragNetFineTuneViz
    from fragnet_b200.dataset.data import collate_fn_pt
    pool, ids, _, b = big
    torch.manual_seed(9)
    m = FragNetFineTuneViz(n_classes=1, num_layer=4, drop_ratio=0.1, edge_features=17, h1=64, h2=64, h3=64, h4=64,
                           act="relu").cuda().eval()
    with torch.no_grad():
        pred, a_atoms, a_frags, a_bonds, a_fbonds = m(b)
    assert pred.shape == (4096, 1) and bool(torch.isfinite(pred).all())
    expect = {
        "atoms": b["x_atoms"].shape[0],                                   # self loops: every atom is a destination
        "bonds": int(torch.unique(b["edge_index_bonds_graph"][0]).numel()),
        "frags": int(torch.unique(b["frag_index"][1]).numel()),
        "fbonds": int(torch.unique(b["edge_index_fbonds"][0]).numel()),
    }
    for name, t in (("atoms", a_atoms), ("bonds", a_bonds), ("frags", a_frags), ("fbonds", a_fbonds)):
        sums = t.double().sum(0)
        assert float((sums - expect[name]).abs().max()) <= 1e-5 * expect[name], (name, sums.tolist(), expect[name])
        assert float(t.min()) >= 0.0
    for sl, mols in ((slice(0, 16), range(16)), (slice(4080, 4096), range(16, 32))):
        small = {k: v.cuda() for k, v in collate_fn_pt([pool[i] for i in mols]).items()}
        with torch.no_grad():
            want = m(small)
        assert rel_err(pred[sl], want[0]) <= FP32_REL_TOL
    # first / last molecule's atom attention rows as well
    n0 = pool[0].x_atoms.shape[0]
    with torch.no_grad():
        one = m({k: v.cuda() for k, v in collate_fn_pt([pool[0]]).items()})
    assert rel_err(a_atoms[:n0], one[1]) <= FP32_REL_TOL


def test_training_step_is_deterministic_at_full_size(big):
    import copy

    from fragnet.model.gat.gat2_pretrain import FragNetPreTrain
    from fragnet_b200.train.fused import FusedPretrainStep
    _, _, _, b = big
    torch.manual_seed(3)
    m1 = FragNetPreTrain(num_layer=4, drop_ratio=0.0, edge_features=17).cuda().train()
    m2 = copy.deepcopy(m1)
    s1, s2 = FusedPretrainStep(m1, lr=1e-4), FusedPretrainStep(m2, lr=1e-4)
    l1 = [float(s1.step(b)) for _ in range(2)]
    l2 = [float(s2.step(b)) for _ in range(2)]
    assert l1 == l2 and all(np.isfinite(l1)) and l1[1] != l1[0]
    for (k, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.equal(p, q), k


def test_stress_shape_batch_independence():
    """BASELINE configs[4]: ~100 atoms, ~20 k fragment-connection edges per molecule, in-degree up to ~114."""
    from fragnet.model.gat.gat2 import FragNetFineTune
    from fragnet_b200 import synth
    from fragnet_b200.dataset.data import collate_fn
    pool = synth.make_dataset("stress", 24, seed=8, with_pretrain_targets=False)
    torch.manual_seed(4)
    m = FragNetFineTune(n_classes=1, num_layer=4, drop_ratio=0.1, h1=64, h2=64, h3=64, h4=64, act="relu").cuda().eval()
    with torch.no_grad():
        full = m({k: v.cuda() for k, v in collate_fn(pool).items()})
        head = m({k: v.cuda() for k, v in collate_fn(pool[:3]).items()})
        tail = m({k: v.cuda() for k, v in collate_fn(pool[-2:]).items()})
    assert rel_err(full[:3], head) <= FP32_REL_TOL and rel_err(full[-2:], tail) <= FP32_REL_TOL


def test_forward_parity_with_the_oracle_at_the_bench_batch_size():
    """BASELINE's per-GPU batch (1 024 UniMol-shaped molecules, the batch bench.py times): the CUDA forward of the
    pretraining model in eval mode against the CPU oracle on the very same batch -- every prediction tensor, the loss and
    the four encoder outputs within 1e-5 (the oracle needs a few seconds per forward at this size, so this runs once)."""
    import bench
    from fragnet.model.gat.gat2_pretrain import FragNetPreTrain
    from fragnet_b200.train.pretrain_utils import pretrain_loss
    from oracle import gat2_oracle as O
    hb = bench.make_batches("unimol", 1024, 1, 512, seed=100)[0]
    torch.manual_seed(1234)
    m = FragNetPreTrain(**bench.PT_KW).eval()
    P = O.params_from_module(m, requires_grad=False)
    m = m.cuda()
    b = {k: v.cuda() for k, v in hb.items()}
    with torch.no_grad():
        preds = m(b)
        enc = m.pretrain(b)
        loss = pretrain_loss(torch.nn.MSELoss(), preds, b)
        want_enc = O.fragnet_forward(P, hb, bench.PT_KW["num_layer"])
        want = O.pretrain_heads_forward(P, want_enc[0], want_enc[1], want_enc[2], hb)
        want_loss = O.pretrain_loss(want, hb)
    for name, a, r in zip(("bond_length", "bond_angle", "dihedral", "energy"), preds, want):
        assert a.shape == r.shape and rel_err(a, r) <= FP32_REL_TOL, (name, rel_err(a, r))
    for name, a, r in zip(("x_atoms", "x_frags", "edge_features", "fedge_features"), enc, want_enc):
        assert rel_err(a, r) <= FP32_REL_TOL, (name, rel_err(a, r))
    assert rel_err(loss, want_loss) <= FP32_REL_TOL
