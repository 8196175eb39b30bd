"""On-device batch assembly from the packed arena (fnb_arena_assemble) against the host collate: the product's
``collate_fn_pt`` and the restatement of the reference's ``collate_fn`` / ``get_incr_*`` (fragnet/dataset/data.py:11-113,
:877-1032).  Integer and byte work: bit-exact, same keys, shapes and dtypes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _same(got, want):
    assert list(got) == list(want)
    for k in want:
        g, w = got[k].cpu(), want[k]
        assert g.dtype == w.dtype, (k, g.dtype, w.dtype)
        assert g.shape == w.shape, (k, g.shape, w.shape)
        assert torch.equal(g, w), k


def _dataset(n=60, seed=3):
    from fragnet_b200 import synth
    ds = synth.make_dataset("unimol", n, seed=seed)
    return ds + [synth.handmade(k) for k in ("two_atom", "ion_pair", "single_frag", "two_frag")]


@pytest.mark.parametrize("pretrain", [True, False])
def test_arena_batches_equal_host_collate(pretrain):
    from fragnet_b200.dataset.arena import MoleculeArena
    from fragnet_b200.dataset.data import collate_fn, collate_fn_pt
    from oracle import collate_oracle
    ds = _dataset()
    arena = MoleculeArena(ds, "cuda", pretrain=pretrain)
    host = collate_fn_pt if pretrain else collate_fn
    rng = np.random.default_rng(0)
    cases = [np.arange(len(ds)), rng.permutation(len(ds))[:17], np.array([5, 5, 5, 0, len(ds) - 1, 5]),
             np.array([len(ds) - 1]), rng.integers(0, len(ds), size=1500)]
    for ids in cases:
        got = arena.batch(ids)
        picked = [ds[int(i)] for i in ids]
        _same(got, host(picked))
        _same(got, collate_oracle.collate(picked, pretrain=pretrain))
    assert int(arena._status.item()) == 0
    # validity contract: a batch lives in one of SLOTS persistent device slots and survives SLOTS - 1 further batches
    first = arena.batch(cases[1])
    for _ in range(MoleculeArena.SLOTS - 1):
        arena.batch(cases[0])
    _same(first, host([ds[int(i)] for i in cases[1]]))
    again = arena.batch(cases[2])
    assert again["x_atoms"].data_ptr() == first["x_atoms"].data_ptr()      # the slot is reused from here on


def test_arena_ragged_and_empty():
    """Single-fragment molecules have no fragment links (zero-column index tensors); an empty id list gives empty
    tensors of the right rank."""
    from fragnet_b200.dataset.arena import MoleculeArena
    from fragnet_b200.dataset.data import collate_fn_pt
    import copy
    from fragnet_b200 import synth
    ds = _dataset(20, seed=11)
    bare = []
    for kind in ("two_atom", "single_frag"):      # the same molecules without any fragment link
        d = copy.deepcopy(synth.handmade(kind))
        d.frag_index = torch.zeros((2, 0), dtype=torch.long)
        d.cnx_attr = torch.zeros((0, d.cnx_attr.shape[1]))
        d.node_feautures_fbondg = torch.zeros((0, d.node_feautures_fbondg.shape[1]))
        d.edge_index_fbondg = torch.zeros((2, 0), dtype=torch.int32)
        d.edge_attr_fbondg = torch.zeros((0, d.edge_attr_fbondg.shape[1]))
        bare.append(d)
    ds = bare[:1] + ds + bare[1:]
    arena = MoleculeArena(ds, "cuda")
    for ids in (np.array([0, len(ds) - 1]), np.array([0, 3, len(ds) - 1, 4, 0])):
        got = arena.batch(ids)
        _same(got, collate_fn_pt([ds[int(i)] for i in ids]))
    got = arena.batch(np.array([0, len(ds) - 1]))
    assert got["frag_index"].shape == (2, 0) and got["edge_index_fbonds"].shape == (2, 0)
    empty = arena.batch(np.zeros(0, dtype=np.int64))
    assert empty["x_atoms"].shape == (0, ds[0].x_atoms.shape[1]) and empty["edge_index"].shape == (2, 0)
    with pytest.raises(IndexError):
        arena.batch([len(ds)])


def test_arena_stress_shape_and_loader():
    from fragnet_b200 import synth
    from fragnet_b200.dataset.arena import ArenaLoader, MoleculeArena
    from fragnet_b200.dataset.data import collate_fn_pt
    ds = synth.make_dataset("stress", 6, seed=2) + synth.make_dataset("esol", 30, seed=2)
    arena = MoleculeArena(ds, "cuda")
    loader = ArenaLoader(arena, batch_size=8, shuffle=False, drop_last=False)
    assert len(loader) == 5
    for b, got in enumerate(loader):
        _same(got, collate_fn_pt(ds[b * 8:(b + 1) * 8]))
    g = torch.Generator().manual_seed(1)
    seen = sum(int(batch["y"].shape[0]) for batch in ArenaLoader(arena, 8, shuffle=True, drop_last=True, generator=g))
    assert seen == 32


def test_model_on_arena_batch_is_bitwise_the_collated_batch():
    from fragnet.model.gat.gat2_pretrain import FragNetPreTrain
    from fragnet_b200.dataset.arena import MoleculeArena
    from fragnet_b200.dataset.data import collate_fn_pt
    ds = _dataset(40, seed=5)
    torch.manual_seed(0)
    m = FragNetPreTrain(num_layer=2, drop_ratio=0.0, edge_features=17).cuda().eval()
    ids = np.arange(len(ds))[::-1].copy()
    a = m(MoleculeArena(ds, "cuda").batch(ids))
    b = m({k: v.cuda() for k, v in collate_fn_pt([ds[int(i)] for i in ids]).items()})
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_prefetcher_widens_the_compact_wire_format_bit_exactly():
    """Compact host batches (uint8 feature matrices, int32 indices) arrive on the device with collate_fn_pt's dtypes and
    values; device-resident batches pass through untouched; ragged element counts exercise the kernel's tail."""
    from fragnet_b200 import synth
    from fragnet_b200.dataset.data import collate_fn_pt, compact_batch, pack_batch
    from fragnet_b200.dataset.prefetch import DevicePrefetcher, staged_bytes
    ds = synth.make_dataset("unimol", 37, seed=13) + [synth.handmade("ion_pair")]
    wide = [collate_fn_pt(ds[i:i + n]) for i, n in ((0, 7), (7, 30), (37, 1), (3, 5))]
    narrow = [compact_batch(b, pin=True) for b in wide]
    assert staged_bytes(narrow[1]) < 0.45 * staged_bytes(wide[1])
    # packed: the same tensors as views of one pinned buffer per batch -> one copy per batch (hot-path mode), the
    # tensor-by-tensor path when the unread tensors are wanted too; wide (unnarrowed) batches pack as well
    packed = [pack_batch(b, pin=True) for b in narrow]
    assert all(p.blob.is_pinned() and p["x_atoms"].data_ptr() >= p.blob.data_ptr() for p in packed)
    packed_wide = [pack_batch(b).pin_memory() for b in wide]
    assert all(p.blob.is_pinned() for p in packed_wide)
    for hot, feed in ((False, narrow), (True, narrow), (True, packed), (False, packed), (True, packed_wide)):
        n_seen = 0
        for g, w in zip(DevicePrefetcher(iter(feed), "cuda", depth=2, hot_path_only=hot), wide):
            n_seen += 1          # (a yielded batch is only valid until `depth` further ones have been requested)
            torch.cuda.synchronize()
            for k, v in w.items():
                if hot and k in ("edge_attr", "cnx_attr"):
                    assert k not in g
                elif hot and k == "x_frags":
                    assert g[k].device.type == "meta" and g[k].shape == v.shape and g[k].dtype == v.dtype
                else:
                    assert g[k].dtype == v.dtype and torch.equal(g[k].cpu(), v), k
        assert n_seen == len(wide)
    dev = [{k: v.cuda() for k, v in b.items()} for b in wide[:2]]
    through = list(DevicePrefetcher(iter(dev), "cuda"))
    assert all(t[k] is d[k] for t, d in zip(through, dev) for k in d)


def test_overlapped_assembly_and_prefetched_collate_train_bitwise_the_same():
    """ArenaLoader(overlap=True) assembles every batch on a side stream underneath the step in flight and the trainer
    queues its collate there too: the batches and a two-epoch training run are bitwise those of the in-stream loader."""
    import copy
    from fragnet.model.gat.gat2_pretrain import FragNetPreTrain
    from fragnet_b200.dataset.arena import ArenaLoader, DeviceBatch, MoleculeArena
    from fragnet_b200.dataset.data import collate_fn_pt
    from fragnet_b200.train.pretrain_utils import Trainer
    ds = _dataset(57, seed=9)
    arena = MoleculeArena(ds, "cuda")
    for b, got in enumerate(ArenaLoader(arena, batch_size=8, overlap=True)):
        assert isinstance(got, DeviceBatch) and got.ready_event is not None
        _same(got, collate_fn_pt(ds[b * 8:(b + 1) * 8]))
    torch.manual_seed(4)
    m0 = FragNetPreTrain(num_layer=2, drop_ratio=0.1, edge_features=17).cuda()
    runs = []
    for overlap in (True, False):
        m = copy.deepcopy(m0)
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        tr = Trainer(torch.nn.MSELoss())
        torch.manual_seed(8)
        losses = [tr.train(m, ArenaLoader(MoleculeArena(ds, "cuda"), batch_size=10, overlap=overlap), opt, "cuda")
                  for _ in range(2)]
        runs.append((losses, {k: v.clone() for k, v in m.state_dict().items()}))
    assert runs[0][0] == runs[1][0]
    for k in runs[0][1]:
        assert torch.equal(runs[0][1][k], runs[1][1][k]), k
