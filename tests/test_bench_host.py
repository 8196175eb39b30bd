"""Host-side pieces of bench.py and of the staging path that need no GPU: the byte model of SURVEY 8d, the compact wire
format, and the reference arm (the unmodified reference modules, from the checkout or from oracle/_ref)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_step_bytes_reproduce_the_survey_worked_totals():
    """SURVEY.md 8d / BASELINE.md section 5: 0.95-0.96 MB forward and 2.43-2.46 MB forward+backward per molecule."""
    import bench
    for shape, fwd_mb, all_mb in (("esol", 0.96, 2.46), ("unimol", 0.95, 2.43)):
        b = bench.make_batches(shape, 64, 1, 64, seed=3)[0]
        c = bench.batch_counts(b)
        fwd, both = bench.step_bytes(c)
        assert abs(fwd / c["G"] / 1e6 - fwd_mb) < 0.15 * fwd_mb, (shape, fwd / c["G"] / 1e6)
        assert abs(both / c["G"] / 1e6 - all_mb) < 0.15 * all_mb, (shape, both / c["G"] / 1e6)
        _, live = bench.step_bytes(c, live_only=True)
        assert 0.85 * both < live < both
    # one attention block, by hand: 10 nodes, 30 edges, 4-byte edge attribute
    assert bench.att_fwd_bytes(10, 30, 120) == 2 * 10 * 512 + 120 + 120 + 44 + 480
    assert bench.att_bwd_bytes(10, 30, 120, 0) == 3 * 10 * 512 + 480 + 360 + 88


def test_compact_batch_is_exact_and_a_third_of_the_bytes():
    from fragnet_b200 import synth
    from fragnet_b200.dataset.data import INDEX_KEYS, collate_fn_pt, collate_fn_pt_compact, compact_batch
    mols = synth.make_dataset("unimol", 32, seed=4)
    wide, narrow = collate_fn_pt(mols), collate_fn_pt_compact(mols)
    assert list(wide) == list(narrow)
    for k in wide:
        assert narrow[k].shape == wide[k].shape
        if k in INDEX_KEYS:
            assert narrow[k].dtype == torch.int32 and torch.equal(narrow[k].long(), wide[k])
        elif narrow[k].dtype == torch.uint8:
            assert torch.equal(narrow[k].float(), wide[k])
        else:
            assert narrow[k].dtype == wide[k].dtype and torch.equal(narrow[k], wide[k])
    assert narrow["x_atoms"].dtype == torch.uint8 and narrow["edge_attr_bonds"].dtype == torch.float32
    size = lambda d: sum(v.numel() * v.element_size() for v in d.values())
    assert size(narrow) < 0.4 * size(wide)
    # values that do not survive the round trip stay wide
    odd = dict(wide)
    odd["x_atoms"] = wide["x_atoms"] * 0.5
    odd["edge_index"] = wide["edge_index"] + 2 ** 31
    kept = compact_batch(odd)
    assert kept["x_atoms"].dtype == torch.float32 and kept["edge_index"].dtype == torch.int64


def test_packed_batch_is_the_same_dict_over_one_buffer():
    import pickle
    from fragnet_b200 import synth
    from fragnet_b200.dataset.data import PACK_SKIP, PackedBatch, collate_fn_pt, collate_fn_pt_packed, compact_batch
    mols = synth.make_dataset("unimol", 16, seed=2) + [synth.handmade("ion_pair")]
    narrow, packed = compact_batch(collate_fn_pt(mols)), collate_fn_pt_packed(mols)
    assert isinstance(packed, PackedBatch) and list(packed) == list(narrow)
    lo, hi = packed.blob.data_ptr(), packed.blob.data_ptr() + packed.blob.numel()
    for k, v in narrow.items():
        assert packed[k].dtype == v.dtype and torch.equal(packed[k], v), k
        inside = lo <= packed[k].data_ptr() < hi
        assert inside == (k not in PACK_SKIP) or v.numel() == 0, k
    assert all(off % 64 == 0 for _, _, _, off, _ in packed.layout)
    again = pickle.loads(pickle.dumps(packed))            # what a DataLoader worker hands over
    assert isinstance(again, PackedBatch) and again.layout == packed.layout and torch.equal(again.blob, packed.blob)
    assert packed.pin_memory is not None and PackedBatch(narrow).pin_memory() is not None   # (no buffer: returns itself)


def test_reference_arm_runs_the_unmodified_reference():
    """``bench.py --impl reference`` / ``cpu_baseline``: kind "reference" wherever the reference modules are available
    (checkout here, oracle/_ref bytecode on the GPU box), and its loss equals the oracle port's on the same batch."""
    import bench
    from oracle import gat2_oracle as O, ref_import
    if not ref_import.available():
        pytest.skip("no reference build")
    rate, sec, kind, how = bench.cpu_pretrain_rate("esol", 8, 1, 0)
    assert kind == "reference" and rate > 0 and sec > 0 and "unmodified" in how
    ns = ref_import.load()
    torch.manual_seed(0)
    m = ns.pretrain_heads.FragNetPreTrain(**bench.PT_KW).eval()
    b = bench.make_batches("esol", 6, 1, 6, seed=2)[0]
    with ref_import.quiet():
        bl, ba, da, e = m(b)
    mse = torch.nn.MSELoss()
    want = 2 * mse(da, b["dh_angl"]) + mse(ba, b["bnd_angl"]) + mse(e.view(-1), b["y"])
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    got = O.pretrain_loss(O.pretrain_forward(P, b), b)
    assert abs(float(got) - float(want)) <= 1e-6 * abs(float(want))


def test_compiled_reference_build_loads(tmp_path, monkeypatch):
    """oracle/build_ref.py: bytecode compiled from the checkout loads through ref_import without the checkout."""
    from oracle import build_ref, ref_import
    if ref_import.kind() != "checkout":
        pytest.skip("needs the reference checkout")
    dest = build_ref.build()
    assert os.path.isfile(os.path.join(dest, "fragnet/model/gat/gat2.py.bin"))
    assert not any(f.endswith(".py") for _, _, fs in os.walk(dest) for f in fs)      # no reference source is copied
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); from oracle import ref_import as r; "
            "assert r.kind() == 'compiled', r.kind(); ns = r.load(); print(ns.gat2.FragNetLayerA.__name__)" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], env={**os.environ, "FRAGNET_REFERENCE": str(tmp_path)},
                         capture_output=True, text=True)
    assert out.returncode == 0 and "FragNetLayerA" in out.stdout, out.stderr
