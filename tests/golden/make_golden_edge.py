"""Generate ``tests/golden/gat2_edge_golden.pt`` from the UNMODIFIED reference ``fragnet/model/gat/gat2_edge.py``.

Build container only (needs ``/root/reference``):  ``python tests/golden/make_golden_edge.py``.  Same conventions as
``make_golden.py``: weights and inputs are regenerated from seeds by the tests; stored are the reference's outputs
(predictions, encoder outputs, attention sums of a bare layer, gradients of the small live parameters, checksums of the
rest).  ``cnx_attr`` is 8 wide here (gat2_edge.py:46): the synthetic 6-wide connection one-hot plus two more columns.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from make_golden import checksums, fix_bias, is_small  # noqa: E402
from fragnet_b200 import synth  # noqa: E402
from fragnet_b200.dataset.data import collate_fn  # noqa: E402
from oracle import ref_import  # noqa: E402

WEIGHT_SEED, DATA_SEED = 2468, 91
EDGE_KW = dict(n_classes=1, num_layer=3, drop_ratio=0.1, edge_features=17, h1=128, h2=256, h3=256, h4=128, act="relu",
               fthead="FTHead3")


def edge_batch():
    mols = synth.make_dataset("esol", 6, seed=DATA_SEED, with_pretrain_targets=False)
    mols += synth.make_dataset("unimol", 4, seed=DATA_SEED + 1, with_pretrain_targets=False)
    mols += [synth.handmade(k) for k in ("two_atom", "ion_pair", "single_frag", "two_frag")]
    b = collate_fn(mols)
    gen = torch.Generator().manual_seed(DATA_SEED + 2)
    extra = torch.rand(b["cnx_attr"].shape[0], 2, generator=gen)            # the two columns the 6-wide featuriser lacks
    b["cnx_attr"] = torch.cat((b["cnx_attr"], extra), dim=1)
    return b


def layer_inputs(b, seed=5):
    """Inputs of one bare ``FragNetLayerA`` call on 128-wide features (a layer >= 1 of the stack)."""
    gen = torch.Generator().manual_seed(seed)
    r = lambda n: torch.randn(n, 128, generator=gen)
    na, nf, nb = b["x_atoms"].shape[0], b["x_frags"].shape[0], b["edge_index"].shape[1]
    bond = r(nb)
    return (r(na), b["edge_index"], bond, b["frag_index"], r(nf), b["atom_to_frag_ids"], bond,
            b["edge_index_bonds_graph"], b["edge_attr_bonds"], b["cnx_attr"])


def main():
    edge = ref_import.load_edge()
    batch = edge_batch()
    torch.manual_seed(WEIGHT_SEED)
    with ref_import.quiet():
        model = edge.FragNetFineTune(**EDGE_KW)
    fix_bias(model)
    model.eval()
    with ref_import.quiet():
        enc = model.pretrain(batch)
        pred = model(batch)
        target = torch.linspace(-1.0, 1.0, pred.numel()).view_as(pred)
        torch.nn.functional.mse_loss(pred, target).backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    torch.manual_seed(WEIGHT_SEED + 1)
    layer = edge.FragNetLayerA(num_heads=4, return_attentions=True)
    layer.bias.data.zero_()             # uninitialised upstream (gat2_edge.py:35), never read
    with ref_import.quiet(), torch.no_grad():
        layer_out = layer(*layer_inputs(batch))
    out = {
        "weight_seed": WEIGHT_SEED, "data_seed": DATA_SEED, "kwargs": EDGE_KW, "torch_version": str(torch.__version__),
        "batch_checksums": checksums(batch),
        "state_keys": list(model.state_dict()),
        "state_checksums": checksums(model.state_dict()),
        "pred": pred.detach().clone(),
        "encoder": [t.detach().clone() for t in enc],
        "grads": {k: v.clone() for k, v in grads.items() if is_small(k) or "cnx_attr_transform" in k},
        "grad_checksums": checksums(grads),
        "grad_none": sorted(k for k, p in model.named_parameters() if p.grad is None),
        "layer_state_keys": list(layer.state_dict()),
        "layer_out": [t.detach().clone() for t in layer_out],
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gat2_edge_golden.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path), "bytes;", len(out["grads"]), "gradients stored;", len(out["grad_none"]), "params without gradient")


if __name__ == "__main__":
    main()
