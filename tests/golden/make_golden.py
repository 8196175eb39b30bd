"""Generate ``tests/golden/gat2_golden.pt`` from the UNMODIFIED reference code.

Run in the build container only (needs ``/root/reference``):  ``python tests/golden/make_golden.py``.
The reference ships no golden vectors for this path, so these are outputs of the
reference's own ``gat2.py`` / ``pretrain_heads.py`` (loaded through ``oracle/ref_import.py`` with the
third-party shims) on a small seeded synthetic batch.  Weights and inputs are NOT stored: both are
regenerated from seeds by the tests (the product modules consume the global RNG in the same order as the
reference's constructors, which the tests assert through the stored per-tensor checksums).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from fragnet_b200 import synth  # noqa: E402
from fragnet_b200.dataset.data import collate_fn_pt  # noqa: E402
from oracle import ref_import  # noqa: E402

WEIGHT_SEED, DATA_SEED = 1234, 77
FT_KW = dict(n_classes=1, num_layer=4, drop_ratio=0.1, h1=128, h2=1024, h3=1024, h4=512, act="relu", fthead="FTHead3")
PT_KW = dict(num_layer=4, drop_ratio=0.2, num_heads=4, emb_dim=128, atom_features=167, frag_features=167,
             edge_features=17, fedge_in=6, fbond_edge_in=6)


def golden_batch():
    mols = synth.make_dataset("esol", 6, seed=DATA_SEED) + synth.make_dataset("stress", 1, seed=DATA_SEED + 1)
    mols += [synth.handmade(k) for k in ("two_atom", "ion_pair", "single_frag", "two_frag")]
    return collate_fn_pt(mols)


def checksums(tensors):
    return {k: float(v.double().abs().sum()) for k, v in tensors.items() if torch.is_tensor(v)}


def fix_bias(model):
    """The reference leaves ``bias`` uninitialised (gat2.py:81); the product zero-fills it."""
    for name, p in model.named_parameters():
        if name.endswith("bias") and name.split(".")[-2].isdigit() and name.split(".")[-3] == "layers":
            p.data.zero_()


def is_small(key):
    """Parameters whose full gradient is stored (the rest are pinned by checksums)."""
    parts = key.split(".")
    return parts[-1] in ("a_b", "a", "f", "f_a_b") or parts[-2] in ("edge_attr_bond_embed", "edge_attr_fbond_embed") \
        or key.endswith("projection_a.bias")


def viz_forward(encoder, batch):
    """The layer loop of the reference's vizualize/model.py:72-142 (last layer returns attentions)."""
    x_atoms, x_frags = encoder.dropout(batch["x_atoms"]), encoder.dropout(batch["x_frags"])
    bond_nodes, fbond_nodes, edge_attr = batch["node_features_bonds"], batch["node_features_fbonds"], batch["edge_attr"]
    post = lambda t: encoder.act(encoder.dropout(t))
    attn = None
    for li, layer in enumerate(encoder.layers):
        layer.return_attentions = li == len(encoder.layers) - 1
        out = layer(x_atoms, batch["edge_index"], edge_attr, batch["frag_index"], x_frags, batch["atom_to_frag_ids"],
                    bond_nodes, batch["edge_index_bonds_graph"], batch["edge_attr_bonds"],
                    fbond_nodes, batch["edge_index_fbonds"], batch["edge_attr_fbonds"])
        x_atoms, x_frags, bond_nodes, fbond_nodes = (post(t) for t in out[:4])
        edge_attr = bond_nodes
        attn = out[4:] if len(out) > 4 else attn
        layer.return_attentions = False
    return (x_atoms, x_frags, bond_nodes, fbond_nodes) + tuple(attn)


def main():
    ref = ref_import.load()
    batch = golden_batch()
    out = {"weight_seed": WEIGHT_SEED, "data_seed": DATA_SEED, "ft_kwargs": FT_KW, "pt_kwargs": PT_KW,
           "batch_checksums": checksums(batch), "torch_version": str(torch.__version__)}
    with ref_import.quiet():
        torch.manual_seed(WEIGHT_SEED)
        ft = ref.gat2.FragNetFineTune(**FT_KW).eval()
        fix_bias(ft)
        out["ft_state_keys"] = list(ft.state_dict().keys())
        out["ft_state_checksums"] = checksums(ft.state_dict())
        pred = ft(batch)
        out["ft_pred"] = pred.detach().clone()
        pred.sum().backward()
        out["ft_grads"] = {k: p.grad.clone() for k, p in ft.named_parameters() if p.grad is not None and is_small(k)}
        out["ft_grads"]["pretrain.layers.1.projection_b.weight"] = ft.pretrain.layers[1].projection_b.weight.grad.clone()
        out["ft_grad_checksums"] = {k: float(p.grad.double().abs().sum()) for k, p in ft.named_parameters()
                                    if p.grad is not None}
        out["ft_grad_none"] = [k for k, p in ft.named_parameters() if p.grad is None]
        with torch.no_grad():
            enc = viz_forward(ft.pretrain, batch)
            names = ("x_atoms", "x_frags", "edge_features", "fedge_features",
                     "attn_atoms", "attn_frags", "attn_bonds", "attn_fbonds")
            out["encoder"] = {n: t.clone() for n, t in zip(names, enc)}
            masked = {}
            for attr, val in (("bond_mask", 2), ("atom_mask_individual", 3), ("frag_bond_mask", 0)):
                for layer in ft.pretrain.layers:
                    setattr(layer, attr, val)
                masked[attr] = ft(batch).clone()
                for layer in ft.pretrain.layers:
                    setattr(layer, attr, None)
            out["ft_masked_pred"] = masked
        torch.manual_seed(WEIGHT_SEED + 1)
        pt = ref.pretrain_heads.FragNetPreTrain(**PT_KW).eval()
        fix_bias(pt)
        out["pt_state_keys"] = list(pt.state_dict().keys())
        out["pt_state_checksums"] = checksums(pt.state_dict())
        preds = pt(batch)
        out["pt_preds"] = [t.detach().clone() for t in preds]
        mse = torch.nn.MSELoss()
        l_dh = mse(preds[2], batch["dh_angl"])
        loss = l_dh + mse(preds[1], batch["bnd_angl"]) + l_dh + mse(preds[3].view(-1), batch["y"])
        out["pt_loss"] = loss.detach().clone()
        loss.backward()
        out["pt_grad_checksums"] = {k: float(p.grad.double().abs().sum()) for k, p in pt.named_parameters()
                                    if p.grad is not None}
        out["pt_grads"] = {k: p.grad.clone() for k, p in pt.named_parameters() if p.grad is not None and is_small(k)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gat2_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
