"""Generate ``tests/golden/gat2_lite_golden.pt`` from the UNMODIFIED reference ``fragnet/model/gat/gat2_lite.py``.

Build container only (needs ``/root/reference``):  ``python tests/golden/make_golden_lite.py``.  Same conventions as
``make_golden.py``: weights and inputs are regenerated from seeds by the tests; stored are the reference's outputs
(predictions, encoder outputs, gradients of the small live parameters, checksums of the rest).
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

WEIGHT_SEED, DATA_SEED = 4321, 78
LITE_KW = dict(n_classes=1, num_layer=3, drop_ratio=0.1, edge_features=17, h1=128, h2=256, h3=256, h4=128, act="relu",
               fthead="FTHead3")


def lite_batch():
    mols = synth.make_dataset("esol", 6, seed=DATA_SEED, with_pretrain_targets=False)
    mols += synth.make_dataset("unimol", 3, seed=DATA_SEED + 1, with_pretrain_targets=False)
    mols += [synth.handmade(k) for k in ("two_atom", "ion_pair", "single_frag", "two_frag")]
    return collate_fn(mols)


def main():
    lite = ref_import.load_lite()
    batch = lite_batch()
    torch.manual_seed(WEIGHT_SEED)
    with ref_import.quiet():
        model = lite.FragNetFineTune(**LITE_KW)
    fix_bias(model)
    model.eval()
    with ref_import.quiet():
        enc = model.pretrain(batch)
        pred = model(batch)
        target = torch.linspace(-1.0, 1.0, pred.numel()).view_as(pred)
        torch.nn.functional.mse_loss(pred, target).backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    out = {
        "weight_seed": WEIGHT_SEED, "data_seed": DATA_SEED, "kwargs": LITE_KW, "torch_version": str(torch.__version__),
        "batch_checksums": checksums(batch),
        "state_keys": list(model.state_dict()),
        "state_checksums": checksums(model.state_dict()),
        "pred": pred.detach().clone(),
        "encoder": [t.detach().clone() for t in enc[:3]],
        "grads": {k: v.clone() for k, v in grads.items() if is_small(k)},
        "grad_checksums": checksums(grads),
        "grad_none": sorted(k for k, p in model.named_parameters() if p.grad is None),
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gat2_lite_golden.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path), "bytes;", len(out["grads"]), "gradients stored;", len(out["grad_none"]), "params without gradient")


if __name__ == "__main__":
    main()
