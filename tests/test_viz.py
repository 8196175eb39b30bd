"""Attention-returning task models (reference fragnet/vizualize/model.py) and the screening pipeline."""
import numpy as np
import pytest
import torch

from conftest import FP32_REL_TOL, rel_err

KW = dict(n_classes=1, num_layer=3, drop_ratio=0.1, edge_features=17, h1=64, h2=64, h3=64, h4=64, act="relu")


def test_viz_models_share_the_training_models_state_dict():
    """viz.py:560-575 loads FragNetFineTune / FragNetPreTrain checkpoints into the *Viz classes strictly."""
    from fragnet.model.gat.gat2 import FragNetFineTune
    from fragnet.model.gat.gat2_pretrain import FragNetPreTrain
    from fragnet.vizualize.model import FragNetFineTuneBaseViz, FragNetFineTuneViz, FragNetPreTrainViz
    torch.manual_seed(0)
    ft = FragNetFineTune(**KW)
    torch.manual_seed(0)
    vz = FragNetFineTuneViz(**KW)
    assert list(ft.state_dict()) == list(vz.state_dict())
    assert all(torch.equal(a, b) for a, b in zip(ft.state_dict().values(), vz.state_dict().values()))
    assert vz.pretrain.layers[-1].return_attentions and not vz.pretrain.layers[0].return_attentions
    FragNetFineTuneBaseViz(**KW).load_state_dict(ft.state_dict(), strict=True)
    pt = FragNetPreTrain(num_layer=2, edge_features=17)
    FragNetPreTrainViz(num_layer=2, edge_features=17).load_state_dict(pt.state_dict(), strict=True)


@pytest.mark.gpu
def test_viz_model_outputs_match_the_cpu_restatement():
    from fragnet.vizualize.model import FragNetFineTuneViz, FragNetPreTrainViz
    from fragnet_b200 import synth
    from fragnet_b200.dataset.data import collate_fn_pt
    from oracle import gat2_oracle as O
    hb = collate_fn_pt(synth.make_dataset("esol", 7, seed=2) + [synth.handmade("ion_pair"), synth.handmade("two_frag")])
    b = {k: v.cuda() for k, v in hb.items()}
    torch.manual_seed(3)
    m = FragNetFineTuneViz(**KW).cuda().eval()
    with torch.no_grad():
        out = m(b)
        P = O.params_from_module(m, False)
        enc = O.fragnet_forward(P, hb, 3, return_attentions=True)
        want = (O.fthead_forward(P, O.readout(enc[0], enc[1], hb)),) + tuple(enc[4:])
    assert len(out) == 5
    for a, w in zip(out, want):
        assert a.shape == w.shape and rel_err(a, w) <= FP32_REL_TOL
    torch.manual_seed(4)
    mp = FragNetPreTrainViz(num_layer=2, edge_features=17, drop_ratio=0.0).cuda().eval()
    with torch.no_grad():
        outp = mp(b)
        Pp = O.params_from_module(mp, False)
        encp = O.fragnet_forward(Pp, hb, 2, return_attentions=True)
        wantp = O.pretrain_heads_forward(Pp, encp[0], encp[1], encp[2], hb)[3]
    assert rel_err(outp[0], wantp) <= FP32_REL_TOL
    for a, w in zip(outp[1:], encp[4:]):
        assert rel_err(a, w) <= FP32_REL_TOL


@pytest.mark.gpu
def test_screening_pipeline_equals_batch_by_batch_inference():
    from fragnet.vizualize.model import FragNetFineTuneViz
    from fragnet_b200 import synth
    from fragnet_b200.dataset.arena import MoleculeArena
    from fragnet_b200.dataset.data import collate_fn
    from fragnet_b200.screen import screen
    ds = synth.make_dataset("unimol", 70, seed=6, with_pretrain_targets=False)
    torch.manual_seed(5)
    m = FragNetFineTuneViz(**KW).cuda().eval()
    arena = MoleculeArena(ds, "cuda", pretrain=False)
    order = np.random.default_rng(1).permutation(len(ds))
    seen = []
    for ids, outs in screen(m, arena, batch_size=16, ids=order):
        with torch.no_grad():
            want = m({k: v.cuda() for k, v in collate_fn([ds[int(i)] for i in ids]).items()})
        assert len(outs) == 5 and not outs[0].is_cuda
        for a, w in zip(outs, want):
            assert torch.equal(a, w.cpu())
        seen.append(np.asarray(ids))
    assert np.array_equal(np.concatenate(seen), order) and len(seen) == 5
    assert not m.training
