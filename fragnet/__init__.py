"""Import-path compatibility: the reference's module paths, served by ``fragnet_b200``.

``from fragnet.model.gat.gat2 import FragNetFineTune`` (reference train/finetune/finetune_gat2.py:121),
``from fragnet.model.gat.gat2_pretrain import FragNetPreTrain`` (:217), ``from fragnet.dataset.data import
collate_fn`` (:9) etc. resolve to the B200 implementations, so the reference's entry scripts and
visualisation code import unchanged.  Only the GAT2 hot path is provided.
"""
