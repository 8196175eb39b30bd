"""Import-path compatibility: the reference's module paths, served by ``fragnet_b200``.

``from fragnet.model.gat.gat2 import FragNetFineTune`` (reference train/finetune/finetune_gat2.py:121),
``from fragnet.model.gat.gat2_pretrain import FragNetPreTrain`` (:217), ``from fragnet.dataset.data import
collate_fn`` (:9) etc. resolve to the B200 implementations, so the reference's entry scripts and
visualisation code import unchanged.  Only the GAT2 hot path is provided; with the reference
itself further down ``sys.path`` every other module and name (``fragnet.dataset.dataset``, ``FragNetPreTrainMasked``
...) falls through to it (``fragnet_b200/_compat.py``).
"""
from fragnet_b200._compat import overlay as _overlay

__path__ = _overlay(__path__, __name__)   # modules outside the GAT2 hot path: the reference's own, if importable
