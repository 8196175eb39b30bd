"""``FragNetPreTrain`` under the module path ``finetune_gat2.py`` loads pretraining checkpoints from
(reference fragnet/model/gat/gat2_pretrain.py:7-27, used at train/finetune/finetune_gat2.py:216-229)."""
from .pretrain_heads import FragNetPreTrain, PretrainTask  # noqa: F401
from .gat2 import FragNet  # noqa: F401
