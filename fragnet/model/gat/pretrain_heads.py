from fragnet_b200.model.gat.pretrain_heads import PretrainTask, FragNetPreTrain  # noqa: F401
from fragnet_b200.model.gat.gat2 import FragNet  # noqa: F401
