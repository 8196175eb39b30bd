from fragnet_b200.train.pretrain_utils import Trainer, pretrain_loss  # noqa: F401
