from fragnet_b200.train.pretrain_utils import Trainer, pretrain_loss  # noqa: F401
from fragnet_b200._compat import reference_fallback as _fallback  # noqa: E402

__getattr__ = _fallback(__name__, __file__)   # names outside the hot path: the reference's module of the same path
