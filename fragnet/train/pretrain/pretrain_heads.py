from fragnet_b200.model.gat.pretrain_heads import PretrainTask, FragNetPreTrain  # noqa: F401
from fragnet_b200._compat import reference_fallback as _fallback  # noqa: E402

__getattr__ = _fallback(__name__, __file__)   # names outside the hot path: the reference's module of the same path
