from fragnet_b200.model.gat.gat2_lite import *  # noqa: F401,F403
from fragnet_b200.model.gat.gat2_lite import FragNetLayerA, FragNet, FragNetFineTune, FTHead1, FTHead2, FTHead3, FTHead4, FTHead5  # noqa: F401
from fragnet_b200._compat import reference_fallback as _fallback  # noqa: E402

__getattr__ = _fallback(__name__, __file__)   # names outside the hot path: the reference's module of the same path
