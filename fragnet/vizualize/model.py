from fragnet_b200.vizualize.model import FragNetViz, FragNetFineTuneViz, FragNetFineTuneBaseViz, FragNetPreTrainViz  # noqa: F401
from fragnet_b200.model.gat.gat2 import FragNetLayerA, FragNet, FTHead1, FTHead2, FTHead3, FTHead4  # noqa: F401
from fragnet_b200._compat import reference_fallback as _fallback  # noqa: E402

__getattr__ = _fallback(__name__, __file__)   # names outside the hot path: the reference's module of the same path
