from fragnet_b200.model.gat.gat2_edge import *  # noqa: F401,F403
from fragnet_b200.model.gat.gat2_edge import FragNetLayerA, FragNet, FragNetFineTune, FTHead1, FTHead2, FTHead3, FTHead4, FTHead5  # noqa: F401
