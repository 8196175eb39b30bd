from fragnet_b200.vizualize.model import FragNetViz, FragNetFineTuneViz, FragNetFineTuneBaseViz, FragNetPreTrainViz  # noqa: F401
from fragnet_b200.model.gat.gat2 import FragNetLayerA, FragNet, FTHead1, FTHead2, FTHead3, FTHead4  # noqa: F401
