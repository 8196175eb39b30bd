from fragnet_b200.model.gat.pretrain_heads import PretrainTask, FragNetPreTrain  # noqa: F401
