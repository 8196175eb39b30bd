from fragnet_b200.model.gat.gat2_pretrain import FragNetPreTrain, PretrainTask, FragNet  # noqa: F401
