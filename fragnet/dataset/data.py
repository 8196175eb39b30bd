from fragnet_b200.dataset.data import collate_fn, collate_fn_pt  # noqa: F401
