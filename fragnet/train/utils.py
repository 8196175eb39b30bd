from fragnet_b200.train.utils import EarlyStopping, TrainerFineTune, compute_bce_loss, test_fn  # noqa: F401
