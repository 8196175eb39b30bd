"""Pretraining step loop with the reference's semantics (fragnet/train/pretrain/pretrain_utils.py:4-57).

Kept behaviour: the loss is ``2*MSE(dihedral) + MSE(angle) + MSE(energy)`` because the reference overwrites
``loss_lngth`` with the dihedral term before summing (pretrain_utils.py:22-26); the epoch loss is the sum of the
per-batch losses divided by the dataset size; ``train`` / ``validate`` keep their signatures, so
``pretrain_gat2.py`` drives this class unchanged.

What is different underneath: when the model is the drop-in ``FragNetPreTrain`` on a CUDA device, the criterion is
``nn.MSELoss()`` and the optimizer a plain ``torch.optim.Adam`` -- the configuration of pretrain_gat2.py:98-165 --
every batch runs as ONE library call (``FusedPretrainStep``: collate, forward, loss, backward, Adam) fed by a
``DevicePrefetcher`` (pinned, double-buffered host-to-device copies on a side stream).  Anything else takes the
generic ``model(batch)`` / ``loss.backward()`` / ``optimizer.step()`` path below, which is the reference's loop.
"""
from __future__ import annotations

import torch


def pretrain_loss(loss_fn, preds, batch):
    """``loss_fn(dihedral) + loss_fn(angle) + loss_fn(dihedral) + loss_fn(energy)`` (pretrain_utils.py:22-26).  With
    the reference's ``nn.MSELoss()`` on CUDA tensors the four terms, their sum and the gradients of the predictions
    come from one ``fnb_mse_sum_loss`` launch."""
    _, angle, dihedral, energy = preds
    t_dh, t_ba, y = batch["dh_angl"], batch["bnd_angl"], batch["y"]
    if (isinstance(loss_fn, torch.nn.MSELoss) and loss_fn.reduction == "mean" and dihedral.is_cuda
            and all(t.is_cuda and t.dtype == torch.float32 for t in (t_dh, t_ba, y))
            and dihedral.shape == t_dh.shape and angle.shape == t_ba.shape and energy.numel() == y.numel()
            and min(dihedral.numel(), angle.numel(), energy.numel()) > 0):
        from ..autograd import MseSumLossFn
        return MseSumLossFn.apply((2.0, 1.0, 1.0), dihedral, t_dh, angle, t_ba, energy.view(-1), y)
    l_dh = loss_fn(dihedral, t_dh)
    return l_dh + loss_fn(angle, t_ba) + l_dh + loss_fn(energy.view(-1), y)


def _plain_adam(optimizer) -> bool:
    if type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
        return False
    g = optimizer.param_groups[0]
    return not (g.get("amsgrad") or g.get("maximize") or g.get("capturable") or g.get("differentiable")
                or g.get("decoupled_weight_decay"))


class Trainer:
    def __init__(self, loss_fn=None, fused: bool = True):
        self.loss_fn = loss_fn
        self.fused = fused
        # FusedPretrainStep per (model, optimizer), held through weak references: an id() can be recycled after garbage
        # collection, a dead weakref cannot be mistaken for a new object
        self._fused_steps = []       # [(weakref(model), weakref(optimizer) | None, FusedPretrainStep | None)]

    # ---- fused path -----------------------------------------------------------------------------
    def _fused_for(self, model, optimizer, device):
        """The ``FusedPretrainStep`` bound to (model, optimizer), or None when the fused path does not apply."""
        if not self.fused or torch.device(device).type != "cuda":
            return None
        if not (isinstance(self.loss_fn, torch.nn.MSELoss) and self.loss_fn.reduction == "mean"):
            return None
        if not torch.is_tensor(next(model.parameters(), None)):
            return None
        self._fused_steps = [e for e in self._fused_steps if e[0]() is not None and (e[1] is None or e[1]() is not None)]
        for wm, wo, fs in self._fused_steps:
            if wm() is model and (wo() if wo is not None else None) is optimizer:
                if fs is not None and not fs.bound_to_model():
                    # the parameters were moved or re-assigned (model.to(), .cpu(), p.data = ...): the flat buffers no
                    # longer are the model -- hand this pair to the generic loop for good rather than train an orphan
                    import warnings
                    warnings.warn("fragnet_b200: the model's parameters were moved after the fused pretraining step was "
                                  "built; falling back to the generic training loop", RuntimeWarning)
                    self._fused_steps = [e for e in self._fused_steps if e[2] is not fs]
                    self._fused_steps.append((wm, wo, None))
                    return None
                return fs
        from ..model.gat.pretrain_heads import FragNetPreTrain
        from .fused import FusedPretrainStep
        ok = isinstance(model, FragNetPreTrain) and model.head._library_shapes() and \
            next(model.parameters()).is_cuda and (optimizer is None or _plain_adam(optimizer))
        fs = None
        if ok:
            g = optimizer.param_groups[0] if optimizer is not None else {}
            fs = FusedPretrainStep(model, lr=g.get("lr", 1e-3), betas=g.get("betas", (0.9, 0.999)),
                                   eps=g.get("eps", 1e-8), weight_decay=g.get("weight_decay", 0.0))
            if optimizer is not None:
                # the caller's optimizer stays the owner of the state: its entries become views of the fused buffers
                # (optimizer.state_dict() checkpoints them; moments loaded before the first step are adopted)
                fs.attach_optimizer(optimizer)
        import weakref
        self._fused_steps.append((weakref.ref(model), weakref.ref(optimizer) if optimizer is not None else None, fs))
        return fs

    @staticmethod
    def _run_pipelined(loader, device, launch, prefetch=None):
        """``launch(batch) -> loss tensor`` per batch; the next batches are staged while the GPU works and every
        loss is read back (the reference's ``loss.item()`` per step), one step behind the launches so that the GPU
        never waits for the host.  The sum runs in step order over the same fp32 values: identical to the reference's
        ``total_loss += loss.item()``."""
        from ..dataset.prefetch import DevicePrefetcher
        from .fused import LaggedScalars
        # this path only runs the drop-in FragNetPreTrain: tensors the GAT2 path never reads stay on the host
        staged = DevicePrefetcher(loader, device, depth=2, hot_path_only=True)
        feed = iter(staged)
        reader = LaggedScalars(lag=1)
        total, batch = 0.0, next(feed, None)
        while batch is not None:
            loss = launch(batch)
            batch = next(feed, None)
            if batch is not None and prefetch is not None:
                prefetch(batch, staged.last_event)       # its on-device collate runs underneath the step just launched
            for v in reader.push(loss):
                total += v
        for v in reader.drain():
            total += v
        return total

    # ---- the reference's interface --------------------------------------------------------------
    def train(self, model, loader, optimizer, device):
        model.train()
        fs = self._fused_for(model, optimizer, device)
        if fs is not None:
            def launch(batch):
                g = optimizer.param_groups[0]                        # LR schedulers / edited hyper-parameters keep working
                fs.lr, fs.betas, fs.eps, fs.weight_decay = float(g["lr"]), g["betas"], float(g["eps"]), float(g["weight_decay"])
                return fs.step(batch)
            return self._run_pipelined(loader, device, launch, fs.prefetch_plan) / len(loader.dataset)
        total = 0.0
        for batch in loader:
            for k in batch:
                batch[k] = batch[k].to(device)
            optimizer.zero_grad()
            loss = pretrain_loss(self.loss_fn, model(batch), batch)
            loss.backward()
            total += loss.item()
            optimizer.step()
        return total / len(loader.dataset)

    def validate(self, loader, model, device):
        model.eval()
        fs = next((f for wm, _, f in self._fused_steps if wm() is model and f is not None and f.bound_to_model()), None)
        if fs is None:
            fs = self._fused_for(model, None, device)
        if fs is not None:
            return self._run_pipelined(loader, device, fs.evaluate, fs.prefetch_plan) / len(loader.dataset)
        total = 0.0
        with torch.no_grad():
            for batch in loader:
                for k in batch:
                    batch[k] = batch[k].to(device)
                total += pretrain_loss(self.loss_fn, model(batch), batch).item()
        return total / len(loader.dataset)
