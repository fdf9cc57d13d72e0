"""KL-annealing schedules (reference: hulc/utils/kl_callbacks.py:5-60, selected by conf/callbacks/kl_schedule/{constant,linear,sigmoid}.yaml).

Once per training epoch the trainer calls `on_train_epoch_start(trainer, pl_module)`, which sets `pl_module.set_kl_beta(beta(epoch))`.
The classes keep the reference's names, constructor arguments and hook signature, so a callbacks YAML whose `_target_` points here (or
at the reference's own module, which works unchanged on `hulc_b200.models.hulc.Hulc`) drives the fused step: `Hulc.set_kl_beta` hands the
coefficient to the loss kernels and drops CUDA graphs captured under the previous value.  No Lightning import is needed: the base class
only has to provide the hook the trainer looks up by name.
"""
from __future__ import annotations



def sigmoid(scale: float, shift: float, x: int) -> float:
    """float32 torch.sigmoid of the shifted, scaled epoch — the same call as kl_callbacks.py:5-6, so the ramp values are bit-identical."""
    import torch

    return torch.sigmoid(torch.Tensor([(x - shift) / (scale / 12)])).item()


class KLSchedule:
    """Base class for KL annealing (kl_callbacks.py:9-26)."""

    def __init__(self, start_epoch: int, end_epoch: int, max_kl_beta: float):
        self.start_epoch = start_epoch
        self.end_epoch = end_epoch
        self.max_kl_beta = max_kl_beta

    def on_train_epoch_start(self, trainer, pl_module) -> None:
        epoch = pl_module.current_epoch
        pl_module.set_kl_beta(self._anneal_fn(epoch))

    def _anneal_fn(self, epoch: int):
        raise NotImplementedError


class KLConstantSchedule(KLSchedule):
    """kl_beta stays what the model config says (kl_callbacks.py:29-37)."""

    def __init__(self):
        pass

    def on_train_epoch_start(self, trainer, pl_module) -> None:
        pass

    def _anneal_fn(self, epoch: int) -> None:
        pass


class KLSigmoidSchedule(KLSchedule):
    """0 before start_epoch, max_kl_beta after end_epoch, a sigmoid ramp (12 widths wide) in between (kl_callbacks.py:40-50)."""

    def _anneal_fn(self, epoch: int) -> float:
        if epoch < self.start_epoch:
            return 0.0
        if epoch > self.end_epoch:
            return self.max_kl_beta
        scale = self.end_epoch - self.start_epoch
        shift = (self.end_epoch + self.start_epoch) / 2
        return sigmoid(scale=scale, shift=shift, x=epoch) * self.max_kl_beta


class KLLinearSchedule(KLSchedule):
    """0 before start_epoch, max_kl_beta after end_epoch, linear in between (kl_callbacks.py:53-60)."""

    def _anneal_fn(self, epoch: int) -> float:
        if epoch < self.start_epoch:
            return 0.0
        if epoch > self.end_epoch:
            return self.max_kl_beta
        return self.max_kl_beta * (epoch - self.start_epoch) / (self.end_epoch - self.start_epoch)
