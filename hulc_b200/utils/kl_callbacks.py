"""KL-annealing schedules for the fused step.

Reference: hulc/utils/kl_callbacks.py:5-60, selected by conf/callbacks/kl_schedule/{constant,linear,sigmoid}.yaml.  Once per training
epoch the trainer looks up `on_train_epoch_start(trainer, pl_module)` on every callback; the schedules answer with
`pl_module.set_kl_beta(beta(epoch))`.  `Hulc.set_kl_beta` hands the coefficient to the loss kernels and drops the CUDA graphs that were
captured under the previous value, so nothing else has to know about the schedule.

The class names, constructor arguments and the hook signature are the reference's (a callbacks YAML may point its `_target_` here; the
reference's own classes also work unchanged on `hulc_b200.models.hulc.Hulc`, they touch nothing but `current_epoch` and `set_kl_beta`).
No Lightning import is needed: a callback only has to provide the hook the trainer looks up by name.

Shape of a schedule: beta(epoch) = max_kl_beta * ramp((epoch - start) / (end - start)) inside [start_epoch, end_epoch], 0 before, max_kl_beta
after.  Subclasses only say what `ramp` is.
"""
from __future__ import annotations

from typing import Callable, Optional


def _logistic12(u: float) -> float:
    """The reference's sigmoid ramp (kl_callbacks.py:5-6, 46-49) on the unit interval: sigmoid(12 (u - 1/2)) — twelve "widths" across the
    ramp, evaluated with torch.sigmoid on a float32 tensor exactly as the reference does, so the values are bit-identical."""
    import torch

    return torch.sigmoid(torch.Tensor([12.0 * (u - 0.5)])).item()


class KLSchedule:
    """Base of the annealing callbacks (kl_callbacks.py:9-26)."""

    ramp: Optional[Callable[[float], float]] = None  # unit interval -> fraction of max_kl_beta

    def __init__(self, start_epoch: int, end_epoch: int, max_kl_beta: float):
        self.start_epoch, self.end_epoch, self.max_kl_beta = start_epoch, end_epoch, max_kl_beta

    def _anneal_fn(self, epoch: int) -> float:
        if type(self).ramp is None:
            raise NotImplementedError
        if epoch < self.start_epoch:
            return 0.0
        if epoch > self.end_epoch:
            return self.max_kl_beta
        return self._inside(epoch)

    def _inside(self, epoch: int) -> float:
        u = (epoch - self.start_epoch) / (self.end_epoch - self.start_epoch)
        return type(self).ramp(u) * self.max_kl_beta

    def on_train_epoch_start(self, trainer, pl_module) -> None:
        pl_module.set_kl_beta(self._anneal_fn(pl_module.current_epoch))


class KLConstantSchedule(KLSchedule):
    """Leaves kl_beta at what the model config says (kl_callbacks.py:29-37): takes no arguments and never calls `set_kl_beta`."""

    def __init__(self):  # noqa: D107 — no ramp parameters
        pass

    def _anneal_fn(self, epoch: int) -> None:
        return None

    def on_train_epoch_start(self, trainer, pl_module) -> None:
        return None


class KLLinearSchedule(KLSchedule):
    """Straight ramp (kl_callbacks.py:53-60): max_kl_beta * (epoch - start) / (end - start)."""

    ramp = staticmethod(lambda u: u)

    def _inside(self, epoch: int) -> float:  # the reference's order of operations (multiply, then divide): identical rounding
        return self.max_kl_beta * (epoch - self.start_epoch) / (self.end_epoch - self.start_epoch)


class KLSigmoidSchedule(KLSchedule):
    """Sigmoid ramp, twelve widths across [start_epoch, end_epoch] (kl_callbacks.py:40-50)."""

    ramp = staticmethod(_logistic12)

    def _inside(self, epoch: int) -> float:  # the reference's argument, (epoch - midpoint) / (span / 12), in its order of operations
        import torch

        span, mid = self.end_epoch - self.start_epoch, (self.end_epoch + self.start_epoch) / 2
        return torch.sigmoid(torch.Tensor([(epoch - mid) / (span / 12)])).item() * self.max_kl_beta
