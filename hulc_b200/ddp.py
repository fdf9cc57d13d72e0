"""Data-parallel training of the HULC step: one process per GPU, full replica, batch sharded across ranks.

The reference trains with Lightning's DDP strategy (hulc/training.py:67 → `Trainer(strategy="ddp")`,
conf/trainer/play_trainer.yaml), i.e. torch DDP's bucketed all-reduce of every parameter's `.grad`.  Here the gradients
of all 47 M parameters already live in ONE flat fp32 buffer (`engine.ParamStore.grad`), so the whole exchange is a single
`all_reduce(sum)` over that buffer (NCCL over NVLink on the GPU box; gloo in the CPU tests), and the mean is folded into
the fused Adam launch as `grad_scale = 1 / world_size`.  Nothing else on the path is collective: the CLIP loss of the
reference is rank-local (hulc/models/hulc.py:685-692) and the logged scalars are reduced at epoch end only.

Parameters without a gradient (GCBC never touches the prior) contribute zeros on every rank — the flat
buffer is zero-filled at the start of each step — so all ranks reduce buffers of identical layout.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class FlatGradientSync:
    """`sync()` after the engine's backward, `step()` instead of the optimizer: all-reduce + Adam on the flat buffers."""

    def __init__(self, engine, process_group: Optional[dist.ProcessGroup] = None):
        self.engine = engine
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._work = None

    def broadcast_parameters(self, src: int = 0):
        """Every rank starts from rank `src`'s weights and Adam state (what torch DDP does in its constructor)."""
        if self.world == 1:
            return
        ps = self.engine.ps
        for t in (ps.flat, ps.exp_avg, ps.exp_avg_sq):
            dist.broadcast(t, src=src, group=self.group)
        step = torch.tensor([ps.step_count], dtype=torch.int64, device=ps.flat.device)
        dist.broadcast(step, src=src, group=self.group)
        ps.step_count = int(step.item())
        ps.step_dev.fill_(ps.step_count)

    def sync(self, async_op: bool = False):
        """Sum the flat gradient buffer over the ranks (the one collective of the step)."""
        if self.world == 1:
            return None
        self._work = dist.all_reduce(self.engine.ps.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        return self._work

    def sync_head(self):
        """Overlapped variant, part 1: start the all-reduce of everything behind the perceptual encoders' gradients (97 % of the bytes; final
        once the decoder / prior / posterior / goal-encoder backward is done, see HulcEngine.capture_split).  NCCL runs on its own stream,
        ordered after the work already queued on the current stream, so the encoders' backward that follows overlaps it."""
        if self.world == 1:
            return None
        split = self.engine.encoder_grad_split()
        self._work = dist.all_reduce(self.engine.ps.grad[split:], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return self._work

    def sync_tail(self):
        """Part 2, after the encoders' backward: the remaining (small) slice."""
        if self.world == 1:
            return None
        split = self.engine.encoder_grad_split()
        dist.all_reduce(self.engine.ps.grad[:split], op=dist.ReduceOp.SUM, group=self.group)

    def step(self, lr: Optional[float] = None):
        """Adam on the averaged gradient: waits for an outstanding async all-reduce first."""
        if self._work is not None:
            self._work.wait()
            self._work = None
        self.engine.ps.adam_step(lr=self.engine.lr if lr is None else lr, grad_scale=1.0 / self.world)
