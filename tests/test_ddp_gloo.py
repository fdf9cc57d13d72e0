"""Host-side logic of the data-parallel path at world_size 2 on CPU (gloo): the single flat-gradient all-reduce + fused Adam
of hulc_b200.ddp.FlatGradientSync must equal torch.optim.Adam on the rank-averaged gradients (what the reference's
Lightning-DDP + torch.optim.Adam does, hulc/training.py:67, hulc/models/hulc.py:239-252), including parameters that have
no gradient on any rank (GCBC's unused prior) and the initial parameter broadcast.  The Adam kernel source runs on the host
SIMT emulator build here; the NCCL path is exercised by bench.py on the GPU box."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
SPEC = {"a.weight": (33, 7), "a.bias": (33,), "unused.weight": (5, 3), "logit_scale": (), "b.weight": (64, 64)}
STEPS = 3


def _grads(rank, step):
    g = torch.Generator().manual_seed(1000 * step + rank)
    return {k: (torch.zeros(s) if k.startswith("unused") else torch.randn(s, generator=g)) for k, s in SPEC.items()}


def _worker(rank, world, port, emu_path, q):
    try:
        sys.path.insert(0, str(ROOT))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from hulc_b200 import _lib, ops
        from hulc_b200.ddp import FlatGradientSync
        from hulc_b200.engine import ParamStore

        _lib._LIB = _lib.Library(emu_path, allow_missing=True)
        ops._DEVICE_TYPE = "cpu"

        class Eng:  # the part of HulcEngine the sync object touches
            lr = 1e-2

            def encoder_grad_split(self):  # chunked variant: [split, end) is reduced first (asynchronously), [0, split) afterwards
                return self.ps.offsets["logit_scale"][0]

        eng = Eng()
        eng.ps = ParamStore(SPEC, "cpu")
        # ranks start from different weights: the broadcast must make them rank 0's
        g0 = torch.Generator().manual_seed(7 + rank)
        for k in SPEC:
            eng.ps.p[k].copy_(torch.randn(SPEC[k], generator=g0))
        sync = FlatGradientSync(eng)
        sync.broadcast_parameters(0)
        g_ref = torch.Generator().manual_seed(7)
        ref = {k: torch.randn(SPEC[k], generator=g_ref).requires_grad_(True) for k in SPEC}
        for k in SPEC:
            assert torch.equal(eng.ps.p[k], ref[k].detach()), f"broadcast {k}"
        opt = torch.optim.Adam(list(ref.values()), lr=Eng.lr)
        for step in range(STEPS):
            eng.ps.zero_grad()
            mine = _grads(rank, step)
            for k in SPEC:
                eng.ps.g[k].copy_(mine[k])
            if step == 2:  # the overlapped two-chunk exchange (HulcEngine.capture_split)
                sync.sync_head()
                sync.sync_tail()
            else:
                sync.sync(async_op=bool(step & 1))
            sync.step()
            allg = [_grads(r, step) for r in range(world)]
            for k in SPEC:
                ref[k].grad = sum(g[k] for g in allg) / world
            opt.step()
            for k in SPEC:
                torch.testing.assert_close(eng.ps.p[k], ref[k].detach(), rtol=1e-5, atol=1e-6, msg=lambda m: f"step {step} {k}: {m}")
        assert eng.ps.step_count == STEPS
        # every rank holds the same parameters bit for bit
        mine = eng.ps.flat.clone()
        other = eng.ps.flat.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(mine, other)
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        import traceback

        q.put((rank, traceback.format_exc()))
        raise e


@pytest.mark.timeout(300)
def test_flat_allreduce_adam_world2(emu_lib_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(emu_lib_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
