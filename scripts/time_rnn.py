"""Time the persistent recurrence kernels alone (B = 64, S = 32, H = 2048), CUDA events over back-to-back launches.
    HULC_B200_RNN_GEN=1|2  HULC_B200_RNN_POLL=0|1  python scripts/time_rnn.py"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from hulc_b200 import ops  # noqa: E402

H, B, S = 2048, 64, 32
g = torch.Generator().manual_seed(0)
W = ((torch.rand(H, H, generator=g) * 2 - 1) / H ** 0.5).cuda()
W16 = W.to(torch.bfloat16)
pre = (torch.randn(S, B, H, generator=g) * 0.5).cuda()
hbuf = torch.zeros(S + 2, B, H, device="cuda")
dbuf = torch.zeros(S + 1, B, H, device="cuda")
x16 = torch.empty((S + 1) * B * H, dtype=torch.bfloat16, device="cuda")
st, sp, sd = hbuf.stride(0), pre.stride(0), dbuf.stride(0)


def fwd32():
    ops.rnn_tc_seq(W, hbuf[0], hbuf[1], pre[0], S, prev_step=st, out_step=st, add_step=sp, act=1)


def bwd32():
    ops.rnn_tc_seq(W, dbuf[S], dbuf[S - 1], pre[S - 1], S, prev_step=-sd, out_step=-sd, add_step=-sp, gate0=hbuf[S], gate_step=-st, act=0, transW=True)


def fwd16():
    ops.rnn_seq_bf16(W16, hbuf[0], x16, hbuf[1], pre[0], S, out_step=st, add_step=sp, act=1)


def bwd16():
    ops.rnn_seq_bf16(W16, dbuf[S], x16, dbuf[S - 1], pre[S - 1], S, out_step=-sd, add_step=-sp, gate0=hbuf[S], gate_step=-st, act=0, transW=True)


for name, fn in (("tf32 fwd", fwd32), ("tf32 bwd", bwd32), ("bf16 fwd", fwd16), ("bf16 bwd", bwd16)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"gen={os.environ.get('HULC_B200_RNN_GEN', '2')} poll={os.environ.get('HULC_B200_RNN_POLL', '0')} {name}: {ms * 1e3:8.1f} us per launch (incl. the init launch) = "
          f"{ms * 1e3 / S:6.2f} us per step")
