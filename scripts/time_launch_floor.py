"""What does one dependent kernel launch cost inside a replayed CUDA graph?  Chains of N tiny launches (a 4 KB cast; a 64 x 128 x 128 bf16
product; a 64 x 2048 x 2048 bf16 product — the prior / goal-encoder shape), captured once, replayed, CUDA-event timed.
    python scripts/time_launch_floor.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from hulc_b200 import ops  # noqa: E402

dev = "cuda"
N = 200
x = torch.randn(8, 128, device=dev)
xb = torch.zeros(8, 128, dtype=torch.bfloat16, device=dev)
A = torch.randn(64, 128, device=dev).to(torch.bfloat16)
B = torch.randn(128, 128, device=dev).to(torch.bfloat16)
C = torch.zeros(64, 128, device=dev)
A2 = torch.randn(64, 2048, device=dev).to(torch.bfloat16)
B2 = torch.randn(2048, 2048, device=dev).to(torch.bfloat16)
C2 = torch.zeros(64, 2048, device=dev)
A3 = torch.randn(2048, 2048, device=dev).to(torch.bfloat16)
C3 = torch.zeros(2048, 2048, device=dev)


def chain(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (10 * N)


print("cast 4 KB              : %.2f us per launch in a graph chain" % chain(lambda: ops.cast_bf16(x, xb)))
print("gemm_bf16 64x128x128   : %.2f us" % chain(lambda: ops.gemm_bf16(A, B, C, None, transB=True)))
print("gemm_bf16 64x2048x2048 : %.2f us" % chain(lambda: ops.gemm_bf16(A2, B2, C2, None, transB=True)))
print("gemm_bf16 2048^3       : %.2f us" % chain(lambda: ops.gemm_bf16(A3, B2, C3, None, transB=True)))
