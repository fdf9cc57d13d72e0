"""CUDA-core vs tensor-core GEMM on the small / skinny products of the step (time per launch back to back, error vs float64).
    python scripts/dbg_gemm_small.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from hulc_b200 import ops  # noqa: E402

shapes = [  # M N K tA tB beta
    (2048, 2048, 32, 1, 0, 1), (2048, 384, 32, 1, 0, 1), (2048, 128, 32, 1, 0, 1), (2048, 128, 64, 1, 0, 1), (4096, 128, 64, 1, 0, 1),
    (2048, 128, 128, 0, 1, 0), (2048, 128, 128, 0, 0, 0), (128, 128, 2048, 1, 0, 1), (32, 32, 2048, 0, 1, 0), (32, 128, 4096, 0, 1, 0),
    (64, 128, 4096, 0, 0, 0), (64, 128, 2048, 0, 0, 0), (32, 128, 2048, 0, 0, 0), (64, 32, 2048, 0, 0, 0), (32, 2048, 384, 0, 1, 0),
    (64, 2048, 128, 0, 1, 0), (64, 4096, 128, 0, 1, 0), (32, 2048, 128, 0, 1, 0), (128, 4096, 32, 1, 0, 1), (32, 4096, 128, 0, 0, 0),
    (2048, 32, 64, 1, 0, 1), (32, 2048, 32, 1, 0, 1), (32, 2048, 32, 0, 0, 0), (64, 2048, 32, 0, 1, 0), (32, 128, 32, 0, 1, 0), (128, 32, 32, 1, 0, 1),
    (32, 32, 128, 0, 1, 0), (32, 128, 32, 1, 0, 1), (2048, 2048, 184, 0, 0, 0), (184, 2048, 2048, 1, 0, 1), (2048, 184, 2048, 0, 1, 0),
]
g = torch.Generator().manual_seed(0)
print(f"{'M':>5} {'N':>5} {'K':>5} tA tB |  simt us   tc1 us   tc3 us | err simt / tc1 / tc3 (max abs, ref scale)")
for M, N, K, ta, tb, beta in shapes:
    A = torch.randn((K, M) if ta else (M, K), generator=g).cuda()
    B = torch.randn((N, K) if tb else (K, N), generator=g).cuda()
    C0 = torch.randn(M, N, generator=g).cuda()
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double()) + beta * C0.double()
    res, errs = [], []
    for tc in (0, 1, 3):
        C = C0.clone()
        ops.gemm(A, B, C, transA=bool(ta), transB=bool(tb), beta=float(beta), tc=tc)
        errs.append(float((C.double() - ref).abs().max()))
        for _ in range(5):
            ops.gemm(A, B, C, transA=bool(ta), transB=bool(tb), beta=0.0, tc=tc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20):
                ops.gemm(A, B, C, transA=bool(ta), transB=bool(tb), beta=0.0, tc=tc)
        gr.replay()
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 1e3 / 20)
    print(f"{M:5d} {N:5d} {K:5d} {ta:2d} {tb:2d} | {res[0]:8.1f} {res[1]:8.1f} {res[2]:8.1f} | {errs[0]:.2e} {errs[1]:.2e} {errs[2]:.2e}  ({float(ref.abs().max()):.1f})")
