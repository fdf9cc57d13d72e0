import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from hulc_b200 import ops
M, N, K = 64, 2048, 2048
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda"); pre = torch.randn(M, N, device="cuda")
def timeit(fn, n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, act in (("full", 1),):
    for tc in (1, 3):
        print(name, "tc", tc, f"{timeit(lambda: ops.gemm(A, B, C, transB=True, addend=pre, act=act, tc=tc)):.1f} us")
print("simt", f"{timeit(lambda: ops.gemm(A, B, C, transB=True, addend=pre, act=1)):.1f} us")
torch.backends.cuda.matmul.allow_tf32 = True
print("cublas tf32", f"{timeit(lambda: torch.matmul(A, B.t(), out=C)):.1f} us")
torch.backends.cuda.matmul.allow_tf32 = False
print("cublas fp32", f"{timeit(lambda: torch.matmul(A, B.t(), out=C)):.1f} us")
ref = (A.double() @ B.double().t() + pre.double()).relu()
for tc in (1, 3):
    ops.gemm(A, B, C, transB=True, addend=pre, act=1, tc=tc); torch.cuda.synchronize()
    print("tc", tc, "max err", float((C.double() - ref).abs().max()))
for (M2, N2, K2) in [(2048, 182, 2048), (64, 6144, 2048), (32, 2048, 2048)]:
    A2 = torch.randn(M2, K2, device="cuda"); B2 = torch.randn(N2, K2, device="cuda"); C2 = torch.empty(M2, N2, device="cuda")
    for tc in (0, 1, 3):
        print((M2, N2, K2), "tc", tc, f"{timeit(lambda: ops.gemm(A2, B2, C2, transB=True, tc=tc), 20):.1f} us")
