"""Timing of hulc_gemm (CUDA cores) vs hulc_gemm_tc (tcgen05) on the step's big GEMM shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from hulc_b200 import ops

def t(fn, n=20):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (M, N, K, tA, tB) in [(64, 2048, 2048, 0, 1), (64, 2048, 2048, 0, 0), (64, 6144, 2048, 0, 1), (64, 2048, 160, 0, 1), (2048, 2048, 1120, 0, 1), (2048, 2048, 2048, 0, 1), (2048, 2048, 2048, 0, 0), (2048, 2048, 2048, 1, 0), (8192, 8192, 4096, 0, 1), (2048, 182, 2048, 0, 1), (2048, 4096, 128, 0, 1)]:
    A = torch.randn((K, M) if tA else (M, K), device="cuda"); B = torch.randn((N, K) if tB else (K, N), device="cuda"); C = torch.empty(M, N, device="cuda")
    fl = 2.0 * M * N * K
    r = {}
    for name, kw in (("simt", {}), ("tf32", {"tc": 1}), ("3xtf32", {"tc": 3})):
        if name in ("simt", "3xtf32") and M * N * K > 3e10: continue
        ms = t(lambda: ops.gemm(A, B, C, transA=bool(tA), transB=bool(tB), **kw))
        r[name] = f"{ms*1e3:8.1f}us {fl/ms/1e9:7.1f}TF"
    torch.backends.cuda.matmul.allow_tf32 = True
    a = A.t() if tA else A; b = B.t() if tB else B
    ms = t(lambda: torch.matmul(a, b, out=C)); r["cublas_tf32"] = f"{ms*1e3:8.1f}us {fl/ms/1e9:7.1f}TF"
    print(M, N, K, tA, tB, r, flush=True)
