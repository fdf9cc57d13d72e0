"""Which dense products of one full-size step go to which GEMM path (tensor-core 1-pass / 3-pass / CUDA cores), with timings.
    python scripts/dbg_gemm_log.py"""
import collections
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from hulc_b200 import engine as E, ops  # noqa: E402
from hulc_b200.utils import synthetic  # noqa: E402

log = []
real = ops.gemm


def spy(A, B, C=None, *, transA=False, transB=False, tc=0, **kw):
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N = B.shape[0] if transB else B.shape[1]
    eff = tc if (tc and ops._tc_ok(A, B, M, N, K, transA, transB)) else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = real(A, B, C, transA=transA, transB=transB, tc=tc, **kw)
    e1.record()
    log.append(((M, N, K, int(transA), int(transB), tc, eff, A.stride(0), B.stride(0)), e0, e1))
    return r


ops.gemm = spy
E.gemm = spy
eng = E.HulcEngine("hulc", "rnn_decoder", device="cuda", dropout_p=0.1)
eng.load_state_dict(synthetic.make_state_dict("hulc", "rnn_decoder"))
batch = synthetic.make_batch(32, 32, seed=1, device="cuda")
for i in range(3):
    log.clear()
    eng.step(batch, seed=i)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for key, e0, e1 in log:
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1) * 1e3
tot = collections.Counter()
print(f"{'M':>6} {'N':>6} {'K':>6} tA tB want got   lda   ldb   n   us/call")
for (M, N, K, ta, tb, tc, eff, lda, ldb), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{M:6d} {N:6d} {K:6d} {ta:2d} {tb:2d} {tc:4d} {eff:3d} {lda:5d} {ldb:5d} {n:3d} {us / n:8.1f}")
    tot[eff] += us
print({k: round(v) for k, v in tot.items()}, "us per step by path (eager, includes launch gaps)")
