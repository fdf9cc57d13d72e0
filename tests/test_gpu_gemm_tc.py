"""hulc_gemm_tc (tcgen05 / TMEM) against float64 torch on the B200: every operand layout, ragged sizes, the fused
epilogue, 1-pass tf32 (tolerance of a 10-bit mantissa) and 3xTF32 (fp32-level accuracy)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    # M, N, K
    (128, 64, 32),
    (128, 128, 256),
    (256, 64, 96),
    (300, 200, 100),     # ragged everywhere
    (2048, 2048, 1120),  # decoder input projection
    (2048, 182, 2048),   # heads
    (2048, 384, 128),    # qkv
    (131, 70, 37),       # K not a multiple of 4 (scalar loads)
]


def _ref(A, B, tA, tB):
    a = A.t() if tA else A
    b = B.t() if tB else B
    return a.double() @ b.double()


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, True), (True, False)])
@pytest.mark.parametrize("passes", [1, 3])
def test_gemm_tc_layouts(M, N, K, tA, tB, passes):
    from hulc_b200 import ops

    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g).cuda()
    B = torch.randn((N, K) if tB else (K, N), generator=g).cuda()
    C = ops.gemm(A, B, transA=tA, transB=tB, tc=passes)
    ref = _ref(A, B, tA, tB)
    err = (C.double() - ref).abs().max().item()
    scale = K**0.5
    tol = (5e-3 if passes == 1 else 1e-5 + 2e-8 * K) * scale  # 1 pass: operands truncated to tf32; 3 passes: the tensor
    # core accumulates fp32 with truncation, a bias that grows with the number of k-steps  # 3xTF32: the tensor core accumulates with truncation, ~1e-5 relative
    assert err < tol, f"max err {err:.3e} (tol {tol:.3e})"
    print(f"M={M} N={N} K={K} tA={tA} tB={tB} passes={passes}: max err {err:.3e} ({err / scale:.2e} x sqrt(K))")


def test_gemm_tc_epilogue_and_strides():
    from hulc_b200 import ops

    g = torch.Generator().manual_seed(5)
    M, N, K = 384, 200, 160
    Abig = torch.randn(M, K + 24, generator=g).cuda()
    A = Abig[:, 8 : 8 + K]
    B = torch.randn(N, K, generator=g).cuda()
    Cbig = torch.randn(M, N + 8, generator=g).cuda()
    Cpad = Cbig.clone()
    C = Cbig[:, 4 : 4 + N]
    C0 = C.clone()
    bias = torch.randn(N, generator=g).cuda()
    addend = torch.randn(32, N, generator=g).cuda()
    gate = torch.randn(M, N, generator=g).cuda()
    keep = (torch.rand(M, N, generator=g) > 0.3).to(torch.uint8).cuda()
    ops.gemm(A, B, C, transB=True, alpha=0.5, beta=2.0, bias=bias, addend=addend, add_mod=32, act=1, gate=gate, drop=ops.Drop(0.3, keep=keep), tc=3)
    v = 0.5 * (A.double() @ B.double().t()) + bias.double() + addend.double()[torch.arange(M).cuda() % 32] + 2.0 * C0.double()
    v = v.relu()
    v = torch.where(gate > 0, v, torch.zeros_like(v)) * keep.double() / 0.7
    torch.testing.assert_close(C.double(), v, rtol=1e-4, atol=5e-4)
    assert torch.equal(Cbig[:, :4], Cpad[:, :4]) and torch.equal(Cbig[:, 4 + N :], Cpad[:, 4 + N :])


def test_gemm_tc_repeatable():
    from hulc_b200 import ops

    A, B = torch.randn(1024, 512, device="cuda"), torch.randn(768, 512, device="cuda")
    C1 = ops.gemm(A, B, transB=True, tc=3)
    C2 = ops.gemm(A, B, transB=True, tc=3)
    assert torch.equal(C1, C2)


@pytest.mark.parametrize("passes,tB", [(3, True), (1, False), (1, True)])
def test_gemm_tc_skinny_split_k(passes, tB):
    """The recurrent step shape (64 sequences x 2048 hidden): split along K, reduced by the last CTA of each tile."""
    from hulc_b200 import ops

    g = torch.Generator().manual_seed(11)
    M, N, K = 64, 2048, 2048
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K**0.5).cuda()
    B = W if tB else W.t().contiguous()
    pre = torch.randn(M, N, generator=g).cuda()
    gate = torch.randn(M, N, generator=g).cuda()
    ref = torch.where(gate.double() > 0, (A.double() @ W.double().t() + pre.double()).relu(), torch.zeros((), dtype=torch.float64, device="cuda"))
    kw = {}
    C = ops.gemm(A, B, transB=tB, addend=pre, act=1, gate=gate, tc=passes, **kw)
    err = (C.double() - ref).abs().max().item()
    assert err < (2e-5 if passes == 3 else 2e-2), err
    C2 = ops.gemm(A, B, transB=tB, addend=pre, act=1, gate=gate, tc=passes, **kw)
    assert torch.equal(C, C2)
