"""hulc_gemm_bf16 (TMA-fed tcgen05 kind::f16, fp32 accumulation) against float64 torch on bf16-rounded operands: every operand layout,
ragged sizes, tile widths 64 / 128 / 256, cluster split-K, the fused epilogue with fp32 and bf16 outputs, and the cast kernels."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    # M, N, K
    (128, 64, 64),
    (128, 128, 256),
    (256, 64, 96),       # K not a multiple of the 64-wide k-block
    (300, 200, 104),     # ragged everywhere
    (2048, 2048, 1120),  # decoder input projection (BN = 256 tiles)
    (2048, 184, 2048),   # heads
    (2048, 384, 128),    # qkv
    (64, 2048, 2048),    # prior / goal MLP layer: cluster split-K
    (64, 1024, 4096),    # posterior state head: split-K
    (32, 32, 128),       # CLIP head
    (2048, 64, 32),      # K smaller than one k-block
]


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")


def _operands(M, N, K, tA, tB, seed):
    g = torch.Generator().manual_seed(seed)
    r8 = lambda n: (n + 7) // 8 * 8  # TMA needs 16-byte aligned rows: pad the leading dimension of the storage
    A = torch.randn((K, r8(M)) if tA else (M, r8(K)), generator=g).cuda().to(torch.bfloat16)[:, : (M if tA else K)]
    B = torch.randn((N, r8(K)) if tB else (K, r8(N)), generator=g).cuda().to(torch.bfloat16)[:, : (K if tB else N)]
    return A, B


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, True), (True, False)])
def test_gemm_bf16_layouts(M, N, K, tA, tB):
    from hulc_b200 import ops

    A, B = _operands(M, N, K, tA, tB, M * 7 + N * 3 + K)
    C = torch.full((M, N), float("nan"), device="cuda")
    Cb = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm_bf16(A, B, C, Cb, transA=tA, transB=tB)
    ref = (A.t() if tA else A).double() @ (B.t() if tB else B).double()
    err = (C.double() - ref).abs().max().item()
    tol = 2e-6 * K**0.5 * 16 + 1e-5 * K  # exact bf16 products, fp32 accumulation (the tensor core truncates when aligning addends)
    assert err < tol, f"max err {err:.3e} (tol {tol:.3e})"
    torch.testing.assert_close(Cb.float(), C.to(torch.bfloat16).float(), rtol=0, atol=0)
    print(f"M={M} N={N} K={K} tA={tA} tB={tB}: max err {err:.3e}")


def test_gemm_bf16_epilogue_and_strides():
    from hulc_b200 import ops

    g = torch.Generator().manual_seed(5)
    M, N, K = 384, 200, 160
    Abig = torch.randn(M, K + 24, generator=g).cuda().to(torch.bfloat16)
    A = Abig[:, 8 : 8 + K]
    B = torch.randn(N, K, generator=g).cuda().to(torch.bfloat16)
    Cbig = torch.randn(M, N + 8, generator=g).cuda()
    Cpad = Cbig.clone()
    C = Cbig[:, 4 : 4 + N]
    C0 = C.clone()
    Cb = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(N, generator=g).cuda()
    addend = torch.randn(32, N, generator=g).cuda()
    gate = torch.randn(M, N, generator=g).cuda()
    keep = (torch.rand(M, N, generator=g) > 0.3).to(torch.uint8).cuda()
    for gt in (gate, gate.to(torch.bfloat16)):
        C.copy_(C0)
        ops.gemm_bf16(A, B, C, Cb, transB=True, alpha=0.5, beta=2.0, bias=bias, addend=addend, add_mod=32, act=1, gate=gt, drop=ops.Drop(0.3, keep=keep))
        v = 0.5 * (A.double() @ B.double().t()) + bias.double() + addend.double()[torch.arange(M).cuda() % 32] + 2.0 * C0.double()
        v = v.relu()
        v = torch.where(gate > 0, v, torch.zeros_like(v)) * keep.double() / 0.7
        torch.testing.assert_close(C.double(), v, rtol=1e-4, atol=5e-4)
        torch.testing.assert_close(Cb.float(), C.to(torch.bfloat16).float(), rtol=0, atol=0)
        assert torch.equal(Cbig[:, :4], Cpad[:, :4]) and torch.equal(Cbig[:, 4 + N :], Cpad[:, 4 + N :])
    # tanh and the tanh' gate; bf16-only output
    Cb.zero_()
    ops.gemm_bf16(A, B, None, Cb, transB=True, act=2 | 4, gate=gate)
    v = torch.tanh(A.double() @ B.double().t()) * (1 - gate.double() ** 2)
    torch.testing.assert_close(Cb.double(), v, rtol=1e-2, atol=1e-2)


def test_gemm_bf16_philox_dropout_matches_the_fp32_kernels():
    """The in-epilogue Philox dropout draws the same keep decisions as hulc_gemm's dropout pass (forward and backward regenerate them)."""
    from hulc_b200 import ops

    M, N, K = 256, 512, 128
    A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    d = ops.Drop(0.25, seed=1234, site=7)
    ops.set_rng_offset(None)
    ref = ops.gemm(A, B, transB=True, drop=d, tc=0)
    got = ops.gemm_bf16(A.to(torch.bfloat16), B.to(torch.bfloat16), torch.empty(M, N, device="cuda"), transB=True, drop=d)
    assert torch.equal(ref == 0, got == 0)
    assert 0.2 < float((got == 0).float().mean()) < 0.3


def test_gemm_bf16_repeatable_and_rejects_unaligned():
    from hulc_b200 import _lib, ops

    A, B = torch.randn(64, 2048, device="cuda").to(torch.bfloat16), torch.randn(768, 2048, device="cuda").to(torch.bfloat16)
    C1 = ops.gemm_bf16(A, B, torch.empty(64, 768, device="cuda"), transB=True)
    C2 = ops.gemm_bf16(A, B, torch.empty(64, 768, device="cuda"), transB=True)
    assert torch.equal(C1, C2)  # cluster split-K is reduced in a fixed order
    bad = torch.randn(64, 2052, device="cuda").to(torch.bfloat16)[:, :2048]  # rows not 16-byte aligned
    assert not ops.gemm_bf16_ok(bad, B)
    with pytest.raises(_lib.HulcError):
        ops.gemm_bf16(bad, B, torch.empty(64, 768, device="cuda"), transB=True)


def test_cast_bf16():
    from hulc_b200 import ops

    x = torch.randn(1000, 77, device="cuda")
    torch.testing.assert_close(ops.cast_bf16(x), x.to(torch.bfloat16), rtol=0, atol=0)
    v = x[:, 5:70]  # a view: rows with a leading dimension
    out = torch.zeros(1000, 80, device="cuda", dtype=torch.bfloat16)
    ops.cast_bf16(v, out[:, 8:73])
    torch.testing.assert_close(out[:, 8:73], v.to(torch.bfloat16), rtol=0, atol=0)
    assert float(out[:, :8].abs().sum()) == 0 and float(out[:, 73:].abs().sum()) == 0
