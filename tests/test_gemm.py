"""hulc_gemm / hulc_colsum against torch: on the host SIMT emulator (CPU suite) and on the B200 (`-m gpu`), via `K`."""
import pytest
import torch


def ref_gemm(A, B, tA, tB, alpha, beta, C0, bias, addend, add_mod, act, gate, keep, p):
    a = A.t() if tA else A
    b = B.t() if tB else B
    v = alpha * (a.double() @ b.double())
    M, N = v.shape
    if bias is not None:
        v = v + bias.double()
    if addend is not None:
        idx = torch.arange(M) % add_mod if add_mod else torch.arange(M)
        v = v + addend.double()[idx]
    if beta:
        v = v + beta * C0.double()
    if act:
        v = v.relu()
    if gate is not None:
        v = torch.where(gate > 0, v, torch.zeros_like(v))
    if keep is not None:
        v = v * keep.double() / (1 - p)
    return v.float()


CASES = [
    # M, N, K, tA, tB
    (64, 64, 64, False, True),
    (70, 50, 33, False, True),
    (5, 130, 257, False, False),
    (130, 7, 40, True, False),
    (300, 200, 100, False, True),  # big tile config
    (260, 100, 19, True, False),
    (64, 128, 2048, False, True),  # split-K
    (33, 60, 1030, False, False),  # split-K, ragged
    (90, 70, 600, True, True),
]


@pytest.mark.parametrize("M,N,Kd,tA,tB", CASES)
def test_gemm_plain(K, M, N, Kd, tA, tB):
    g = torch.Generator().manual_seed(M * 131 + N * 7 + Kd)
    A = torch.randn((Kd, M) if tA else (M, Kd), generator=g)
    B = torch.randn((N, Kd) if tB else (Kd, N), generator=g)
    C = K.gemm(A, B, transA=tA, transB=tB)
    ref = ref_gemm(A, B, tA, tB, 1.0, 0.0, None, None, None, 0, 0, None, None, 0)
    torch.testing.assert_close(C, ref, rtol=1e-4, atol=1e-4 * Kd**0.5)


def test_gemm_epilogue_and_strides(K):
    g = torch.Generator().manual_seed(5)
    M, N, Kd = 96, 72, 200
    Abig = torch.randn(M, Kd + 24, generator=g)
    A = Abig[:, 8 : 8 + Kd]  # unaligned-by-row view with a leading dimension
    B = torch.randn(N, Kd, generator=g)
    Cbig = torch.randn(M, N + 8, generator=g)
    Cpad = Cbig.clone()
    C = Cbig[:, 4 : 4 + N]
    C0 = C.clone()
    bias = torch.randn(N, generator=g)
    addend = torch.randn(32, N, generator=g)
    gate = torch.randn(M, N, generator=g)
    keep = (torch.rand(M, N, generator=g) > 0.3).to(torch.uint8)
    K.gemm(A, B, C, transB=True, alpha=0.5, beta=2.0, bias=bias, addend=addend, add_mod=32, act=1, gate=gate,
             drop=K.Drop(0.3, keep=keep))
    ref = ref_gemm(A, B, False, True, 0.5, 2.0, C0, bias, addend, 32, 1, gate, keep, 0.3)
    torch.testing.assert_close(C, ref, rtol=1e-4, atol=1e-3)
    # padding columns of the strided output are untouched
    assert torch.equal(Cbig[:, :4], Cpad[:, :4]) and torch.equal(Cbig[:, 4 + N :], Cpad[:, 4 + N :])


def test_gemm_philox_dropout_rate_and_determinism(K):
    A = torch.ones(128, 16)
    B = torch.ones(16, 256)
    d = K.Drop(0.25, seed=123, site=7)
    C1 = K.gemm(A, B, drop=d)
    C2 = K.gemm(A, B, drop=d)
    assert torch.equal(C1, C2)
    kept = (C1 != 0).float().mean().item()
    assert abs(kept - 0.75) < 0.02
    torch.testing.assert_close(C1[C1 != 0], torch.full_like(C1[C1 != 0], 16 / 0.75))
    C3 = K.gemm(A, B, drop=K.Drop(0.25, seed=124, site=7))
    assert not torch.equal(C1, C3)


def test_colsum(K):
    X = torch.randn(5000, 70)
    out = K.colsum(X)
    torch.testing.assert_close(out, X.sum(0), rtol=1e-4, atol=1e-3)
    out2 = torch.ones(70)
    K.colsum(X[:100], out2, beta=1.0)
    torch.testing.assert_close(out2, 1 + X[:100].sum(0), rtol=1e-4, atol=1e-4)
