"""Thin tensor-level wrappers over the C ABI (include/hulc_b200.h).  Plumbing only: argument checks, leading dimensions,
workspace and stream lookup.  Every wrapper launches CUDA kernels from libhulc_b200.so on the current stream; nothing
here computes on the host and there is no fallback path."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib

_DEVICE_TYPE = "cuda"  # tensors must live here; (tests/emu swaps the library AND this for the host-emulated build)
_WORKSPACE_BYTES = 96 << 20
_workspaces: Dict[Tuple, torch.Tensor] = {}


class Drop:
    """Dropout description: p, and either an injected uint8 keep-mask or a (seed, site) pair for the Philox stream."""

    __slots__ = ("p", "seed", "site", "keep")

    def __init__(self, p: float = 0.0, seed: int = 0, site: int = 0, keep: Optional[torch.Tensor] = None):
        self.p, self.seed, self.site, self.keep = float(p), int(seed), int(site), keep
        if keep is not None:
            assert keep.dtype == torch.uint8 and keep.is_contiguous()

    def args(self):
        return (self.p, self.seed, self.site, _ptr(self.keep))


NO_DROP = Drop()


def _L():
    return _lib.lib()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    if _DEVICE_TYPE != "cuda":
        return None
    return torch.cuda.current_stream().cuda_stream


def _chk(*ts, dtype=torch.float32):
    for t in ts:
        if t is None:
            continue
        if t.device.type != _DEVICE_TYPE:
            raise _lib.HulcError(f"hulc_b200 kernels need {_DEVICE_TYPE} tensors, got {t.device}")
        if dtype is not None and t.dtype != dtype:
            raise TypeError(f"expected {dtype}, got {t.dtype}")


def _rowmajor(t: torch.Tensor) -> int:
    """leading dimension of a 2-D view whose rows are contiguous"""
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), f"rows must be contiguous, strides {t.stride()}"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def workspace(device) -> torch.Tensor:
    """Zero-initialised scratch for split-K partials and counters, one per (device, stream)."""
    key = (str(device), _stream())
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(_WORKSPACE_BYTES // 4, dtype=torch.float32, device=device)
        _workspaces[key] = ws
    return ws


def empty(*shape, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(*shape, dtype=torch.float32, device=like.device)


def zeros(*shape, like: torch.Tensor) -> torch.Tensor:
    return torch.zeros(*shape, dtype=torch.float32, device=like.device)


# ----------------------------------------------------------------------------------------------------------------------
def gemm(A, B, C=None, *, transA=False, transB=False, alpha=1.0, beta=0.0, bias=None, addend=None, add_mod=0, act=0,
         gate=None, drop: Drop = NO_DROP):
    """C = epi(alpha * op(A) @ op(B)); see hulc_gemm in include/hulc_b200.h.  A, B, C, addend, gate are 2-D views with
    contiguous rows (arbitrary leading dimension)."""
    _chk(A, B, C, bias, addend, gate)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N = B.shape[0] if transB else B.shape[1]
    assert (B.shape[1] if transB else B.shape[0]) == K, (A.shape, B.shape, transA, transB)
    if C is None:
        C = empty(M, N, like=A)
    assert C.shape == (M, N)
    ws = workspace(A.device)
    _L().hulc_gemm(
        _ptr(A), _ptr(B), _ptr(C), M, N, K, _rowmajor(A), _rowmajor(B), _rowmajor(C), int(transA), int(transB),
        float(alpha), float(beta), _ptr(bias), _ptr(addend), _rowmajor(addend) if addend is not None else 0, int(add_mod),
        int(act), _ptr(gate), _rowmajor(gate) if gate is not None else 0, *drop.args(), _ptr(ws), ws.numel() * 4, _stream(),
    )
    return C


def colsum(X, out=None, beta=0.0):
    _chk(X, out)
    rows, cols = X.shape
    if out is None:
        out = empty(cols, like=X)
    _L().hulc_colsum(_ptr(X), rows, cols, _rowmajor(X), _ptr(out), float(beta), _stream())
    return out
