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
    """Every tensor argument must live on the device the kernels are about to be launched on (the CURRENT device, index included:
    a module built on cuda:0 and moved to cuda:1 must not hand the kernels a stale cuda:0 pointer) and have the expected dtype."""
    cur = torch.cuda.current_device() if _DEVICE_TYPE == "cuda" else None
    for t in ts:
        if t is None:
            continue
        if t.device.type != _DEVICE_TYPE or (cur is not None and t.device.index != cur):
            want = _DEVICE_TYPE if cur is None else f"{_DEVICE_TYPE}:{cur} (the current device)"
            raise _lib.HulcError(f"hulc_b200 kernels need tensors on {want}, got {t.device}")
        if dtype is not None and t.dtype != dtype:
            raise TypeError(f"expected {dtype}, got {t.dtype}")


def _rowmajor(t: torch.Tensor) -> int:
    """leading dimension of a 2-D view whose rows are contiguous"""
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), f"rows must be contiguous, strides {t.stride()}"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def workspace(device) -> torch.Tensor:
    """Zero-initialised scratch for split-K partials and tickets, one per device (calls are issued on one stream at a time;
    keying it by stream would re-allocate — and re-zero on every replay — inside a CUDA-graph capture)."""
    key = str(device)
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
         gate=None, drop: Drop = NO_DROP, tc: int = 0):
    """C = epi(alpha * op(A) @ op(B)); see hulc_gemm in include/hulc_b200.h.  A, B, C, addend, gate are 2-D views with
    contiguous rows (arbitrary leading dimension).  tc = 0: exact-fp32 CUDA-core kernel; tc = 1 / 3: tensor cores
    (hulc_gemm_tc) with tf32 operands / 3xTF32 split products."""
    _chk(A, B, C, bias, addend, gate)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N = B.shape[0] if transB else B.shape[1]
    assert (B.shape[1] if transB else B.shape[0]) == K, (A.shape, B.shape, transA, transB)
    if C is None:
        C = empty(M, N, like=A)
    assert C.shape == (M, N)
    if tc and not _tc_ok(A, B, M, N, K, transA, transB):
        tc = 0  # shapes the 16-byte cp.async producers cannot take go to the CUDA-core kernel
    if tc:
        ws = workspace(A.device)
        _L().hulc_gemm_tc(
            _ptr(A), _ptr(B), _ptr(C), M, N, K, _rowmajor(A), _rowmajor(B), _rowmajor(C), int(transA), int(transB), float(alpha),
            float(beta), _ptr(bias), _ptr(addend), _rowmajor(addend) if addend is not None else 0, int(add_mod), int(act), _ptr(gate),
            _rowmajor(gate) if gate is not None else 0, *drop.args(), int(tc), _ptr(ws), ws.numel() * 4, _stream(),
        )
        return C
    ws = workspace(A.device)
    _L().hulc_gemm(
        _ptr(A), _ptr(B), _ptr(C), M, N, K, _rowmajor(A), _rowmajor(B), _rowmajor(C), int(transA), int(transB),
        float(alpha), float(beta), _ptr(bias), _ptr(addend), _rowmajor(addend) if addend is not None else 0, int(add_mod),
        int(act), _ptr(gate), _rowmajor(gate) if gate is not None else 0, *drop.args(), _ptr(ws), ws.numel() * 4, _stream(),
    )
    return C


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = bf16(x) (hulc_cast_bf16 / hulc_cast_bf16_rows): the narrowing torch.autocast applies to the inputs of nn.Linear / nn.Conv2d."""
    _chk(x)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _chk(out, dtype=torch.bfloat16)
    assert out.shape == x.shape
    if x.is_contiguous() and out.is_contiguous() and x.data_ptr() % 32 == 0 and out.data_ptr() % 16 == 0:
        _L().hulc_cast_bf16(_ptr(x), _ptr(out), x.numel(), _stream())
    else:
        x2, o2 = (x, out) if x.dim() == 2 else (x.reshape(-1, x.shape[-1]) if x.dim() > 2 else x.view(1, -1), out.reshape(-1, out.shape[-1]) if out.dim() > 2 else out.view(1, -1))
        assert x2.data_ptr() == x.data_ptr() and o2.data_ptr() == out.data_ptr(), "views that cannot be flattened to rows need a contiguous copy first"
        _L().hulc_cast_bf16_rows(_ptr(x2), _rowmajor(x2), _ptr(o2), _rowmajor(o2), x2.shape[0], x2.shape[1], _stream())
    return out


def gemm_bf16_ok(A, B) -> bool:
    """Operand requirements of the TMA-fed bf16 kernel: 16-byte aligned bases and rows."""
    return A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0 and _rowmajor(A) % 8 == 0 and _rowmajor(B) % 8 == 0


def gemm_bf16(A, B, C=None, Cb=None, *, transA=False, transB=False, alpha=1.0, beta=0.0, bias=None, addend=None, add_mod=0, act=0, gate=None,
              drop: Drop = NO_DROP):
    """C (fp32) and / or Cb (bf16) = epi(alpha * op(A) @ op(B)) with bf16 operands on the tensor cores (hulc_gemm_bf16); the epilogue
    contract of `gemm`.  `gate` may be fp32 or bf16.  At least one of C / Cb must be given."""
    _chk(A, B, dtype=torch.bfloat16)
    _chk(C, bias, addend)
    _chk(Cb, dtype=torch.bfloat16)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N = B.shape[0] if transB else B.shape[1]
    assert (B.shape[1] if transB else B.shape[0]) == K, (A.shape, B.shape, transA, transB)
    assert C is not None or Cb is not None
    assert (C is None or C.shape == (M, N)) and (Cb is None or Cb.shape == (M, N))
    g32 = g16 = None
    if gate is not None:
        _chk(gate, dtype=None)
        g32, g16 = (gate, None) if gate.dtype == torch.float32 else (None, gate)
        assert gate.dtype in (torch.float32, torch.bfloat16)
    _L().hulc_gemm_bf16(
        _ptr(A), _ptr(B), _ptr(C), _ptr(Cb), M, N, K, _rowmajor(A), _rowmajor(B), _rowmajor(C) if C is not None else 0, _rowmajor(Cb) if Cb is not None else 0,
        int(transA), int(transB), float(alpha), float(beta), _ptr(bias), _ptr(addend), _rowmajor(addend) if addend is not None else 0, int(add_mod), int(act),
        _ptr(g32), _ptr(g16), _rowmajor(gate) if gate is not None else 0, *drop.args(), _stream(),
    )
    return C if C is not None else Cb


def _tc_ok(A, B, M, N, K, transA, transB) -> bool:
    lda, ldb = _rowmajor(A), _rowmajor(B)
    return (A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0 and lda % 4 == 0 and ldb % 4 == 0
            and (M if transA else K) % 4 == 0 and (K if transB else N) % 4 == 0)


def launch_count() -> int:
    """Kernel launches issued by libhulc_b200.so since it was loaded."""
    import ctypes

    n = ctypes.c_ulonglong(0)
    _L().hulc_launch_count(ctypes.addressof(n))
    return int(n.value)


def colsum(X, out=None, beta=0.0):
    _chk(X, out)
    rows, cols = X.shape
    if out is None:
        out = empty(cols, like=X)
    ws = workspace(X.device)
    _L().hulc_colsum(_ptr(X), rows, cols, _rowmajor(X), _ptr(out), float(beta), _ptr(ws), ws.numel() * 4, _stream())
    return out


_colsum_tables: Dict[tuple, tuple] = {}


def colsum_multi(jobs):
    """jobs: list of (X [rows, cols] view with contiguous rows, out [cols], beta).  One launch for all of them (hulc_colsum_multi); the
    device-side job table is cached per set of pointers (persistent buffers: built once, also valid inside a captured graph)."""
    import struct

    if not jobs:
        return
    rows_, key, blk = [], [], 0
    for X, out, beta in jobs:
        _chk(X, out)
        r, c = X.shape
        assert out.numel() == c and out.is_contiguous()
        bits = struct.unpack("<i", struct.pack("<f", float(beta)))[0]
        rows_.append([X.data_ptr(), out.data_ptr(), r, c, _rowmajor(X), bits, blk, 0])
        key.append((X.data_ptr(), out.data_ptr(), r, c, _rowmajor(X), bits))
        blk += (c + 31) // 32
    key = (str(jobs[0][0].device), tuple(key))
    ent = _colsum_tables.get(key)
    if ent is None:
        if len(_colsum_tables) > 64:
            _colsum_tables.clear()
        ent = _colsum_tables[key] = (torch.tensor(rows_, dtype=torch.int64, device=jobs[0][0].device), blk)
    table, total = ent
    _L().hulc_colsum_multi(_ptr(table), len(jobs), total, _stream())


# ----------------------------------------------------------------------------------------------------------------------
# GRU gates
# ----------------------------------------------------------------------------------------------------------------------
def gru_gates_fwd(gi, gh, hprev, h, saved):
    _chk(gi, gh, hprev, h, saved)
    B, H = h.shape
    _L().hulc_gru_gates_fwd(_ptr(gi), _rowmajor(gi), _ptr(gh), _rowmajor(gh), _ptr(hprev), _rowmajor(hprev) if hprev is not None else 0,
                            _ptr(h), _rowmajor(h), _ptr(saved), B, H, _stream())


def gru_gates_bwd(dh_above, dh_rec, saved, hprev, dgi, dgh, dh_carry):
    _chk(dh_above, dh_rec, saved, hprev, dgi, dgh, dh_carry)
    B, H = dh_carry.shape
    ld = lambda t: _rowmajor(t) if t is not None else 0
    _L().hulc_gru_gates_bwd(_ptr(dh_above), ld(dh_above), _ptr(dh_rec), ld(dh_rec), _ptr(saved), _ptr(hprev), ld(hprev), _ptr(dgi), ld(dgi),
                            _ptr(dgh), ld(dgh), _ptr(dh_carry), ld(dh_carry), B, H, _stream())


def rnn_tc_seq_ok(B: int, H: int) -> bool:
    """Shapes the persistent recurrence kernel takes (hulc_rnn_tc_seq): hidden size 2048, at most 64 sequences."""
    return _DEVICE_TYPE == "cuda" and H == 2048 and B <= 64


def rnn_tc_seq(W, prev0, out0, add0, S: int, *, prev_step: int, out_step: int, add_step: int, gate0=None, gate_step: int = 0, act: int = 0,
               transW: bool = False):
    """All S dependent steps of one Elman layer / direction in one launch (see hulc_rnn_tc_seq in include/hulc_b200.h).
    prev0 / out0 / add0 / gate0 are the [B, H] views of step 0; *_step the element strides from one step to the next."""
    _chk(W, prev0, out0, add0, gate0)
    B, H = out0.shape
    ws = workspace(W.device)
    _L().hulc_rnn_tc_seq(_ptr(W), _rowmajor(W), int(transW), _ptr(prev0), int(prev_step), _rowmajor(prev0), _ptr(out0), int(out_step),
                         _rowmajor(out0), _ptr(add0), int(add_step), _rowmajor(add0), _ptr(gate0), int(gate_step),
                         _rowmajor(gate0) if gate0 is not None else 0, int(act), B, H, int(S), _ptr(ws), ws.numel() * 4, _stream())


_push_clusters = None


def rnn_seq_bf16_ok(B: int, H: int) -> bool:
    """hulc_rnn_seq_bf16 takes hidden size 2048, at most 64 sequences, on a device that holds its 32 four-CTA clusters at once."""
    global _push_clusters
    if not rnn_tc_seq_ok(B, H):
        return False
    if _push_clusters is None:
        import ctypes
        out = (ctypes.c_int * 2)()
        _L().hulc_rnn_push_max_clusters(ctypes.addressof(out))
        _push_clusters = (out[0], out[1])
    return _push_clusters[0] >= 32


def rnn_seq_bf16(W16, prev0, x16, out0, add0, S: int, *, out_step: int, add_step: int, gate0=None, gate_step: int = 0, act: int = 0,
                 transW: bool = False):
    """bf16 variant of `rnn_tc_seq` (hulc_rnn_seq_bf16): W16 is the bf16 copy of weight_hh, x16 a bf16 workspace of (S + 1) * B * H elements
    through which the hidden state travels between the steps; the fp32 result of step s lands in out0 + s * out_step."""
    _chk(prev0, out0, add0, gate0)
    _chk(W16, x16, dtype=torch.bfloat16)
    B, H = prev0.shape
    assert x16.numel() >= (S + 1) * B * H and x16.is_contiguous() and W16.stride(1) == 1
    _L().hulc_rnn_seq_bf16(_ptr(W16), W16.stride(0), int(transW), _ptr(prev0), _rowmajor(prev0), _ptr(x16), _ptr(out0), int(out_step),
                           _rowmajor(out0) if out0 is not None else 0, _ptr(add0), int(add_step), _rowmajor(add0), _ptr(gate0), int(gate_step),
                           _rowmajor(gate0) if gate0 is not None else 0, int(act), B, H, int(S), _stream())


# ----------------------------------------------------------------------------------------------------------------------
# convolutions
# ----------------------------------------------------------------------------------------------------------------------
def _conv_out(n, k, s):
    return (n - k) // s + 1


def conv2d_fwd(x, w, b, stride, y=None, relu=True):
    _chk(x, w, b, y)
    assert x.is_contiguous() and w.is_contiguous()
    N, CIN, H, W = x.shape
    COUT, _, KS, _ = w.shape
    if y is None:
        y = empty(N, COUT, _conv_out(H, KS, stride), _conv_out(W, KS, stride), like=x)
    assert y.is_contiguous()
    ws = workspace(x.device)
    _L().hulc_conv2d_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(y), N, CIN, H, W, COUT, KS, stride, int(relu), _ptr(ws), ws.numel() * 4, _stream())
    return y


def conv2d_dgrad(dy, w, x_shape, stride, gate=None, dx=None):
    _chk(dy, w, gate, dx)
    N, CIN, H, W = x_shape
    COUT, _, KS, _ = w.shape
    if dx is None:
        dx = empty(N, CIN, H, W, like=dy)
    assert dy.is_contiguous() and dx.is_contiguous() and (gate is None or gate.is_contiguous())
    ws = workspace(dy.device)
    _L().hulc_conv2d_dgrad(_ptr(dy), _ptr(w), _ptr(gate), _ptr(dx), N, CIN, H, W, COUT, KS, stride, _ptr(ws), ws.numel() * 4, _stream())
    return dx


def conv2d_wgrad(x, dy, dw, stride, beta=0.0):
    _chk(x, dy, dw)
    N, CIN, H, W = x.shape
    COUT, _, KS, _ = dw.shape
    assert x.is_contiguous() and dy.is_contiguous() and dw.is_contiguous()
    ws = workspace(x.device)
    _L().hulc_conv2d_wgrad(_ptr(x), _ptr(dy), _ptr(dw), float(beta), N, CIN, H, W, COUT, KS, stride, _ptr(ws), ws.numel() * 4, _stream())
    return dw


# channels-last tensor-core variants: activations are NHWC [N,H,W,C]; the first layer reads the NCHW frames directly
def conv2d_tc_fwd(x, w, b, stride, y, relu=True, relu_bits=None):
    _chk(x, w, b, y)
    assert x.is_contiguous() and w.is_contiguous() and y.is_contiguous()
    COUT, CIN, KS, _ = w.shape
    N, H, W = (x.shape[0], x.shape[2], x.shape[3]) if CIN == 3 else (x.shape[0], x.shape[1], x.shape[2])
    assert tuple(y.shape) == (N, _conv_out(H, KS, stride), _conv_out(W, KS, stride), COUT), y.shape
    ws = workspace(x.device)
    _chk(relu_bits, dtype=torch.int32)
    _L().hulc_conv2d_tc_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(y), N, CIN, H, W, COUT, KS, stride, int(relu), _ptr(relu_bits), _ptr(ws), ws.numel() * 4, _stream())
    return y


def conv2d_tc_dgrad(dy, w, dx, stride, gate=None, gate_bits=None):
    _chk(dy, w, gate, dx)
    COUT, CIN, KS, _ = w.shape
    N, H, W, _ = dx.shape
    assert dy.is_contiguous() and dx.is_contiguous() and (gate is None or (gate.is_contiguous() and gate.shape == dx.shape))
    ws = workspace(dy.device)
    _chk(gate_bits, dtype=torch.int32)
    assert gate_bits is None or gate is not None
    _L().hulc_conv2d_tc_dgrad(_ptr(dy), _ptr(w), _ptr(gate), _ptr(gate_bits), _ptr(dx), N, CIN, H, W, COUT, KS, stride, _ptr(ws), ws.numel() * 4, _stream())
    return dx


def conv2d_tc_wgrad(x, dy, dw, stride, beta=0.0, db=None):
    """dw = beta * dw + dL/dw; db (optional) += sum over pixels of dy (the bias gradient, from the same pass where possible)."""
    _chk(x, dy, dw, db)
    COUT, CIN, KS, _ = dw.shape
    nchw = CIN == 3
    N, H, W = (x.shape[0], x.shape[2], x.shape[3]) if nchw else (x.shape[0], x.shape[1], x.shape[2])
    assert x.is_contiguous() and dy.is_contiguous() and dw.is_contiguous()
    ws = workspace(x.device)
    _L().hulc_conv2d_tc_wgrad(_ptr(x), _ptr(dy), _ptr(dw), float(beta), _ptr(db), N, CIN, H, W, COUT, KS, stride, int(nchw), _ptr(ws), ws.numel() * 4, _stream())
    return dw


# bf16 activations between the conv layers (bf16 path): layer 1 reads the fp32 NCHW frames, layers 2 / 3 bf16 NHWC; outputs bf16 NHWC
def conv2d_bf16_fwd(x, w, b, stride, y, relu=True, relu_bits=None):
    _chk(w, b)
    COUT, CIN, KS, _ = w.shape
    _chk(x, dtype=torch.float32 if CIN == 3 else torch.bfloat16)
    _chk(y, dtype=torch.bfloat16)
    _chk(relu_bits, dtype=torch.int32)
    assert x.is_contiguous() and w.is_contiguous() and y.is_contiguous()
    N, H, W = (x.shape[0], x.shape[2], x.shape[3]) if CIN == 3 else (x.shape[0], x.shape[1], x.shape[2])
    assert tuple(y.shape) == (N, _conv_out(H, KS, stride), _conv_out(W, KS, stride), COUT), y.shape
    ws = workspace(x.device)
    _L().hulc_conv2d_bf16_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(y), N, CIN, H, W, COUT, KS, stride, int(relu), _ptr(relu_bits), _ptr(ws), ws.numel() * 4, _stream())
    return y


def conv2d_bf16_dgrad(dy, w, dx, stride, gate_bits):
    _chk(w)
    _chk(dy, dx, dtype=torch.bfloat16)
    _chk(gate_bits, dtype=torch.int32)
    COUT, CIN, KS, _ = w.shape
    N, H, W, _ = dx.shape
    assert dy.is_contiguous() and dx.is_contiguous()
    ws = workspace(dy.device)
    _L().hulc_conv2d_bf16_dgrad(_ptr(dy), _ptr(w), _ptr(gate_bits), _ptr(dx), N, CIN, H, W, COUT, KS, stride, _ptr(ws), ws.numel() * 4, _stream())
    return dx


def conv2d_bf16_wgrad(x, dy, dw, stride, beta=0.0, db=None):
    """dw = beta * dw + dL/dw (fp32); db (optional, fp32) += sum over pixels of dy, out of the same tensor-core pass."""
    _chk(dw, db)
    COUT, CIN, KS, _ = dw.shape
    _chk(x, dtype=torch.float32 if CIN == 3 else torch.bfloat16)
    _chk(dy, dtype=torch.bfloat16)
    N, H, W = (x.shape[0], x.shape[2], x.shape[3]) if CIN == 3 else (x.shape[0], x.shape[1], x.shape[2])
    assert x.is_contiguous() and dy.is_contiguous() and dw.is_contiguous()
    ws = workspace(x.device)
    _L().hulc_conv2d_bf16_wgrad(_ptr(x), _ptr(dy), _ptr(dw), float(beta), _ptr(db), N, CIN, H, W, COUT, KS, stride, _ptr(ws), ws.numel() * 4, _stream())
    return dw


def spatial_softmax_nhwc_bf16_fwd(x, out, temperature=1.0):
    _chk(x, dtype=torch.bfloat16)
    _chk(out)
    N, H, W, C = x.shape
    assert x.is_contiguous() and out.is_contiguous()
    _L().hulc_spatial_softmax_nhwc_bf16_fwd(_ptr(x), _ptr(out), N, C, H, W, 1.0 / float(temperature), _stream())
    return out


def spatial_softmax_nhwc_bf16_bwd(x, dout, dx, temperature=1.0, relu_gate=True):
    _chk(x, dx, dtype=torch.bfloat16)
    _chk(dout)
    N, H, W, C = x.shape
    assert x.is_contiguous() and dout.is_contiguous() and dx.is_contiguous()
    _L().hulc_spatial_softmax_nhwc_bf16_bwd(_ptr(x), _ptr(dout), _ptr(dx), N, C, H, W, 1.0 / float(temperature), int(relu_gate), _stream())
    return dx


def nchw_channel_sum(x, out):
    _chk(x, out)
    N, C = x.shape[:2]
    P = x[0, 0].numel()
    assert x.is_contiguous()
    _L().hulc_nchw_channel_sum(_ptr(x), _ptr(out), N, C, P, _stream())
    return out


def spatial_softmax_fwd(x, out=None, temperature=1.0):
    _chk(x, out)
    N, C, H, W = x.shape
    if out is None:
        out = empty(N, 2 * C, like=x)
    assert x.is_contiguous() and out.is_contiguous()
    _L().hulc_spatial_softmax_fwd(_ptr(x), _ptr(out), N * C, H, W, 1.0 / float(temperature), _stream())
    return out


def spatial_softmax_bwd(x, dout, dx=None, temperature=1.0, relu_gate=True):
    _chk(x, dout, dx)
    N, C, H, W = x.shape
    if dx is None:
        dx = empty(N, C, H, W, like=x)
    assert x.is_contiguous() and dout.is_contiguous() and dx.is_contiguous()
    _L().hulc_spatial_softmax_bwd(_ptr(x), _ptr(dout), _ptr(dx), N * C, H, W, 1.0 / float(temperature), int(relu_gate), _stream())
    return dx


def spatial_softmax_nhwc_fwd(x, out=None, temperature=1.0):
    _chk(x, out)
    N, H, W, C = x.shape
    if out is None:
        out = empty(N, 2 * C, like=x)
    assert x.is_contiguous() and out.is_contiguous()
    _L().hulc_spatial_softmax_nhwc_fwd(_ptr(x), _ptr(out), N, C, H, W, 1.0 / float(temperature), _stream())
    return out


def spatial_softmax_nhwc_bwd(x, dout, dx=None, temperature=1.0, relu_gate=True):
    _chk(x, dout, dx)
    N, H, W, C = x.shape
    if dx is None:
        dx = empty(N, H, W, C, like=x)
    assert x.is_contiguous() and dout.is_contiguous() and dx.is_contiguous()
    _L().hulc_spatial_softmax_nhwc_bwd(_ptr(x), _ptr(dout), _ptr(dx), N, C, H, W, 1.0 / float(temperature), int(relu_gate), _stream())
    return dx


# ----------------------------------------------------------------------------------------------------------------------
# LayerNorm / transformer pieces
# ----------------------------------------------------------------------------------------------------------------------
def layernorm_fwd(x, w, b, y, stats, res=None, z=None, eps=1e-5, drop: Drop = NO_DROP):
    """y = LN(res + drop(x)) (res None: LN(x)); z receives the pre-norm sum when given; stats [rows,2]."""
    _chk(x, w, b, y, stats, res, z)
    rows, D = x.shape
    ld = lambda t: _rowmajor(t) if t is not None else 0
    _L().hulc_layernorm_fwd(_ptr(x), ld(x), _ptr(res), ld(res), _ptr(w), _ptr(b), _ptr(y), ld(y), _ptr(z), ld(z), _ptr(stats), rows, D, float(eps),
                            *drop.args(), _stream())
    return y


def layernorm_bwd(dy, z, stats, w, dw, db, dz=None, dx=None, drop: Drop = NO_DROP):
    """dz = dLN/dz (goes to the residual branch), dx = dz * dropout factor; dw, db are ACCUMULATED into."""
    _chk(dy, z, stats, w, dw, db, dz, dx)
    rows, D = dy.shape
    ld = lambda t: _rowmajor(t) if t is not None else 0
    _L().hulc_layernorm_bwd(_ptr(dy), ld(dy), _ptr(z), ld(z), _ptr(stats), _ptr(w), _ptr(dz), ld(dz), _ptr(dx), ld(dx), _ptr(dw), _ptr(db), rows, D,
                            *drop.args(), _stream())


def add_posemb_fwd(x, pos, y, drop: Drop = NO_DROP):
    _chk(x, pos, y)
    B, S, D = x.shape
    assert x.is_contiguous() and y.is_contiguous() and pos.is_contiguous() and pos.shape[0] >= S
    _L().hulc_add_posemb_fwd(_ptr(x), _ptr(pos), _ptr(y), B, S, D, *drop.args(), _stream())
    return y


def dropout_apply(x, y, drop: Drop):
    _chk(x, y)
    assert x.is_contiguous() and y.is_contiguous()
    _L().hulc_dropout_apply(_ptr(x), _ptr(y), x.numel(), *drop.args(), _stream())
    return y


def attention_fwd(qkv, out, probs, B, S, H, drop: Drop = NO_DROP):
    _chk(qkv, out, probs)
    dh = qkv.shape[1] // (3 * H)
    assert qkv.is_contiguous() and out.is_contiguous() and probs.is_contiguous()
    _L().hulc_attention_fwd(_ptr(qkv), _ptr(out), _ptr(probs), B, S, H, dh, *drop.args(), _stream())
    return out


def attention_bwd(qkv, probs, dout, dqkv, B, S, H, drop: Drop = NO_DROP):
    _chk(qkv, probs, dout, dqkv)
    dh = qkv.shape[1] // (3 * H)
    assert dout.is_contiguous() and dqkv.is_contiguous()
    _L().hulc_attention_bwd(_ptr(qkv), _ptr(probs), _ptr(dout), _ptr(dqkv), B, S, H, dh, *drop.args(), _stream())
    return dqkv


# ----------------------------------------------------------------------------------------------------------------------
# data movement
# ----------------------------------------------------------------------------------------------------------------------
def strided_copy(dst, src, alpha=1.0, accumulate=False):
    """dst (+)= alpha * src for <=3-D views of fp32 storage (src may be an expanded/permuted view)."""
    _chk(dst, src)
    assert dst.shape == src.shape and dst.dim() <= 3
    shape = [1] * (3 - dst.dim()) + list(dst.shape)
    ds = [0] * (3 - dst.dim()) + list(dst.stride())
    ss = [0] * (3 - src.dim()) + list(src.stride())
    _L().hulc_strided_copy(_ptr(dst), _ptr(src), *shape, *ds, *ss, float(alpha), int(accumulate), _stream())
    return dst


def reduce_mid(x, out, scale=1.0):
    _chk(x, out)
    B, S, D = x.shape
    assert x.is_contiguous() and out.is_contiguous()
    _L().hulc_reduce_mid(_ptr(x), _ptr(out), B, S, D, float(scale), _stream())
    return out


def sum_to(x, out, scale=1.0):
    _chk(x, out)
    assert x.is_contiguous()
    _L().hulc_sum(_ptr(x), x.numel(), _ptr(out), float(scale), _stream())
    return out


def frames_u8_to_f32(src, dst, mean=0.5, std=0.5):
    """dst = ((src / 255) - mean) / std for uint8 camera frames (hulc_frames_u8_to_f32)."""
    _chk(src, dtype=torch.uint8)
    _chk(dst)
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    _L().hulc_frames_u8_to_f32(_ptr(src), _ptr(dst), src.numel(), float(mean), float(std), _stream())
    return dst


def frames_u8_shift_to_f32(src, dst, pad, shifts=None, seed=0, site=0, mean=0.5, std=0.5):
    """RandomShiftsAug + scale + normalise of uint8 frames [N, C, H, W] (hulc_frames_u8_shift_to_f32); shifts [N, 2] int32 (sx, sy) or Philox."""
    _chk(src, dtype=torch.uint8)
    _chk(dst)
    _chk(shifts, dtype=torch.int32)
    N, C, H, W = src.shape
    assert src.is_contiguous() and dst.is_contiguous() and tuple(dst.shape) == (N, C, H, W) and (shifts is None or tuple(shifts.shape) == (N, 2))
    _L().hulc_frames_u8_shift_to_f32(_ptr(src), _ptr(dst), N, C, H, W, int(pad), _ptr(shifts), int(seed), int(site), float(mean), float(std), _stream())
    return dst


def scale_(x, alpha):
    _chk(x)
    assert x.is_contiguous()
    _L().hulc_scale(_ptr(x), x.numel(), float(alpha), _stream())
    return x


def scale_dev_(x, alpha_dev):
    """x *= alpha with alpha a 1-element device tensor (no host read; a factor of exactly 1 costs one empty launch)."""
    _chk(x, alpha_dev)
    assert x.is_contiguous() and alpha_dev.numel() == 1
    _L().hulc_scale_dev(_ptr(x), x.numel(), _ptr(alpha_dev), _stream())
    return x


# ----------------------------------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------------------------------
def world_to_tcp(actions, robot_obs, out, nan_flag):
    _chk(actions, robot_obs, out)
    _chk(nan_flag, dtype=torch.int32)
    assert actions.is_contiguous() and robot_obs.is_contiguous() and actions.shape[-1] == 7
    _L().hulc_world_to_tcp(_ptr(actions), _ptr(robot_obs), robot_obs.shape[-1], _ptr(out), actions.numel() // 7, _ptr(nan_flag), _stream())
    return out


def tcp_to_world(actions, robot_obs, out, nan_flag):
    """Sampled actions (TCP frame) -> world frame (hulc_tcp_to_world; gripper_control.py:39-63)."""
    _chk(actions, robot_obs, out)
    _chk(nan_flag, dtype=torch.int32)
    assert actions.is_contiguous() and robot_obs.is_contiguous() and out.is_contiguous() and actions.shape[-1] == 7
    _L().hulc_tcp_to_world(_ptr(actions), _ptr(robot_obs), robot_obs.shape[-1], _ptr(out), actions.numel() // 7, _ptr(nan_flag), _stream())
    return out


def logistic_sample(heads, out, B, S, b0, Bm, *, time_major, n_dims, n_mix, log_scale_min=-7.0, has_gripper=True, grip_bounds=(-1.0, 1.0), u_mix=None,
                    u_inv=None, seed=0, site=0):
    """Actions sampled from the logistic mixture the heads describe (hulc_logistic_sample; logistic_decoder_rnn.py:234-258).
    out: [B, S, n_dims + has_gripper] (only sequences [b0, b0 + Bm) are written); u_mix [Bm, S, n_dims, n_mix] / u_inv [Bm, S, n_dims] inject
    the uniforms, otherwise Philox(seed, site)."""
    _chk(heads, out, u_mix, u_inv)
    assert out.is_contiguous() and (u_mix is None or u_mix.is_contiguous()) and (u_inv is None or u_inv.is_contiguous())
    _L().hulc_logistic_sample(_ptr(heads), _rowmajor(heads), _ptr(u_mix), _ptr(u_inv), _ptr(out), B, S, b0, Bm, int(time_major), n_dims, n_mix,
                              float(log_scale_min), int(has_gripper), float(grip_bounds[0]), float(grip_bounds[1]), int(seed), int(site), _stream())
    return out


def val_metrics(pred, actions, mae, hits):
    """mae [B, n_dims] = mean over time of |pred - actions|; hits [B] = steps with the right gripper command (hulc_val_metrics)."""
    _chk(pred, actions, mae, hits)
    B, S, A = pred.shape
    assert pred.is_contiguous() and actions.is_contiguous() and tuple(actions.shape) == (B, S, A) and mae.is_contiguous()
    _L().hulc_val_metrics(_ptr(pred), _ptr(actions), _ptr(mae), _ptr(hits), B, S, A - 1, _stream())
    return mae, hits


def logistic_loss(heads, actions, dheads, losses, B, S, b0, Bm, *, time_major, n_dims, n_mix, num_classes, log_scale_min=-7.0, act_min=-1.0,
                  act_max=1.0, has_gripper=True, gripper_alpha=1.0, grad_scale=1.0):
    _chk(heads, actions, dheads, losses)
    assert actions.is_contiguous() and _rowmajor(heads) == _rowmajor(dheads)
    ws = workspace(heads.device)
    _L().hulc_logistic_loss(_ptr(heads), _rowmajor(heads), _ptr(actions), actions.shape[-1], _ptr(dheads), _ptr(losses), B, S, b0, Bm,
                            int(time_major), n_dims, n_mix, num_classes, float(log_scale_min), float(act_min), float(act_max), int(has_gripper),
                            float(gripper_alpha), float(grad_scale), _ptr(ws), ws.numel() * 4, _stream())


def plan_discrete_fwd(pr_logit, pp_logit, plan, kl_rows, *, u=None, idx_in=None, idx_out=None, class_size=32, seed=0, site=0):
    _chk(pr_logit, pp_logit, plan, kl_rows, u)
    _chk(idx_in, idx_out, dtype=torch.int32)
    rows = pr_logit.numel() // class_size
    _L().hulc_plan_discrete_fwd(_ptr(pr_logit), _ptr(pp_logit), _ptr(u), _ptr(idx_in), _ptr(plan), _ptr(idx_out), _ptr(kl_rows), rows, class_size,
                                int(seed), int(site), _stream())


def plan_discrete_bwd(pr_logit, pp_logit, dplan, d_pr, d_pp, coef_lhs, coef_rhs, class_size=32, dkl=None):
    _chk(pr_logit, pp_logit, dplan, d_pr, d_pp, dkl)
    rows = pr_logit.numel() // class_size
    _L().hulc_plan_discrete_bwd(_ptr(pr_logit), _ptr(pp_logit), _ptr(dplan), _ptr(dkl), float(coef_lhs), float(coef_rhs), _ptr(d_pr), _ptr(d_pp),
                                rows, class_size, _stream())


def plan_cont_fwd(pr_state, pp_state, plan, kl_elem, *, eps=None, seed=0, site=0):
    _chk(pr_state, pp_state, plan, kl_elem, eps)
    Bn, P = plan.shape
    _L().hulc_plan_cont_fwd(_ptr(pr_state), _ptr(pp_state), _ptr(eps), _ptr(plan), _ptr(kl_elem), Bn, P, int(seed), int(site), _stream())


def plan_cont_bwd(pr_state, pp_state, dplan, d_pr, d_pp, coef_lhs, coef_rhs, *, eps=None, seed=0, site=0, dkl=None):
    _chk(pr_state, pp_state, dplan, d_pr, d_pp, eps, dkl)
    Bn, P2 = pr_state.shape
    _L().hulc_plan_cont_bwd(_ptr(pr_state), _ptr(pp_state), _ptr(eps), _ptr(dplan), _ptr(dkl), float(coef_lhs), float(coef_rhs), _ptr(d_pr),
                            _ptr(d_pp), Bn, P2 // 2, int(seed), int(site), _stream())


def clip_loss(im, tx, logit_scale, mask, loss, d_im, d_tx, d_logit_scale, grad_scale=1.0):
    _chk(im, tx, logit_scale, loss, d_im, d_tx, d_logit_scale)
    _chk(mask, dtype=torch.uint8)
    n, D = im.shape
    assert im.is_contiguous() and tx.is_contiguous() and d_im.is_contiguous() and d_tx.is_contiguous()
    _L().hulc_clip_loss(_ptr(im), _ptr(tx), _ptr(logit_scale), _ptr(mask), _ptr(loss), _ptr(d_im), _ptr(d_tx), _ptr(d_logit_scale), n, D,
                        float(grad_scale), _stream())


def cosine_loss(pred, target, dpred, loss, grad_scale=1.0):
    """loss[0] = mean cosine distance of the rows of pred / target [B, D]; dpred = grad_scale * its gradient (hulc_cosine_loss)."""
    _chk(pred, target, dpred, loss)
    B, D = pred.shape
    _L().hulc_cosine_loss(_ptr(pred), _rowmajor(pred), _ptr(target), _rowmajor(target), _ptr(dpred), _rowmajor(dpred), _ptr(loss), B, D, float(grad_scale), _stream())


def bce_logits_loss(logits, dlogits, loss, n_pos: int, n_neg: int, grad_scale=1.0):
    """Binary cross entropy with logits over n_pos scores labelled 1 followed by n_neg labelled 0 (hulc_bce_logits_loss)."""
    _chk(logits, dlogits, loss)
    assert logits.is_contiguous() and dlogits.is_contiguous() and logits.numel() == n_pos + n_neg
    _L().hulc_bce_logits_loss(_ptr(logits), _ptr(dlogits), _ptr(loss), int(n_pos), int(n_neg), float(grad_scale), _stream())


def set_rng_offset(t: Optional[torch.Tensor]):
    """Point the library's device-resident RNG offset at a 1-element int64 tensor (None clears it)."""
    if t is not None:
        _chk(t, dtype=torch.int64)
    _L().hulc_set_rng_offset_ptr(_ptr(t))


def adam_step(p, g, m, v, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, step=1, grad_scale=1.0, step_dev=None, p_bf16=None):
    _chk(p, g, m, v)
    _chk(step_dev, dtype=torch.int32)
    assert p.is_contiguous() and g.is_contiguous() and m.is_contiguous() and v.is_contiguous()
    if p_bf16 is not None:
        assert p_bf16.dtype in (torch.bfloat16, torch.int16) and p_bf16.numel() == p.numel() and p_bf16.is_contiguous()
        _L().hulc_adam_step_bf16(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), p.numel(), float(lr), float(beta1), float(beta2), float(eps), int(step),
                                 _ptr(step_dev), float(grad_scale), _stream())
        return
    _L().hulc_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps), int(step),
                        _ptr(step_dev), float(grad_scale), _stream())
