"""Tensor-core (tcgen05, tf32) channels-last convolutions against torch fp32/fp64 on the B200: forward, data gradient
(with the ReLU mask) and weight gradient for the three layer shapes of the perceptual encoders, ragged sizes included.
Tolerances are those of tf32 operands (10-bit mantissa, truncated) with fp32 accumulation."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# cin, cout, ks, stride, hw, n
CASES = [(3, 32, 8, 4, 200, 5), (3, 32, 8, 4, 84, 7), (3, 32, 8, 4, 36, 3), (3, 32, 8, 4, 100, 2), (3, 32, 8, 4, 84, 200), (3, 32, 8, 4, 200, 40), (32, 64, 4, 2, 49, 300), (64, 64, 3, 1, 7, 37), (32, 64, 4, 2, 49, 5), (32, 64, 4, 2, 20, 9), (64, 64, 3, 1, 23, 5), (64, 64, 3, 1, 9, 11),
         (32, 64, 4, 2, 12, 1), (64, 64, 3, 1, 5, 2)]


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("cin,cout,ks,st,hw,n", CASES)
def test_conv_tc(cin, cout, ks, st, hw, n):
    from hulc_b200 import ops

    g = torch.Generator().manual_seed(cin * 100 + hw)
    x = torch.randn(n, cin, hw, hw, generator=g).cuda()
    if cin != 3:
        x = x.relu()
    w = (torch.randn(cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks)).cuda()
    b = torch.randn(cout, generator=g).cuda()
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    ref = F.relu(F.conv2d(xr, wr, b.double(), stride=st))
    ho = ref.shape[-1]
    xin = x if cin == 3 else nhwc(x)
    bits = torch.zeros(n, ho, ho, cout // 32, dtype=torch.int32, device="cuda")
    y = ops.conv2d_tc_fwd(xin, w, b, st, torch.empty(n, ho, ho, cout, device="cuda"), relu_bits=bits)
    # the sign mask written next to the output: bit c % 32 of word [pixel][c / 32] = (y > 0)
    want = ((y > 0).view(n, ho, ho, cout // 32, 32).to(torch.int64) << torch.arange(32, device="cuda")).sum(-1)
    assert torch.equal(bits.to(torch.int64) & 0xFFFFFFFF, want)
    scale = float(ref.abs().max())
    err = float((y.double() - nhwc(ref.detach())).abs().max())
    assert err < 4e-3 * scale, f"fwd err {err:.3e} (scale {scale:.2f})"
    dy = (torch.randn(ref.shape, generator=g).cuda() * (ref.detach() > 0)).float()
    ref.backward(dy.double())
    dw = torch.zeros_like(w)
    db = torch.ones(cout, device="cuda")
    ops.conv2d_tc_wgrad(xin, nhwc(dy), dw, st, db=db)  # db accumulates the bias gradient (sum of dy over pixels)
    dbref = dy.double().sum((0, 2, 3)) + 1
    assert float((db.double() - dbref).abs().max()) < 4e-3 * float(dbref.abs().max() + dy.abs().sum((0, 2, 3)).max() * 1e-3), "bias gradient"
    gs = float(wr.grad.abs().max())
    err = float((dw.double() - wr.grad).abs().max())
    assert err < 4e-3 * gs, f"wgrad err {err:.3e} (scale {gs:.2f})"
    ops.conv2d_tc_wgrad(xin, nhwc(dy), dw, st, beta=1.0)
    err = float((dw.double() - 2 * wr.grad).abs().max())
    assert err < 8e-3 * gs
    if cin != 3:
        dx = ops.conv2d_tc_dgrad(nhwc(dy), w, torch.empty(n, hw, hw, cin, device="cuda"), st, gate=xin)
        gbits = ((xin > 0).view(n, hw, hw, cin // 32, 32).to(torch.int64) << torch.arange(32, device="cuda")).sum(-1)
        gbits = torch.where(gbits >= 2 ** 31, gbits - 2 ** 32, gbits).to(torch.int32)
        dxb = ops.conv2d_tc_dgrad(nhwc(dy), w, torch.empty(n, hw, hw, cin, device="cuda"), st, gate=xin, gate_bits=gbits)
        assert torch.equal(dxb, dx), "gating on the sign mask must equal gating on the activation"
        refdx = nhwc(xr.grad * (x > 0))
        ds = float(refdx.abs().max())
        err = float((dx.double() - refdx).abs().max())
        assert err < 4e-3 * ds, f"dgrad err {err:.3e} (scale {ds:.2f})"
        dx2 = ops.conv2d_tc_dgrad(nhwc(dy), w, torch.empty(n, hw, hw, cin, device="cuda"), st)
        err = float((dx2.double() - nhwc(xr.grad)).abs().max())
        assert err < 4e-3 * ds


def test_conv1_rejects_unaligned_width():
    """The tensor-core first layer reads 16-byte pieces of the NCHW rows: a width that is not a multiple of 4 is refused with an error code
    (no launch, no sticky CUDA error) instead of faulting."""
    from hulc_b200 import _lib, ops

    x = torch.randn(2, 3, 50, 50, device="cuda")
    w = torch.randn(32, 3, 8, 8, device="cuda")
    with pytest.raises(_lib.HulcError):
        ops.conv2d_tc_fwd(x, w, torch.zeros(32, device="cuda"), 4, torch.empty(2, 11, 11, 32, device="cuda"))
    with pytest.raises(_lib.HulcError):
        ops.conv2d_tc_wgrad(x, torch.randn(2, 11, 11, 32, device="cuda"), torch.zeros_like(w), 4)
    torch.cuda.synchronize()  # nothing was launched, the context is healthy
    assert float(torch.ones(4, device="cuda").sum()) == 4.0


def test_conv1_band_kernel_fallback():
    """HULC_B200_CONV1_VIEW=0 (read when the library loads, hence the fresh interpreter) routes conv1's forward through the band-staging kernel
    instead of the raw-row view kernel: same result within the tf32 tolerance, same sign mask."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    code = """
import torch, torch.nn.functional as F
from hulc_b200 import ops
g = torch.Generator().manual_seed(5)
for n, hw in ((3, 200), (40, 84)):
    x = torch.randn(n, 3, hw, hw, generator=g).cuda(); w = (torch.randn(32, 3, 8, 8, generator=g) / 14).cuda(); b = torch.randn(32, generator=g).cuda()
    ho = (hw - 8) // 4 + 1
    bits = torch.zeros(n, ho, ho, 1, dtype=torch.int32, device='cuda')
    y = ops.conv2d_tc_fwd(x, w, b, 4, torch.empty(n, ho, ho, 32, device='cuda'), relu_bits=bits)
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), stride=4)).permute(0, 2, 3, 1)
    assert float((y.double() - ref).abs().max()) < 4e-3 * float(ref.abs().max())
    want = ((y > 0).to(torch.int64) << torch.arange(32, device='cuda')).sum(-1)
    assert torch.equal(bits[..., 0].to(torch.int64) & 0xFFFFFFFF, want)
print('ok')
"""
    env = dict(os.environ, HULC_B200_CONV1_VIEW="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=str(Path(__file__).resolve().parent.parent), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
