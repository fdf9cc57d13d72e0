"""The bf16 convolutions of the perceptual encoders (hulc_conv2d_bf16_{fwd,dgrad,wgrad}, hulc_spatial_softmax_nhwc_bf16_*) against torch
on the same bf16-rounded operands (fp32 arithmetic): both cameras' geometries, ragged frame counts, bias gradient out of the wgrad pass."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

LAYERS = {1: (3, 32, 8, 4), 2: (32, 64, 4, 2), 3: (64, 64, 3, 1)}


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")


def _rel(a, b):
    return float((a.double() - b.double()).abs().max()) / max(float(b.double().abs().max()), 1e-12)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("layer,hw,n", [(1, 200, 3), (1, 84, 5), (2, 49, 7), (2, 20, 9), (3, 23, 7), (3, 9, 33), (2, 49, 300), (3, 23, 300)])
def test_conv_bf16_fwd_dgrad_wgrad(layer, hw, n):
    from hulc_b200 import ops

    cin, cout, ks, st = LAYERS[layer]
    g = torch.Generator().manual_seed(layer * 1000 + hw + n)
    x = torch.randn(n, cin, hw, hw, generator=g).cuda()
    w = (torch.randn(cout, cin, ks, ks, generator=g) / (cin * ks * ks) ** 0.5).cuda()
    b = (0.1 * torch.randn(cout, generator=g)).cuda()
    ho = (hw - ks) // st + 1
    if layer > 1:
        x = x.to(torch.bfloat16).float()  # the kernels see bf16 activations
    wb = w.to(torch.bfloat16).float() if layer > 1 else w  # ... and bf16 weights (layer 1 runs tf32 on the fp32 frames)
    xr = x.clone().requires_grad_(True)
    wr = wb.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = F.relu(F.conv2d(xr, wr, br, stride=st))
    # forward
    xin = x if layer == 1 else _nhwc(x).to(torch.bfloat16)
    y = torch.zeros(n, ho, ho, cout, device="cuda", dtype=torch.bfloat16)
    bits = torch.zeros(n, ho, ho, cout // 32, device="cuda", dtype=torch.int32)
    ops.conv2d_bf16_fwd(xin, w, b, st, y, relu_bits=bits)
    yref = _nhwc(ref.detach())
    assert _rel(y.float(), yref) < (6e-3 if layer > 1 else 1e-2), _rel(y.float(), yref)
    sign = (y.float() > 0).view(n, ho, ho, cout // 32, 32)
    want = (sign.long() << torch.arange(32, device="cuda")).sum(-1)
    assert torch.equal(bits.long() & 0xFFFFFFFF, want)
    # backward operands: dy bf16, gated upstream by this layer's own ReLU (as the engine's callers do)
    dy = (torch.randn(n, cout, ho, ho, generator=g).cuda() * (ref.detach() > 0)).to(torch.bfloat16).float()
    ref.backward(dy)
    dyh = _nhwc(dy).to(torch.bfloat16)
    dw = torch.zeros_like(w)
    db = torch.zeros_like(b)
    ops.conv2d_bf16_wgrad(xin, dyh, dw, st, beta=0.0, db=db)
    assert _rel(dw, wr.grad) < (5e-3 if layer > 1 else 2e-2), ("dw", _rel(dw, wr.grad))
    assert _rel(db, br.grad) < 5e-3, ("db", _rel(db, br.grad))
    ops.conv2d_bf16_wgrad(xin, dyh, dw, st, beta=1.0, db=db)  # accumulate
    assert _rel(dw, 2 * wr.grad) < (5e-3 if layer > 1 else 2e-2) and _rel(db, 2 * br.grad) < 5e-3
    if layer > 1:  # data gradient, gated by the sign bits of the activation that fed the layer
        gate = (torch.rand(n, hw, hw, cin, generator=g) > 0.3).cuda()
        gbits = (gate.view(n, hw, hw, cin // 32, 32).long() << torch.arange(32, device="cuda")).sum(-1).to(torch.int32)
        dx = torch.full((n, hw, hw, cin), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.conv2d_bf16_dgrad(dyh, w, dx, st, gbits)
        want = _nhwc(xr.grad) * gate
        assert _rel(dx.float(), want) < 6e-3, ("dx", _rel(dx.float(), want))


def test_spatial_softmax_bf16():
    from hulc_b200 import ops

    g = torch.Generator().manual_seed(3)
    n, c, h = 5, 64, 21
    x = F.relu(torch.randn(n, h, h, c, generator=g)).cuda().to(torch.bfloat16)
    out = ops.spatial_softmax_nhwc_bf16_fwd(x, torch.empty(n, 2 * c, device="cuda"))
    ref32 = ops.spatial_softmax_nhwc_fwd(x.float(), torch.empty(n, 2 * c, device="cuda"))
    torch.testing.assert_close(out, ref32, rtol=1e-6, atol=1e-6)  # same arithmetic on the same values
    dout = torch.randn(n, 2 * c, generator=g).cuda()
    dx = ops.spatial_softmax_nhwc_bf16_bwd(x, dout, torch.empty_like(x), relu_gate=True)
    dref = ops.spatial_softmax_nhwc_bwd(x.float(), dout, torch.empty(n, h, h, c, device="cuda"), relu_gate=True)
    # the same arithmetic on the same values, summed in a different lane order (8 vs 16 channel groups per position): a gradient that sits on a
    # bf16 rounding boundary may round the other way
    torch.testing.assert_close(dx.float(), dref.to(torch.bfloat16).float(), rtol=2.0 ** -7, atol=1e-9)
    # against torch autograd on the same bf16-valued map
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    pos = torch.linspace(-1, 1, h, device="cuda")
    sm = torch.softmax(xr.reshape(n, c, h * h), dim=-1).reshape(n, c, h, h)
    ex, ey = (sm.sum(3) * pos).sum(2), (sm.sum(2) * pos).sum(2)  # expected row / column coordinate (vision_network.py:100-108)
    feat = torch.stack([ex, ey], dim=2).reshape(n, 2 * c)
    torch.testing.assert_close(out, feat.detach(), rtol=1e-5, atol=1e-6)
    feat.backward(dout)
    want = (xr.grad * (xr.detach() > 0)).permute(0, 2, 3, 1)
    torch.testing.assert_close(dx.float(), want, rtol=2.0 ** -7, atol=1e-7)
