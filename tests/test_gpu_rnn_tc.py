"""hulc_rnn_tc_seq (the whole Elman recurrence in one persistent launch, W_hh resident in shared memory) against a float64
torch loop on the B200: forward ReLU / tanh, both directions, BPTT with the ReLU and tanh gates, full and ragged batch, and
bit-reproducibility.  Tolerance: tf32 operands (10-bit mantissa, round-to-nearest) with fp32 accumulation."""
import pytest
import torch

pytestmark = pytest.mark.gpu

H = 2048


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")


def _mk(B, S, ld, seed):
    g = torch.Generator().manual_seed(seed)
    W = (torch.rand(H, H, generator=g) * 2 - 1) / H ** 0.5          # torch.nn.RNN's default init
    pre = torch.randn(S, B, H, generator=g) * 0.5
    hbuf = torch.zeros(S + 2, B, ld)
    return W.cuda(), pre.cuda(), hbuf.cuda()


@pytest.mark.parametrize("B,S", [(64, 8), (5, 4), (12, 32)])
@pytest.mark.parametrize("kind", ["relu", "tanh"])
@pytest.mark.parametrize("reverse", [False, True])
def test_rnn_seq_forward(B, S, kind, reverse):
    from hulc_b200 import ops

    ld, col0 = 2 * H, H  # a column slice of a wider buffer, like the bidirectional posterior
    W, pre, hbuf = _mk(B, S, ld, 3 * B + S)
    h = lambda slot: hbuf[slot, :, col0 : col0 + H]
    st, sp = hbuf.stride(0), pre.stride(0)
    act = 2 if kind == "tanh" else 1
    if reverse:
        ops.rnn_tc_seq(W, h(S + 1), h(S), pre[S - 1], S, prev_step=-st, out_step=-st, add_step=-sp, act=act)
    else:
        ops.rnn_tc_seq(W, h(0), h(1), pre[0], S, prev_step=st, out_step=st, add_step=sp, act=act)
    torch.cuda.synchronize()
    f = torch.relu if kind == "relu" else torch.tanh
    Wd, ref = W.double(), torch.zeros(S + 2, B, H, dtype=torch.float64, device="cuda")
    for t in (range(S - 1, -1, -1) if reverse else range(S)):
        prev = ref[t + 2] if reverse else ref[t]
        ref[t + 1] = f(pre[t].double() + prev @ Wd.t())
    got = hbuf[:, :, col0 : col0 + H].double()
    assert float(hbuf[:, :, :col0].abs().max()) == 0.0  # the neighbouring columns are untouched
    err = float((got - ref).abs().max())
    assert err < 3e-3 * max(1.0, float(ref.abs().max())), err
    assert float(got[0].abs().max()) == 0.0 and float(got[S + 1].abs().max()) == 0.0


@pytest.mark.parametrize("B,S", [(64, 8), (6, 5)])
@pytest.mark.parametrize("kind", ["relu", "tanh"])
@pytest.mark.parametrize("reverse", [False, True])
def test_rnn_seq_backward(B, S, kind, reverse):
    from hulc_b200 import ops

    W, dh_above, hbuf = _mk(B, S, H, 11 * B + S)
    g = torch.Generator().manual_seed(5)
    hvals = torch.randn(S, B, H, generator=g).cuda()
    hbuf[1 : S + 1] = torch.relu(hvals) if kind == "relu" else torch.tanh(hvals)
    dbuf = torch.zeros(S + 1, B, H, device="cuda")
    act = 4 if kind == "tanh" else 0
    sd, sh, sa = dbuf.stride(0), hbuf.stride(0), dh_above.stride(0)
    if reverse:
        ops.rnn_tc_seq(W, dbuf[0], dbuf[1], dh_above[0], S, prev_step=sd, out_step=sd, add_step=sa, gate0=hbuf[1], gate_step=sh, act=act, transW=True)
    else:
        ops.rnn_tc_seq(W, dbuf[S], dbuf[S - 1], dh_above[S - 1], S, prev_step=-sd, out_step=-sd, add_step=-sa, gate0=hbuf[S], gate_step=-sh, act=act,
                       transW=True)
    torch.cuda.synchronize()
    Wd, ref = W.double(), torch.zeros(S + 1, B, H, dtype=torch.float64, device="cuda")
    for t in (range(S) if reverse else range(S - 1, -1, -1)):
        nxt = ref[t] if reverse else ref[t + 1]
        hh = hbuf[t + 1].double()
        gate = (1 - hh * hh) if kind == "tanh" else (hh > 0).double()
        v = (dh_above[t].double() + nxt @ Wd) * gate
        if reverse:
            ref[t + 1] = v
        else:
            ref[t] = v
    err = float((dbuf.double() - ref).abs().max())
    assert err < 3e-3 * max(1.0, float(ref.abs().max())), err


def test_rnn_seq_reproducible():
    from hulc_b200 import ops

    W, pre, hbuf = _mk(64, 16, H, 99)
    outs = []
    for _ in range(2):
        hbuf.zero_()
        ops.rnn_tc_seq(W, hbuf[0], hbuf[1], pre[0], 16, prev_step=hbuf.stride(0), out_step=hbuf.stride(0), add_step=pre.stride(0), act=1)
        outs.append(hbuf.clone())
    assert torch.equal(outs[0], outs[1])


def test_rnn_seq_rejects_unsupported_shapes():
    from hulc_b200 import _lib, ops

    W = torch.zeros(1024, 1024, device="cuda")
    hb = torch.zeros(4, 8, 1024, device="cuda")
    with pytest.raises(_lib.HulcError):
        ops.rnn_tc_seq(W, hb[0], hb[1], hb[2], 1, prev_step=0, out_step=0, add_step=0, act=1)


# ---- bf16 variant (hulc_rnn_seq_bf16): bf16 W_hh and bf16 hidden state between the steps, fp32 accumulate / addend / activation ----
def _bf(x):
    return x.to(torch.bfloat16).double()


@pytest.mark.parametrize("B,S", [(64, 8), (5, 4), (12, 32)])
@pytest.mark.parametrize("kind", ["relu", "tanh"])
@pytest.mark.parametrize("reverse", [False, True])
def test_rnn_seq_bf16_forward(B, S, kind, reverse):
    from hulc_b200 import ops

    ld, col0 = 2 * H, H
    W, pre, hbuf = _mk(B, S, ld, 3 * B + S)
    W16 = W.to(torch.bfloat16)
    x16 = torch.empty((S + 1) * B * H, dtype=torch.bfloat16, device="cuda")
    h = lambda slot: hbuf[slot, :, col0 : col0 + H]
    st, sp = hbuf.stride(0), pre.stride(0)
    act = 2 if kind == "tanh" else 1
    if reverse:
        ops.rnn_seq_bf16(W16, h(S + 1), x16, h(S), pre[S - 1], S, out_step=-st, add_step=-sp, act=act)
    else:
        ops.rnn_seq_bf16(W16, h(0), x16, h(1), pre[0], S, out_step=st, add_step=sp, act=act)
    torch.cuda.synchronize()
    f = torch.relu if kind == "relu" else torch.tanh
    # reference with the same roundings: bf16 weights, bf16 state into the product, everything else in float64
    Wd, ref = W16.double(), torch.zeros(S + 2, B, H, dtype=torch.float64, device="cuda")
    for t in (range(S - 1, -1, -1) if reverse else range(S)):
        prev = ref[t + 2] if reverse else ref[t]
        ref[t + 1] = f(pre[t].double() + _bf(prev) @ Wd.t())
    got = hbuf[:, :, col0 : col0 + H].double()
    assert float(hbuf[:, :, :col0].abs().max()) == 0.0
    err = float((got - ref).abs().max())
    # a value that sits on a bf16 rounding boundary may round the other way than in the float64 chain (one bf16 ulp of h, spread by W)
    assert err < 2e-2 * max(1.0, float(ref.abs().max())), err
    assert float((got - ref).abs().mean()) < 1e-3
    xs = x16.view(S + 1, B, H)[1:].double()
    out_steps = got[1 : S + 1].flip(0) if reverse else got[1 : S + 1]
    assert torch.equal(xs, _bf(out_steps))  # the exchanged state is the bf16 rounding of the fp32 result


@pytest.mark.parametrize("B,S", [(64, 8), (6, 5)])
@pytest.mark.parametrize("kind", ["relu", "tanh"])
@pytest.mark.parametrize("reverse", [False, True])
def test_rnn_seq_bf16_backward(B, S, kind, reverse):
    from hulc_b200 import ops

    W, dh_above, hbuf = _mk(B, S, H, 11 * B + S)
    W16 = W.to(torch.bfloat16)
    x16 = torch.empty((S + 1) * B * H, dtype=torch.bfloat16, device="cuda")
    g = torch.Generator().manual_seed(5)
    hvals = torch.randn(S, B, H, generator=g).cuda()
    hbuf[1 : S + 1] = torch.relu(hvals) if kind == "relu" else torch.tanh(hvals)
    dbuf = torch.zeros(S + 1, B, H, device="cuda")
    act = 4 if kind == "tanh" else 0
    sd, sh, sa = dbuf.stride(0), hbuf.stride(0), dh_above.stride(0)
    if reverse:
        ops.rnn_seq_bf16(W16, dbuf[0], x16, dbuf[1], dh_above[0], S, out_step=sd, add_step=sa, gate0=hbuf[1], gate_step=sh, act=act, transW=True)
    else:
        ops.rnn_seq_bf16(W16, dbuf[S], x16, dbuf[S - 1], dh_above[S - 1], S, out_step=-sd, add_step=-sa, gate0=hbuf[S], gate_step=-sh, act=act,
                         transW=True)
    torch.cuda.synchronize()
    Wd, ref = W16.double(), torch.zeros(S + 1, B, H, dtype=torch.float64, device="cuda")
    for t in (range(S) if reverse else range(S - 1, -1, -1)):
        nxt = ref[t] if reverse else ref[t + 1]
        hh = hbuf[t + 1].double()
        gate = (1 - hh * hh) if kind == "tanh" else (hh > 0).double()
        v = (dh_above[t].double() + _bf(nxt) @ Wd) * gate
        if reverse:
            ref[t + 1] = v
        else:
            ref[t] = v
    err = float((dbuf.double() - ref).abs().max())
    assert err < 2e-2 * max(1.0, float(ref.abs().max())), err
    assert float((dbuf.double() - ref).abs().mean()) < 1e-3


def test_rnn_seq_bf16_reproducible():
    from hulc_b200 import ops

    W, pre, hbuf = _mk(64, 32, H, 99)
    W16 = W.to(torch.bfloat16)
    x16 = torch.empty(33 * 64 * H, dtype=torch.bfloat16, device="cuda")
    outs = []
    for _ in range(3):
        hbuf.zero_()
        ops.rnn_seq_bf16(W16, hbuf[0], x16, hbuf[1], pre[0], 32, out_step=hbuf.stride(0), add_step=pre.stride(0), act=1)
        outs.append(hbuf.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
