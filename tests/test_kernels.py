"""Each kernel family of hulc_b200/csrc against torch (CPU autograd for the backward kernels).  Every test body runs on both
builds of the kernel sources through the `K` fixture: the host SIMT emulator (CPU suite) and the nvcc-built library on
the B200 (`-m gpu`)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import hulc_oracle as O

CONV_CASES = [(3, 32, 8, 4, 36, 2), (32, 64, 4, 2, 12, 3), (64, 64, 3, 1, 9, 2), (3, 32, 8, 4, 44, 1), (32, 64, 4, 2, 9, 2)]


@pytest.mark.parametrize("cin,cout,ks,st,hw,n", CONV_CASES)
def test_conv_fwd_bwd(K, cin, cout, ks, st, hw, n):
    g = torch.Generator().manual_seed(cin + hw)
    x = torch.randn(n, cin, hw, hw, generator=g).relu_()  # like a ReLU output: exercises the dgrad gate
    w = (torch.randn(cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks)).requires_grad_(True)
    b = torch.randn(cout, generator=g).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.relu(F.conv2d(xr, w, b, stride=st))
    y = K.conv2d_fwd(x, w.detach(), b.detach(), st)
    torch.testing.assert_close(y, ref.detach(), rtol=1e-4, atol=1e-5)
    dy = torch.randn(ref.shape, generator=g) * (ref.detach() > 0)
    ref.backward(dy)
    dw = torch.zeros_like(w)
    K.conv2d_wgrad(x, dy, dw, st)
    torch.testing.assert_close(dw, w.grad, rtol=1e-4, atol=1e-4)
    K.conv2d_wgrad(x, dy, dw, st, beta=1.0)
    torch.testing.assert_close(dw, 2 * w.grad, rtol=1e-4, atol=2e-4)
    db = torch.zeros(cout)
    K.nchw_channel_sum(dy, db)
    torch.testing.assert_close(db, b.grad, rtol=1e-4, atol=1e-4)
    if cin != 3:
        dx = K.conv2d_dgrad(dy, w.detach(), x.shape, st, gate=x)
        torch.testing.assert_close(dx, xr.grad * (x > 0), rtol=1e-4, atol=1e-4)
        dx2 = K.conv2d_dgrad(dy, w.detach(), x.shape, st)
        torch.testing.assert_close(dx2, xr.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("h", [21, 5])
def test_spatial_softmax(K, h):
    x = torch.randn(3, 4, h, h).relu_().requires_grad_(True)
    ref = O.spatial_softmax(x)
    out = K.spatial_softmax_fwd(x.detach())
    torch.testing.assert_close(out, ref.detach(), rtol=1e-5, atol=1e-6)
    dout = torch.randn_like(ref)
    ref.backward(dout)
    dx = K.spatial_softmax_bwd(x.detach(), dout, relu_gate=True)
    torch.testing.assert_close(dx, x.grad * (x.detach() > 0), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("h,c", [(21, 64), (5, 8), (7, 100), (22, 8)])
def test_spatial_softmax_channels_last(K, h, c):
    x = torch.randn(3, c, h, h).relu_().requires_grad_(True)
    ref = O.spatial_softmax(x)
    xh = x.detach().permute(0, 2, 3, 1).contiguous()
    out = K.spatial_softmax_nhwc_fwd(xh)
    torch.testing.assert_close(out, ref.detach(), rtol=1e-5, atol=1e-6)
    dout = torch.randn_like(ref)
    ref.backward(dout)
    dx = K.spatial_softmax_nhwc_bwd(xh, dout, relu_gate=True)
    torch.testing.assert_close(dx.permute(0, 3, 1, 2), x.grad * (x.detach() > 0), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("D,rows", [(64, 37), (32, 5), (128, 70)])
def test_layernorm_residual_dropout(K, D, rows):
    g = torch.Generator().manual_seed(D)
    x = torch.randn(rows, D, generator=g, requires_grad=True)
    res = torch.randn(rows, D, generator=g, requires_grad=True)
    w = torch.randn(D, generator=g, requires_grad=True)
    b = torch.randn(D, generator=g, requires_grad=True)
    keep = (torch.rand(rows, D, generator=g) > 0.2)
    ref = F.layer_norm(res + x * keep / 0.8, (D,), w, b)
    y, z, st = torch.empty(rows, D), torch.empty(rows, D), torch.empty(rows, 2)
    K.layernorm_fwd(x.detach(), w.detach(), b.detach(), y, st, res=res.detach(), z=z, drop=K.Drop(0.2, keep=keep.to(torch.uint8)))
    torch.testing.assert_close(y, ref.detach(), rtol=1e-4, atol=1e-5)
    dy = torch.randn(rows, D, generator=g)
    ref.backward(dy)
    dz, dx, dw, db = torch.empty(rows, D), torch.empty(rows, D), torch.zeros(D), torch.zeros(D)
    K.layernorm_bwd(dy, z, st, w.detach(), dw, db, dz=dz, dx=dx, drop=K.Drop(0.2, keep=keep.to(torch.uint8)))
    torch.testing.assert_close(dz, res.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dx, x.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dw, w.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(db, b.grad, rtol=1e-4, atol=1e-4)
    # plain LN into a strided output view
    big = torch.zeros(rows, 2 * D)
    K.layernorm_fwd(x.detach(), w.detach(), b.detach(), big[:, D:], st)
    torch.testing.assert_close(big[:, D:], F.layer_norm(x.detach(), (D,), w.detach(), b.detach()), rtol=1e-4, atol=1e-5)
    assert float(big[:, :D].abs().max()) == 0


@pytest.mark.parametrize("B,S,H,dh,p", [(3, 8, 8, 16, 0.0), (2, 32, 8, 16, 0.1), (2, 5, 2, 4, 0.3), (2, 64, 8, 16, 0.1), (2, 4, 2, 8, 0.0)])
def test_attention(K, B, S, H, dh, p):
    g = torch.Generator().manual_seed(S)
    D = H * dh
    qkv = torch.randn(B * S, 3 * D, generator=g, requires_grad=True)
    keep = torch.rand(B, H, S, S, generator=g) >= p
    q, k, v = [t.view(B, S, H, dh).transpose(1, 2) for t in qkv.split(D, -1)]
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), -1)
    ref = ((a * keep / (1 - p)) @ v).transpose(1, 2).reshape(B * S, D)
    drop = K.Drop(p, keep=keep.to(torch.uint8).contiguous()) if p > 0 else K.NO_DROP
    out, probs = torch.empty(B * S, D), torch.empty(B, H, S, S)
    K.attention_fwd(qkv.detach(), out, probs, B, S, H, drop)
    torch.testing.assert_close(out, ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(probs, a.detach(), rtol=1e-4, atol=1e-6)
    dout = torch.randn(B * S, D, generator=g)
    ref.backward(dout)
    dqkv = torch.empty(B * S, 3 * D)
    K.attention_bwd(qkv.detach(), probs, dout, dqkv, B, S, H, drop)
    torch.testing.assert_close(dqkv, qkv.grad, rtol=1e-4, atol=1e-5)


def test_posemb_strided_reduce(K):
    x, pos = torch.randn(3, 5, 8), torch.randn(7, 8)
    y = K.add_posemb_fwd(x, pos, torch.empty(3, 5, 8))
    torch.testing.assert_close(y, x + pos[:5])
    dst = torch.zeros(5, 3, 4)
    K.strided_copy(dst, x[:, :, 4:].transpose(0, 1))
    torch.testing.assert_close(dst, x[:, :, 4:].transpose(0, 1).contiguous())
    K.strided_copy(dst, x[:, :, 4:].transpose(0, 1), alpha=2.0, accumulate=True)
    torch.testing.assert_close(dst, 3 * x[:, :, 4:].transpose(0, 1))
    e = torch.empty(3, 5, 8)
    K.strided_copy(e, pos[0].view(1, 1, 8).expand(3, 5, 8), alpha=0.5)
    torch.testing.assert_close(e, 0.5 * pos[0].expand(3, 5, 8))
    torch.testing.assert_close(K.reduce_mid(x, torch.empty(3, 8), 0.2), x.mean(1))
    torch.testing.assert_close(K.sum_to(x, torch.empty(1), 0.5), 0.5 * x.sum().view(1))
    z = x.clone()
    torch.testing.assert_close(K.scale_(z, 3.0), 3 * x)


def test_world_to_tcp(K):
    from hulc_b200.utils import synthetic
    d = synthetic.make_modality("vis", 3, 7)
    ref = O.world_to_tcp_frame(d["actions"], d["state_info"]["robot_obs"])
    flag = torch.zeros(1, dtype=torch.int32)
    out = K.world_to_tcp(d["actions"], d["state_info"]["robot_obs"], torch.empty(3, 7, 7), flag)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-4)
    assert int(flag) == 0
    bad = d["actions"].clone()
    bad[0, 0, 0] = float("nan")
    K.world_to_tcp(bad, d["state_info"]["robot_obs"], torch.empty(3, 7, 7), flag)
    assert int(flag) == 1


def test_tcp_to_world(K):
    from hulc_b200.utils import synthetic
    d = synthetic.make_modality("vis", 3, 7)
    obs = d["state_info"]["robot_obs"]
    ref = O.tcp_to_world_frame(d["actions"], obs)
    flag = torch.zeros(1, dtype=torch.int32)
    out = K.tcp_to_world(d["actions"], obs, torch.empty(3, 7, 7), flag)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-4)
    assert int(flag) == 0
    # the two frame changes are inverses of each other (gripper_control.py:16-63)
    back = K.world_to_tcp(out.contiguous(), obs, torch.empty(3, 7, 7), flag)
    torch.testing.assert_close(back, d["actions"], rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("time_major,has_grip", [(True, True), (False, True), (True, False)])
def test_logistic_sample(K, time_major, has_grip):
    g = torch.Generator().manual_seed(11)
    B, S, b0, Bm, n_dims, n_mix = 5, 6, 1, 3, 6, 10
    nm = n_dims * n_mix
    n = 3 * nm + (2 if has_grip else 0)
    heads_b = torch.randn(B, S, n, generator=g)
    heads_b[..., 2 * nm : 3 * nm] = heads_b[..., 2 * nm : 3 * nm] * 3 - 4  # some log-scales below the -7 clamp
    sub = heads_b[b0 : b0 + Bm]
    lp, mu, ls = [sub[..., i * nm : (i + 1) * nm].reshape(Bm, S, n_dims, n_mix) for i in range(3)]
    grip = sub[..., 3 * nm :] if has_grip else None
    u_mix, u_inv = torch.rand(Bm, S, n_dims, n_mix, generator=g), torch.rand(Bm, S, n_dims, generator=g)
    ref = O.logistic_mixture_sample(lp, ls.clamp(min=-7.0), mu, grip, u_mix, u_inv)
    ld = n + 2  # padded rows, like the engine's head buffer
    hbuf = torch.zeros(B * S, ld)
    hbuf[:, :n] = heads_b.transpose(0, 1).reshape(S * B, n) if time_major else heads_b.reshape(B * S, n)
    A = n_dims + (1 if has_grip else 0)
    out = torch.full((B, S, A), 7.0)
    K.logistic_sample(hbuf[:, :n], out, B, S, b0, Bm, time_major=time_major, n_dims=n_dims, n_mix=n_mix, has_gripper=has_grip, u_mix=u_mix, u_inv=u_inv)
    torch.testing.assert_close(out[b0 : b0 + Bm], ref, rtol=1e-5, atol=1e-5)
    assert float((out[:b0] - 7.0).abs().max()) == 0.0 and float((out[b0 + Bm :] - 7.0).abs().max()) == 0.0  # other sequences untouched
    # Philox path: reproducible for a seed, different across seeds, finite
    o1, o2, o3 = (torch.zeros(B, S, A) for _ in range(3))
    for o, sd in ((o1, 5), (o2, 5), (o3, 6)):
        K.logistic_sample(hbuf[:, :n], o, B, S, 0, B, time_major=time_major, n_dims=n_dims, n_mix=n_mix, has_gripper=has_grip, seed=sd, site=40)
    assert torch.equal(o1, o2) and not torch.equal(o1, o3) and bool(torch.isfinite(o1).all())


def test_val_metrics(K):
    g = torch.Generator().manual_seed(2)
    B, S = 4, 9
    pred, act = torch.randn(B, S, 7, generator=g), torch.rand(B, S, 7, generator=g) * 2 - 1
    act[..., 6] = (torch.rand(B, S, generator=g) < 0.5).float() * 2 - 1
    mae, hits = torch.empty(B, 6), torch.empty(B)
    K.val_metrics(pred, act, mae, hits)  # results land in the caller's tensors (the cuda variant copies storages back)
    ref_mae, ref_sr = O.validation_metrics(pred, act)
    torch.testing.assert_close(mae, ref_mae, rtol=1e-5, atol=1e-6)
    assert abs(float(hits.sum()) / (B * S) - float(ref_sr)) < 1e-6


@pytest.mark.parametrize("time_major,n_dims,num_classes,has_grip", [(True, 6, 10, True), (False, 6, 10, True), (True, 7, 256, False)])
def test_logistic_loss(K, time_major, n_dims, num_classes, has_grip):
    g = torch.Generator().manual_seed(3)
    B, S, b0, Bm, n_mix = 5, 6, 1, 3, 10
    nm = n_dims * n_mix
    n = 3 * nm + (2 if has_grip else 0)
    heads_b = torch.randn(B, S, n, generator=g)
    heads_b[..., 2 * nm : 3 * nm] = heads_b[..., 2 * nm : 3 * nm] * 3 - 4  # some log-scales below the -7 clamp
    acts = torch.rand(B, S, 7, generator=g) * 2 - 1
    acts[..., :3] *= 1.2  # some beyond the bounds -> the edge branches
    acts[0:3, 0, 0] = torch.tensor([-1.0, 1.0, 0.9995])
    acts[..., 6] = (torch.rand(B, S, generator=g) < 0.5).float() * 2 - 1
    hb = heads_b.clone().requires_grad_(True)
    sub = hb[b0 : b0 + Bm]
    lp, mu, ls = [sub[..., i * nm : (i + 1) * nm].reshape(Bm, S, n_dims, n_mix) for i in range(3)]
    grip = sub[..., 3 * nm :] if has_grip else None
    a = acts[b0 : b0 + Bm]
    nll = O.logistic_mixture_nll(lp, ls, mu, a[..., :n_dims], -1.0, 1.0, num_classes, -7.0)
    total = nll
    if has_grip:
        ce = F.cross_entropy(grip.reshape(-1, 2), (a[..., -1] != -1).long().view(-1))
        total = nll + 1.0 * ce
    (0.5 * total).backward()
    heads = heads_b.transpose(0, 1).contiguous().view(S * B, n) if time_major else heads_b.reshape(B * S, n)
    dheads = torch.zeros_like(heads)
    losses = torch.zeros(2)
    K.logistic_loss(heads, acts, dheads, losses, B, S, b0, Bm, time_major=time_major, n_dims=n_dims, n_mix=n_mix, num_classes=num_classes,
                      has_gripper=has_grip, grad_scale=0.5)
    torch.testing.assert_close(losses[0], nll.detach(), rtol=1e-5, atol=1e-5)
    if has_grip:
        torch.testing.assert_close(losses[1], ce.detach(), rtol=1e-5, atol=1e-6)
    d = dheads.view(S, B, n).transpose(0, 1) if time_major else dheads.view(B, S, n)
    torch.testing.assert_close(d, hb.grad, rtol=1e-3, atol=1e-6)


def test_plan_discrete(K):
    g = torch.Generator().manual_seed(9)
    Bn = 5
    pr = torch.randn(Bn, 32, 32, generator=g, requires_grad=True)
    pp = torch.randn(Bn, 32, 32, generator=g, requires_grad=True)
    u = torch.rand(Bn, 32, generator=g)
    idx_ref = O.categorical_sample_indices(pr.detach(), u)
    plan_ref = O.discrete_rsample(pr, idx_ref).flatten(-2)
    kl = O.kl_loss_discrete(pp.view(Bn, -1), pr.view(Bn, -1), 32, 32, 0.01, 0.8)
    dplan = torch.randn(Bn, 1024, generator=g)
    ((plan_ref * dplan).sum() + kl).backward()
    plan, kl_rows, idx = torch.empty(Bn, 1024), torch.empty(Bn * 32), torch.empty(Bn * 32, dtype=torch.int32)
    K.plan_discrete_fwd(pr.detach(), pp.detach(), plan, kl_rows, u=u.view(-1), idx_out=idx)
    assert torch.equal(idx.long().view(Bn, 32), idx_ref)
    torch.testing.assert_close(plan, plan_ref.detach())
    torch.testing.assert_close(0.01 * kl_rows.sum() / Bn, kl.detach(), rtol=1e-5, atol=1e-7)
    d_pr, d_pp = torch.empty(Bn, 1024), torch.empty(Bn, 1024)
    K.plan_discrete_bwd(pr.detach(), pp.detach(), dplan, d_pr, d_pp, 0.01 * 0.8 / Bn, 0.01 * 0.2 / Bn)
    torch.testing.assert_close(d_pr.view(Bn, 32, 32), pr.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(d_pp.view(Bn, 32, 32), pp.grad, rtol=1e-4, atol=1e-8)
    # injected indices; philox sampling is deterministic and in range
    plan2 = torch.empty(Bn, 1024)
    K.plan_discrete_fwd(pr.detach(), pp.detach(), plan2, kl_rows, idx_in=idx)
    assert torch.equal(plan2, plan)
    i1, i2 = torch.empty(Bn * 32, dtype=torch.int32), torch.empty(Bn * 32, dtype=torch.int32)
    K.plan_discrete_fwd(pr.detach(), pp.detach(), plan2, kl_rows, idx_out=i1, seed=5, site=1)
    K.plan_discrete_fwd(pr.detach(), pp.detach(), plan2, kl_rows, idx_out=i2, seed=5, site=1)
    assert torch.equal(i1, i2) and int(i1.min()) >= 0 and int(i1.max()) < 32 and plan2.sum() == Bn * 32


def test_plan_continuous(K):
    g = torch.Generator().manual_seed(4)
    Bn, Pn = 4, 256
    pr = torch.randn(Bn, 2 * Pn, generator=g, requires_grad=True)
    pp = torch.randn(Bn, 2 * Pn, generator=g, requires_grad=True)
    eps = torch.randn(Bn, Pn, generator=g)
    mean, std = O.cont_state(pr)
    plan_ref = mean + std * eps
    kl = O.kl_loss_continuous(pp, pr, 0.01, 0.8)
    dplan = torch.randn(Bn, Pn, generator=g)
    ((plan_ref * dplan).sum() + kl).backward()
    plan, kl_el = torch.empty(Bn, Pn), torch.empty(Bn, Pn)
    K.plan_cont_fwd(pr.detach(), pp.detach(), plan, kl_el, eps=eps)
    torch.testing.assert_close(plan, plan_ref.detach())
    torch.testing.assert_close(0.01 * kl_el.sum() / Bn, kl.detach(), rtol=1e-5, atol=1e-7)
    d_pr, d_pp = torch.empty(Bn, 2 * Pn), torch.empty(Bn, 2 * Pn)
    K.plan_cont_bwd(pr.detach(), pp.detach(), dplan, d_pr, d_pp, 0.01 * 0.8 / Bn, 0.01 * 0.2 / Bn, eps=eps)
    torch.testing.assert_close(d_pr, pr.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(d_pp, pp.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("masked", [None, "some", "all"])
def test_clip_loss(K, masked):
    g = torch.Generator().manual_seed(2)
    n, D = 6, 32
    im = torch.randn(n, D, generator=g, requires_grad=True)
    tx = torch.randn(n, D, generator=g, requires_grad=True)
    ls = torch.tensor(math.log(1 / 0.07), requires_grad=True)
    mask = None if masked is None else (torch.tensor([1, 0, 1, 1, 0, 1], dtype=torch.bool) if masked == "some" else torch.zeros(n, dtype=torch.bool))
    sel = slice(None) if mask is None else mask
    loss, d_im, d_tx, d_ls = torch.empty(1), torch.empty(n, D), torch.empty(n, D), torch.empty(1)
    K.clip_loss(im.detach(), tx.detach(), ls.detach().view(1), None if mask is None else mask.to(torch.uint8), loss, d_im, d_tx, d_ls, grad_scale=3.0)
    if masked == "all":
        assert float(loss) == 0 and float(d_im.abs().max()) == 0 and float(d_ls) == 0
        return
    a, t = im[sel], tx[sel]
    a = a / a.norm(dim=-1, keepdim=True)
    t = t / t.norm(dim=-1, keepdim=True)
    logits = ls.exp() * a @ t.t()
    lab = torch.arange(logits.shape[0])
    ref = (F.cross_entropy(logits, lab) + F.cross_entropy(logits.t(), lab)) / 2
    (3.0 * ref).backward()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d_im, im.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(d_tx, tx.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(d_ls[0], ls.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("B,D,ld", [(5, 384, 384), (1, 20, 24), (64, 33, 40)])
def test_cosine_loss(K, B, D, ld):
    """hulc_cosine_loss: the BC-Z language-regression loss (hulc.py:596-601) — mean cosine distance of the rows, and its gradient."""
    g = torch.Generator().manual_seed(B + D)
    pred = torch.randn(B, ld, generator=g)[:, :D].requires_grad_(True)  # rows with a leading dimension
    tgt = torch.randn(B, ld, generator=g)[:, :D]
    loss, dpred = torch.full((1,), 7.0), torch.zeros(B, ld)[:, :D]
    K.cosine_loss(pred.detach(), tgt, dpred, loss, grad_scale=1.5)
    cos = (pred * tgt).sum(-1) / (torch.linalg.norm(pred, dim=1) * torch.linalg.norm(tgt, dim=1))
    ref = (1 - cos).mean()
    (1.5 * ref).backward()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dpred, pred.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("n_pos,n_neg", [(4, 4), (1, 1), (300, 300)])
def test_bce_logits_loss(K, n_pos, n_neg):
    """hulc_bce_logits_loss: binary_cross_entropy_with_logits over matching (label 1) and rolled (label 0) pairs (hulc.py:637-645)."""
    g = torch.Generator().manual_seed(n_pos)
    x = (torch.randn(n_pos + n_neg, generator=g) * 4).requires_grad_(True)  # large logits exercise the log1p(exp(-|x|)) form
    loss, dx = torch.full((1,), -3.0), torch.zeros(n_pos + n_neg)
    K.bce_logits_loss(x.detach(), dx, loss, n_pos, n_neg, grad_scale=0.5)
    labels = torch.cat([torch.ones(n_pos), torch.zeros(n_neg)])
    ref = F.binary_cross_entropy_with_logits(x, labels)
    (0.5 * ref).backward()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dx, x.grad, rtol=1e-4, atol=1e-7)


def test_gru_gates(K):
    g = torch.Generator().manual_seed(8)
    B, H = 3, 40
    gi = torch.randn(B, 3 * H, generator=g, requires_grad=True)
    gh = torch.randn(B, 3 * H, generator=g, requires_grad=True)
    hp = torch.randn(B, H, generator=g, requires_grad=True)
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H : 2 * H] + gh[:, H : 2 * H])
    n = torch.tanh(gi[:, 2 * H :] + r * gh[:, 2 * H :])
    href = (1 - z) * n + z * hp
    d1, d2 = torch.randn(B, H, generator=g), torch.randn(B, H, generator=g)
    href.backward(d1 + d2)
    h, saved = torch.empty(B, H), torch.empty(B, 4 * H)
    K.gru_gates_fwd(gi.detach(), gh.detach(), hp.detach(), h, saved)
    torch.testing.assert_close(h, href.detach())
    dgi, dgh, carry = torch.empty(B, 3 * H), torch.empty(B, 3 * H), torch.empty(B, H)
    K.gru_gates_bwd(d1, d2, saved, hp.detach(), dgi, dgh, carry)
    torch.testing.assert_close(dgi, gi.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(dgh, gh.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(carry, hp.grad, rtol=1e-4, atol=1e-6)


def test_gemm_tanh_modes(K):
    A, B = torch.randn(9, 20), torch.randn(20, 11)
    torch.testing.assert_close(K.gemm(A, B, act=2), torch.tanh(A @ B), rtol=1e-4, atol=1e-5)
    gate = torch.tanh(torch.randn(9, 11))
    torch.testing.assert_close(K.gemm(A, B, act=4, gate=gate), (A @ B) * (1 - gate * gate), rtol=1e-4, atol=1e-5)


def test_adam_matches_torch(K):
    p = torch.randn(1000)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-4)
    m, v = torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 4):
        g = torch.randn(1000)
        ref.grad = g.clone()
        opt.step()
        K.adam_step(p, g * 4, m, v, lr=2e-4, step=step, grad_scale=0.25)
    torch.testing.assert_close(p, ref.detach(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("n", [16 * 7, 1003])
def test_frames_u8_to_f32(K, n):
    """uint8 frames -> normalised fp32: bit-identical to the reference's CPU transforms (x / 255 - 0.5) / 0.5."""
    g = torch.Generator().manual_seed(n)
    src = torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8)
    dst = torch.empty(n)
    K.frames_u8_to_f32(src, dst)
    ref = (src.float() / 255 - 0.5) / 0.5
    assert torch.equal(dst, ref)
