"""CPU oracle for the HULC training hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional fp32 PyTorch restatement of the reference algorithm (`lukashermann/hulc`, commit 7fdb09f) written from
the reference's formulas; every function cites the reference file:line it follows.  Parameters are read from a plain
`state_dict`-style mapping that uses the reference's key names, gradients come from CPU autograd.

Parity status: **pinned** — `oracle/make_golden.py` runs the UNMODIFIED reference modules (imported from
`/root/reference` under the dependency shims of `oracle/ref_stubs.py`) on the seeded inputs/weights of
`hulc_b200/utils/synthetic.py`, checks this file against them and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` re-checks the oracle against those fixtures wherever the tests run.
(The reference itself ships no tests or golden vectors: SURVEY.md §4.)

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import this file.
The product package `hulc_b200` never does.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Mapping[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------------------------------
# small building blocks
# ----------------------------------------------------------------------------------------------------------------------
def linear(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """y = x W^T + b with W stored (out, in) — torch.nn.Linear as used throughout the reference."""
    return x @ sd[name + ".weight"].t() + sd[name + ".bias"]


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.LayerNorm over the last dim, biased variance, eps inside the sqrt (vision_network.py:53, goal_encoders.py:29)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def spatial_softmax(x: torch.Tensor, temperature: float = 1.0) -> torch.Tensor:
    """vision_network.py:88-108.  x (N,C,H,W) -> (N, 2C) interleaved [E_x(c), E_y(c)].

    `x_map`/`y_map` come from meshgrid(linspace(-1,1,num_cols), linspace(-1,1,num_rows), indexing="ij") flattened, so
    `x_map[p]` varies with the ROW index p // W of the flattened H*W position and `y_map[p]` with the column p % W
    (for the square maps used here num_rows == num_cols).
    """
    n, c, h, w = x.shape
    lin_r = torch.linspace(-1.0, 1.0, h, dtype=x.dtype, device=x.device)
    lin_c = torch.linspace(-1.0, 1.0, w, dtype=x.dtype, device=x.device)
    x_map = lin_r.view(h, 1).expand(h, w).reshape(-1)
    y_map = lin_c.view(1, w).expand(h, w).reshape(-1)
    a = torch.softmax(x.reshape(n * c, h * w) / temperature, dim=1)
    ex = (a * x_map).sum(1, keepdim=True)
    ey = (a * y_map).sum(1, keepdim=True)
    return torch.cat([ex, ey], 1).view(n, 2 * c)


def conv_stack(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """conv(k8,s4)+ReLU -> conv(k4,s2)+ReLU -> conv(k3,s1)+ReLU (vision_network.py:36-47, vision_network_gripper.py:11-17)."""
    for idx, stride in ((0, 4), (2, 2), (4, 1)):
        x = F.relu(F.conv2d(x, sd[f"{prefix}.conv_model.{idx}.weight"], sd[f"{prefix}.conv_model.{idx}.bias"], stride=stride))
    return x


def static_encoder(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """VisionNetwork.forward for the static camera (vision_network.py:55-65): convs -> spatial softmax -> fc1+ReLU
    -> fc2 -> LayerNorm.  dropout_vis_fc = 0, l2_normalize_output = False, use_sinusoid = False (rgb_static/default.yaml)."""
    x = conv_stack(sd, prefix, x)
    x = spatial_softmax(x, float(sd[f"{prefix}.spatial_softmax.temperature"]) if f"{prefix}.spatial_softmax.temperature" in sd else 1.0)
    x = F.relu(linear(sd, f"{prefix}.fc1.0", x))
    x = linear(sd, f"{prefix}.fc2", x)
    return layer_norm(x, sd[f"{prefix}.ln.weight"], sd[f"{prefix}.ln.bias"])


def gripper_encoder(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """VisionNetwork.forward for the gripper camera (vision_network_gripper.py:10-21,49-56): nature-CNN convs ->
    flatten (C,H,W order) -> FC 3136->128 + ReLU -> fc1+ReLU -> fc2 -> LayerNorm."""
    x = conv_stack(sd, prefix, x)
    x = x.flatten(1)
    x = F.relu(linear(sd, f"{prefix}.conv_model.7", x))
    x = F.relu(linear(sd, f"{prefix}.fc1.0", x))
    x = linear(sd, f"{prefix}.fc2", x)
    return layer_norm(x, sd[f"{prefix}.ln.weight"], sd[f"{prefix}.ln.bias"])


def perceptual_encoder(sd: SD, rgb_static: torch.Tensor, rgb_gripper: torch.Tensor, prefix: str = "perceptual_encoder") -> torch.Tensor:
    """ConcatEncoders.forward (concat_encoders.py:59-109): (B,S,C,H,W) -> (B*S,C,H,W) -> encoders -> cat -> (B,S,128)."""
    b, s = rgb_static.shape[:2]
    e_s = static_encoder(sd, f"{prefix}.rgb_static_encoder", rgb_static.reshape(b * s, *rgb_static.shape[2:]))
    e_g = gripper_encoder(sd, f"{prefix}.rgb_gripper_encoder", rgb_gripper.reshape(b * s, *rgb_gripper.shape[2:]))
    return torch.cat([e_s.view(b, s, -1), e_g.view(b, s, -1)], -1)


def goal_encoder(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """VisualGoalEncoder / LanguageGoalEncoder (goal_encoders.py:20-36 / 52-69): FC-ReLU-FC-ReLU-FC-LN.
    The language MLP starts with a Dropout(p=0) module, which shifts its Linear indices to 1,3,5."""
    idx = (0, 2, 4) if f"{prefix}.mlp.0.weight" in sd else (1, 3, 5)
    x = F.relu(linear(sd, f"{prefix}.mlp.{idx[0]}", x))
    x = F.relu(linear(sd, f"{prefix}.mlp.{idx[1]}", x))
    x = linear(sd, f"{prefix}.mlp.{idx[2]}", x)
    return layer_norm(x, sd[f"{prefix}.ln.weight"], sd[f"{prefix}.ln.bias"])


def plan_proposal(sd: SD, emb0: torch.Tensor, goal: torch.Tensor, prefix: str = "plan_proposal") -> torch.Tensor:
    """PlanProposalNetwork.forward (plan_proposal_net.py:42-47): cat -> 4x(FC 2048 + ReLU) -> FC -> state tensor
    (discrete: logits (B,1024); continuous: (B, 2*plan) = [mean | pre-softplus std])."""
    x = torch.cat([emb0, goal], -1)
    for i in (0, 2, 4, 6):
        x = F.relu(linear(sd, f"{prefix}.fc_model.{i}", x))
    return linear(sd, f"{prefix}.fc_state.0", x)


def _drop(x: torch.Tensor, p: float, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """Inverted dropout with an explicit keep-mask (1 = keep)."""
    if p == 0.0 or mask is None:
        return x
    return x * mask.to(x.dtype) / (1.0 - p)


def transformer_layer(sd: SD, prefix: str, x: torch.Tensor, nhead: int, p: float, masks: Optional[Dict[str, torch.Tensor]]) -> torch.Tensor:
    """nn.TransformerEncoderLayer(d, nhead, ff, dropout=p, activation=relu, norm_first=False) in training mode
    (constructed at plan_recognition_net.py:83-85).  x is batch-first (B,S,D) here; the reference runs seq-first, which
    only permutes rows.  masks (keep-masks, optional): attn (B,H,S,S), drop1 (B,S,D), ffn (B,S,FF), drop2 (B,S,D)."""
    B, S, D = x.shape
    dh = D // nhead
    m = masks or {}
    qkv = x @ sd[f"{prefix}.self_attn.in_proj_weight"].t() + sd[f"{prefix}.self_attn.in_proj_bias"]
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, S, nhead, dh).transpose(1, 2) * (1.0 / math.sqrt(dh))
    k = k.view(B, S, nhead, dh).transpose(1, 2)
    v = v.view(B, S, nhead, dh).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2), dim=-1)  # (B,H,S,S)
    a = _drop(a, p, m.get("attn"))
    o = (a @ v).transpose(1, 2).reshape(B, S, D)
    o = linear(sd, f"{prefix}.self_attn.out_proj", o)
    x = layer_norm(x + _drop(o, p, m.get("drop1")), sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"])
    h = _drop(F.relu(linear(sd, f"{prefix}.linear1", x)), p, m.get("ffn"))
    h = linear(sd, f"{prefix}.linear2", h)
    return layer_norm(x + _drop(h, p, m.get("drop2")), sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"])


def plan_recognition_transformer(
    sd: SD, emb: torch.Tensor, nhead: int = 8, nlayers: int = 2, p: float = 0.0,
    masks: Optional[Dict[str, torch.Tensor]] = None, prefix: str = "plan_recognition",
) -> Tuple[torch.Tensor, torch.Tensor]:
    """PlanRecognitionTransformersNetwork.forward (plan_recognition_net.py:94-117): + learned position embedding ->
    dropout -> nlayers post-norm encoder layers -> FC 128->4096 -> mean over time (= seq_feat) -> FC -> state tensor.
    masks keys: "in" (B,S,D) and f"l{i}.attn|drop1|ffn|drop2"."""
    B, S, D = emb.shape
    masks = masks or {}
    x = emb + sd[f"{prefix}.position_embeddings.weight"][:S].unsqueeze(0)
    x = _drop(x, p, masks.get("in"))
    for i in range(nlayers):
        lm = {k.split(".", 1)[1]: v for k, v in masks.items() if k.startswith(f"l{i}.")}
        x = transformer_layer(sd, f"{prefix}.transformer_encoder.layers.{i}", x, nhead, p, lm)
    x = linear(sd, f"{prefix}.fc", x)
    seq_feat = x.mean(1)
    return linear(sd, f"{prefix}.fc_state.0", seq_feat), seq_feat


def elman_rnn(sd: SD, prefix: str, x: torch.Tensor, num_layers: int, nonlinearity: str, reverse_suffix: str = "", h0=None) -> torch.Tensor:
    """torch.nn.RNN, batch_first, one direction: h_t = act(W_ih x_t + b_ih + W_hh h_{t-1} + b_hh), h_0 = 0
    (decoders/utils/rnn.py:5-14 with nonlinearity="relu"; plan_recognition_net.py:27-34 with tanh)."""
    act = torch.relu if nonlinearity == "relu" else torch.tanh
    B, S, _ = x.shape
    for l in range(num_layers):
        Wi, Wh = sd[f"{prefix}.weight_ih_l{l}{reverse_suffix}"], sd[f"{prefix}.weight_hh_l{l}{reverse_suffix}"]
        bi, bh = sd[f"{prefix}.bias_ih_l{l}{reverse_suffix}"], sd[f"{prefix}.bias_hh_l{l}{reverse_suffix}"]
        h = x.new_zeros(B, Wh.shape[0]) if h0 is None else h0[l]
        outs = []
        pre = x @ Wi.t() + bi
        for t in range(S):
            h = act(pre[:, t] + h @ Wh.t() + bh)
            outs.append(h)
        x = torch.stack(outs, 1)
    return x


def birnn_tanh(sd: SD, prefix: str, x: torch.Tensor, num_layers: int = 2) -> torch.Tensor:
    """torch.nn.RNN(bidirectional=True, tanh, batch_first) (plan_recognition_net.py:27-34): each layer runs a forward
    and a time-reversed pass over the previous layer's output and concatenates [fwd | bwd]."""
    B, S, _ = x.shape
    for l in range(num_layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            Wi, Wh = sd[f"{prefix}.weight_ih_l{l}{suffix}"], sd[f"{prefix}.weight_hh_l{l}{suffix}"]
            bi, bh = sd[f"{prefix}.bias_ih_l{l}{suffix}"], sd[f"{prefix}.bias_hh_l{l}{suffix}"]
            pre = x @ Wi.t() + bi
            h = x.new_zeros(B, Wh.shape[0])
            hs = [None] * S
            order = range(S - 1, -1, -1) if rev else range(S)
            for t in order:
                h = torch.tanh(pre[:, t] + h @ Wh.t() + bh)
                hs[t] = h
            outs.append(torch.stack(hs, 1))
        x = torch.cat(outs, -1)
    return x


def gru(sd: SD, prefix: str, x: torch.Tensor, num_layers: int) -> torch.Tensor:
    """torch.nn.GRU, batch_first (decoders/utils/rnn.py:27-36); gate order in the packed weights is (r, z, n):
    r = s(W_ir x + b_ir + W_hr h + b_hr), z likewise, n = tanh(W_in x + b_in + r*(W_hn h + b_hn)), h' = (1-z) n + z h."""
    B, S, _ = x.shape
    for l in range(num_layers):
        Wi, Wh = sd[f"{prefix}.weight_ih_l{l}"], sd[f"{prefix}.weight_hh_l{l}"]
        bi, bh = sd[f"{prefix}.bias_ih_l{l}"], sd[f"{prefix}.bias_hh_l{l}"]
        H = Wh.shape[1]
        h = x.new_zeros(B, H)
        gi_all = x @ Wi.t() + bi
        outs = []
        for t in range(S):
            gi, gh = gi_all[:, t], h @ Wh.t() + bh
            r = torch.sigmoid(gi[:, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, H : 2 * H] + gh[:, H : 2 * H])
            n = torch.tanh(gi[:, 2 * H :] + r * gh[:, 2 * H :])
            h = (1 - z) * n + z * h
            outs.append(h)
        x = torch.stack(outs, 1)
    return x


# ----------------------------------------------------------------------------------------------------------------------
# latent plan distribution
# ----------------------------------------------------------------------------------------------------------------------
def categorical_sample_indices(logits: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """Inverse-CDF categorical sample: index = #{j : cumsum(p)_j <= u}, clamped.  logits (B,C,K), u (B,C) in [0,1).
    (The reference draws with torch.multinomial, hulc.py:289 via distributions.py:23-27; bitwise RNG parity with
    torch is not a goal, so parity runs inject either `u` or the resulting indices.)"""
    p = torch.softmax(logits, -1)
    c = torch.cumsum(p, -1)
    idx = (c <= u.unsqueeze(-1)).sum(-1)
    return idx.clamp(max=logits.shape[-1] - 1)


def discrete_rsample(logits: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """OneHotCategoricalStraightThrough.rsample (torch one_hot_categorical.py): onehot(idx) + p - p.detach()."""
    p = torch.softmax(logits, -1)
    return F.one_hot(idx, logits.shape[-1]).to(p) + (p - p.detach())


def kl_categorical(logits_p: torch.Tensor, logits_q: torch.Tensor) -> torch.Tensor:
    """KL(Independent(OneHotCategorical(p),1) || ...(q)) per batch row = sum_cat sum_cls p (log p - log q)
    (torch/distributions/kl.py _kl_categorical_categorical + _kl_independent_independent)."""
    lp = torch.log_softmax(logits_p, -1)
    lq = torch.log_softmax(logits_q, -1)
    return (lp.exp() * (lp - lq)).sum(-1).sum(-1)


def kl_normal(m_p, s_p, m_q, s_q) -> torch.Tensor:
    """KL(N(m_p,s_p) || N(m_q,s_q)) summed over the last dim (torch/distributions/kl.py _kl_normal_normal)."""
    var_ratio = (s_p / s_q) ** 2
    t1 = ((m_p - m_q) / s_q) ** 2
    return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum(-1)


def cont_state(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Distribution.forward_dist, continuous (distributions.py:55-59): mean, softplus(var) + 1e-4."""
    mean, var = torch.chunk(x, 2, dim=-1)
    return mean, F.softplus(var) + 1e-4


def kl_loss_discrete(pp_logit: torch.Tensor, pr_logit: torch.Tensor, n_cat: int, n_cls: int, kl_beta: float, alpha: float) -> torch.Tensor:
    """Hulc.compute_kl_loss (hulc.py:539-561): beta * (alpha * KL(sg(post)||prior) + (1-alpha) * KL(post||sg(prior)))."""
    pp = pp_logit.view(-1, n_cat, n_cls)
    pr = pr_logit.view(-1, n_cat, n_cls)
    lhs = kl_categorical(pr.detach(), pp).mean()
    rhs = kl_categorical(pr, pp.detach()).mean()
    return kl_beta * (alpha * lhs + (1 - alpha) * rhs)


def kl_loss_continuous(pp_state, pr_state, kl_beta: float, alpha: float) -> torch.Tensor:
    (mp, sp), (mr, sr) = cont_state(pp_state), cont_state(pr_state)
    lhs = kl_normal(mr.detach(), sr.detach(), mp, sp).mean()
    rhs = kl_normal(mr, sr, mp.detach(), sp.detach()).mean()
    return kl_beta * (alpha * lhs + (1 - alpha) * rhs)


# ----------------------------------------------------------------------------------------------------------------------
# action decoder
# ----------------------------------------------------------------------------------------------------------------------
def euler_xyz_to_matrix(e: torch.Tensor) -> torch.Tensor:
    """euler_angles_to_matrix(e, "XYZ") = Rx(e0) Ry(e1) Rz(e2) (pytorch3d_transforms.py:162-218)."""
    a, b, c = e.unbind(-1)
    ca, sa, cb, sb, cc, sc = a.cos(), a.sin(), b.cos(), b.sin(), c.cos(), c.sin()
    one, zero = torch.ones_like(a), torch.zeros_like(a)
    Rx = torch.stack([one, zero, zero, zero, ca, -sa, zero, sa, ca], -1).view(*a.shape, 3, 3)
    Ry = torch.stack([cb, zero, sb, zero, one, zero, -sb, zero, cb], -1).view(*a.shape, 3, 3)
    Rz = torch.stack([cc, -sc, zero, sc, cc, zero, zero, zero, one], -1).view(*a.shape, 3, 3)
    return (Rx @ Ry) @ Rz


def matrix_to_euler_xyz(M: torch.Tensor) -> torch.Tensor:
    """matrix_to_euler_angles(M, "XYZ") (pytorch3d_transforms.py:221-303) resolved for this convention:
    (atan2(-M12, M22), asin(M02), atan2(-M01, M00))."""
    return torch.stack(
        [torch.atan2(-M[..., 1, 2], M[..., 2, 2]), torch.asin(M[..., 0, 2]), torch.atan2(-M[..., 0, 1], M[..., 0, 0])], -1
    )


def world_to_tcp_frame(action: torch.Tensor, robot_obs: torch.Tensor) -> torch.Tensor:
    """gripper_control.py:16-36: rotate the relative position into the TCP frame, re-express the (0.01-scaled)
    relative euler rotation in the TCP frame, wrap to (-pi, pi], scale back by 100, keep the gripper action."""
    b, s, _ = action.shape
    R = euler_xyz_to_matrix(robot_obs[..., 3:6]).float().view(-1, 3, 3)
    Rinv = torch.linalg.inv(R)
    pos = Rinv @ action[..., :3].reshape(-1, 3, 1)
    Rn = euler_xyz_to_matrix(robot_obs[..., 3:6] + action[..., 3:6] * 0.01).float().view(-1, 3, 3)
    orn = matrix_to_euler_xyz(torch.linalg.inv(Rn) @ R).float()
    orn = torch.where(orn < -math.pi, orn + 2 * math.pi, orn)
    orn = torch.where(orn > math.pi, orn - 2 * math.pi, orn)
    orn = orn * 100
    out = torch.cat([pos.view(b, s, -1), orn.view(b, s, -1), action[..., -1:]], -1)
    assert not torch.any(out.isnan())
    return out


def tcp_to_world_frame(action: torch.Tensor, robot_obs: torch.Tensor) -> torch.Tensor:
    """gripper_control.py:39-63: the inverse of world_to_tcp_frame, applied to sampled actions on the validation path."""
    b, s, _ = action.shape
    R = euler_xyz_to_matrix(robot_obs[..., 3:6]).float().view(-1, 3, 3)
    pos = R @ action[..., :3].reshape(-1, 3, 1)
    Rrel = euler_xyz_to_matrix(action[..., 3:6] * 0.01).float().view(-1, 3, 3)
    orn = matrix_to_euler_xyz(R @ torch.linalg.inv(Rrel)).float() - robot_obs[..., 3:6].reshape(-1, 3)
    orn = torch.where(orn < -math.pi, orn + 2 * math.pi, orn)
    orn = torch.where(orn > math.pi, orn - 2 * math.pi, orn)
    orn = orn * 100
    out = torch.cat([pos.view(b, s, -1), orn.view(b, s, -1), action[..., -1:]], -1)
    assert not torch.any(out.isnan())
    return out


def logistic_mixture_sample(logit_probs, log_scales, means, gripper_act, u_mix, u_inv, gripper_bounds=(-1.0, 1.0)) -> torch.Tensor:
    """LogisticDecoderRNN._sample (logistic_decoder_rnn.py:234-258) with the two torch.rand draws passed in: Gumbel-max choice of the
    mixture component, inverse-CDF sample of that logistic, gripper command = gripper_bounds[argmax]."""
    r1, r2 = 1e-5, 1.0 - 1e-5
    temp = logit_probs - torch.log(-torch.log((r1 - r2) * u_mix + r2))
    dist = torch.nn.functional.one_hot(torch.argmax(temp, -1), logit_probs.shape[-1]).to(means.dtype)
    ls = (dist * log_scales).sum(-1)
    mu = (dist * means).sum(-1)
    u = (r1 - r2) * u_inv + r2
    actions = mu + torch.exp(ls) * (torch.log(u) - torch.log(1.0 - u))
    if gripper_act is None:
        return actions
    cmd = torch.tensor(gripper_bounds, dtype=means.dtype, device=means.device)[gripper_act.argmax(-1)]
    return torch.cat([actions, cmd.unsqueeze(-1)], 2)


def validation_metrics(sample_act: torch.Tensor, actions: torch.Tensor):
    """lmp_val (hulc.py:346-385): per-sequence L1 error of the six continuous dims (mean over time) and the gripper success rate."""
    mae = (sample_act[..., :-1] - actions[..., :-1]).abs().mean(1)
    grip = torch.where(sample_act[..., -1] > 0, 1.0, -1.0)
    return mae, (actions[..., -1] == grip).float().mean()


def decoder_heads(sd: SD, h: torch.Tensor, out_features: int, n_mix: int, log_scale_min: float, discrete_gripper: bool, prefix: str = "action_decoder"):
    """LogisticDecoderRNN.forward after the RNN (logistic_decoder_rnn.py:278-287)."""
    B, S, _ = h.shape
    probs = linear(sd, f"{prefix}.prob_fc", h).view(B, S, out_features, n_mix)
    means = linear(sd, f"{prefix}.mean_fc", h).view(B, S, out_features, n_mix)
    log_scales = torch.clamp(linear(sd, f"{prefix}.log_scale_fc", h), min=log_scale_min).view(B, S, out_features, n_mix)
    grip = linear(sd, f"{prefix}.gripper_fc", h) if discrete_gripper else None
    return probs, log_scales, means, grip


def logistic_mixture_nll(logit_probs, log_scales, means, actions, act_min: float, act_max: float, num_classes: int, log_scale_min: float) -> torch.Tensor:
    """LogisticDecoderRNN._logistic_loss (logistic_decoder_rnn.py:184-231) for uniform scalar bounds."""
    log_scales = torch.clamp(log_scales, min=log_scale_min)
    a = actions.unsqueeze(-1).expand_as(means)
    centered = a - means
    inv_std = torch.exp(-log_scales)
    half_bin = (act_max - act_min) / 2.0 / (num_classes - 1)
    plus_in = inv_std * (centered + half_bin)
    min_in = inv_std * (centered - half_bin)
    cdf_plus, cdf_min = torch.sigmoid(plus_in), torch.sigmoid(min_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    mid_in = inv_std * centered
    log_pdf_mid = mid_in - log_scales - 2.0 * F.softplus(mid_in)
    cdf_delta = cdf_plus - cdf_min
    log_probs = torch.where(
        a < act_min + 1e-3,
        log_cdf_plus,
        torch.where(
            a > act_max - 1e-3,
            log_one_minus_cdf_min,
            torch.where(cdf_delta > 1e-5, torch.log(torch.clamp(cdf_delta, min=1e-12)), log_pdf_mid - math.log((num_classes - 1) / 2)),
        ),
    )
    log_probs = log_probs + torch.log_softmax(logit_probs, dim=-1)
    return -torch.logsumexp(log_probs, dim=-1).sum(-1).mean()


def decoder_loss(logit_probs, log_scales, means, grip, actions, *, discrete_gripper: bool, gripper_alpha: float, num_classes: int, log_scale_min: float):
    """LogisticDecoderRNN._loss (logistic_decoder_rnn.py:136-155)."""
    if discrete_gripper:
        nll = logistic_mixture_nll(logit_probs, log_scales, means, actions[..., :-1], -1.0, 1.0, num_classes, log_scale_min)
        label = (actions[..., -1] != -1).long().view(-1)  # -1 -> class 0, +1 -> class 1
        ce = F.cross_entropy(grip.reshape(-1, 2), label)
        return nll + gripper_alpha * ce
    return logistic_mixture_nll(logit_probs, log_scales, means, actions, -1.0, 1.0, num_classes, log_scale_min)


def clip_loss(sd: SD, seq_feat: torch.Tensor, goal: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Hulc.clip_auxiliary_loss (hulc.py:650-695) + ProjVisLang.forward (proj_vis_lang.py:23-27)."""
    skip = False
    if mask is not None:
        if not bool(mask.any()):
            skip, seq_feat, goal = True, seq_feat[0:1], goal[0:1]
        else:
            seq_feat, goal = seq_feat[mask], goal[mask]
    im = linear(sd, "proj_vis_lang.mlp_im.2", F.relu(linear(sd, "proj_vis_lang.mlp_im.0", seq_feat)))
    tx = linear(sd, "proj_vis_lang.mlp_lang.2", F.relu(linear(sd, "proj_vis_lang.mlp_lang.0", goal)))
    im = im / im.norm(dim=-1, keepdim=True)
    tx = tx / tx.norm(dim=-1, keepdim=True)
    logits = sd["logit_scale"].exp() * im @ tx.t()
    labels = torch.arange(logits.shape[0], device=logits.device)
    loss = (F.cross_entropy(logits, labels) + F.cross_entropy(logits.t(), labels)) / 2
    return loss * 0 if skip else loss


def bc_z_loss(sd: SD, seq_feat: torch.Tensor, gt_lang: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Hulc.bc_z_auxiliary_loss (hulc.py:567-604) with BCZLangDecoder (bc_z_lang_decoder.py:5-20): regress the language embedding from the
    sequence feature, mean cosine distance.  The reference indexes `seq_vis_feat` with the mask TWICE (hulc.py:592-594), which only works
    when the mask keeps every sequence; an all-false mask runs a dummy forward on the first sequence and multiplies the loss by 0."""
    if mask is not None:
        if not bool(mask.any()):
            return 0.0 * bc_z_loss(sd, seq_feat[0:1], gt_lang[0:1])
        if not bool(mask.all()):
            raise IndexError("the reference applies use_for_aux_lang_loss twice to seq_vis_feat (hulc.py:592-594): partial masks fail there")
    pred = linear(sd, "bc_z_lang_decoder.mlp.2", torch.relu(linear(sd, "bc_z_lang_decoder.mlp.0", seq_feat)))
    cos = (pred * gt_lang).sum(-1) / (torch.linalg.norm(pred, dim=1) * torch.linalg.norm(gt_lang, dim=1))
    return (1 - cos).mean()


def mia_loss(sd: SD, seq_feat: torch.Tensor, goal: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Hulc.mia_auxiliary_loss (hulc.py:606-648) with MIALangDiscriminator (mia_lang_discriminator.py:5-21): a discriminator over
    [projected image feature | projected language feature] scores matching pairs against pairs whose language row is rolled by one
    (torch.roll(lang, 1, 0)); binary cross entropy with logits over the 2B scores."""
    if mask is not None:
        if not bool(mask.any()):
            return 0.0 * mia_loss(sd, seq_feat[0:1], goal[0:1])
        seq_feat, goal = seq_feat[mask], goal[mask]
    im = linear(sd, "proj_vis_lang.mlp_im.2", torch.relu(linear(sd, "proj_vis_lang.mlp_im.0", seq_feat)))
    tx = linear(sd, "proj_vis_lang.mlp_lang.2", torch.relu(linear(sd, "proj_vis_lang.mlp_lang.0", goal)))

    def disc(a, b):
        return linear(sd, "mia_lang_discriminator.mlp.3", torch.relu(linear(sd, "mia_lang_discriminator.mlp.0", torch.cat([a, b], -1))))

    pos, neg = disc(im, tx), disc(im, torch.roll(tx, shifts=1, dims=0))
    pred = torch.cat([pos, neg], 0)
    labels = torch.cat([torch.ones_like(pos), torch.zeros_like(neg)], 0)
    return F.binary_cross_entropy_with_logits(pred, labels)


def random_shifts_aug(frames_u8: torch.Tensor, shifts: torch.Tensor, pad: int, mean: float = 0.5, std: float = 0.5) -> torch.Tensor:
    """The training-time image pipeline of conf/datamodule/transforms/rand_shift.yaml:2-22 on uint8 frames [N, C, H, W]: RandomShiftsAug
    (hulc/utils/transforms.py:8-29) -> ScaleImageTensor (/255) -> Normalize(mean, std).  The reference replicate-pads by `pad` and bilinearly
    samples a grid displaced by shifts[n] = (sx, sy) whole pixels of the PADDED image (align_corners=False puts every sample on a pixel
    centre), i.e. it crops the padded frame at offset (sy, sx): out[y][x] = in[clamp(y + sy - pad)][clamp(x + sx - pad)]."""
    N, C, H, W = frames_u8.shape
    ys = (torch.arange(H).view(1, H) + shifts[:, 1].view(N, 1) - pad).clamp(0, H - 1)  # [N, H]
    xs = (torch.arange(W).view(1, W) + shifts[:, 0].view(N, 1) - pad).clamp(0, W - 1)  # [N, W]
    x = frames_u8.float()
    x = torch.gather(x, 2, ys.view(N, 1, H, 1).expand(N, C, H, W))
    x = torch.gather(x, 3, xs.view(N, 1, 1, W).expand(N, C, H, W))
    return ((x / 255.0) - mean) / std


# ----------------------------------------------------------------------------------------------------------------------
# the training step
# ----------------------------------------------------------------------------------------------------------------------
def training_step(
    sd: SD,
    batch: Dict[str, Dict],
    *,
    model: str = "hulc",
    rnn_model: str = "rnn_decoder",
    dropout_p: float = 0.0,
    plan_idx: Optional[Dict[str, torch.Tensor]] = None,
    plan_u: Optional[Dict[str, torch.Tensor]] = None,
    plan_eps: Optional[Dict[str, torch.Tensor]] = None,
    dropout_masks: Optional[Dict[str, Dict[str, torch.Tensor]]] = None,
    kl_beta: float = 0.01,
    kl_alpha: float = 0.8,
    clip_beta: float = 3.0,
    bc_z_beta: Optional[float] = None,
    mia_beta: Optional[float] = None,
) -> Dict[str, torch.Tensor]:
    """Hulc.training_step (hulc.py:390-537) + lmp_train (hulc.py:254-299); GCBC.training_step (gcbc.py:50-181) for
    model="gcbc"; conf/model/mcil.yaml for model="mcil".  Returns every loss plus the intermediates the kernel tests
    compare against.  Randomness is injected per modality: `plan_idx[m]` (B,32) sampled class indices or `plan_u[m]`
    (B,32) uniforms (discrete latent), `plan_eps[m]` (B,256) normals (continuous latent), `dropout_masks[m]`."""
    out: Dict[str, torch.Tensor] = {}
    n_mod = len(batch)
    tot = kl_tot = act_tot = 0.0
    clip = torch.zeros(())
    discrete = model != "mcil"
    for m, d in batch.items():
        emb = perceptual_encoder(sd, d["rgb_obs"]["rgb_static"], d["rgb_obs"]["rgb_gripper"])
        goal = goal_encoder(sd, "language_goal", d["lang"]) if "lang" in m else goal_encoder(sd, "visual_goal", emb[:, -1])
        out[f"emb_{m}"], out[f"goal_{m}"] = emb, goal
        B, S, _ = emb.shape
        actions, robot_obs = d["actions"], d["state_info"]["robot_obs"]
        masks = (dropout_masks or {}).get(m)
        if model == "mcil":
            pr_h = birnn_tanh(sd, "plan_recognition.birnn_model", emb)
            seq_feat = pr_h[:, -1]
            pr_state = linear(sd, "plan_recognition.fc_state.0", seq_feat)
        else:
            pr_state, seq_feat = plan_recognition_transformer(sd, emb, p=dropout_p, masks=masks)
        out[f"pr_state_{m}"], out[f"seq_feat_{m}"] = pr_state, seq_feat
        if model == "gcbc":
            plan = emb.new_zeros(B, 0)
            kl = torch.zeros(())
        else:
            pp_state = plan_proposal(sd, emb[:, 0], goal)
            out[f"pp_state_{m}"] = pp_state
            if discrete:
                lg = pr_state.view(B, 32, 32)
                idx = plan_idx[m] if plan_idx is not None else categorical_sample_indices(lg, plan_u[m])
                out[f"plan_idx_{m}"] = idx
                plan = discrete_rsample(lg, idx).flatten(-2)
                kl = kl_loss_discrete(pp_state, pr_state, 32, 32, kl_beta, kl_alpha)
            else:
                mean, std = cont_state(pr_state)
                plan = mean + std * plan_eps[m]
                kl = kl_loss_continuous(pp_state, pr_state, kl_beta, kl_alpha)
        # LogisticDecoderRNN.forward (logistic_decoder_rnn.py:260-287)
        percep = emb[..., 64:128] if model != "mcil" else emb
        x = torch.cat([plan.unsqueeze(1).expand(-1, S, -1), percep, goal.unsqueeze(1).expand(-1, S, -1)], -1)
        h = gru(sd, "action_decoder.rnn", x, 2) if rnn_model == "gru_decoder" else elman_rnn(sd, "action_decoder.rnn", x, 2, "relu")
        dg = model != "mcil"
        lp, ls, mu, grip = decoder_heads(sd, h, 6 if dg else 7, 10, -7.0, dg)
        out[f"logit_probs_{m}"], out[f"log_scales_{m}"], out[f"means_{m}"] = lp, ls, mu
        if dg:
            out[f"gripper_act_{m}"] = grip
            act_t = world_to_tcp_frame(actions, robot_obs)
            out[f"actions_tcp_{m}"] = act_t
        else:
            act_t = actions
        act_loss = decoder_loss(lp, ls, mu, grip, act_t, discrete_gripper=dg, gripper_alpha=1.0, num_classes=10 if dg else 256, log_scale_min=-7.0)
        out[f"action_loss_{m}"], out[f"kl_loss_{m}"] = act_loss, kl
        tot = tot + act_loss + kl
        kl_tot, act_tot = kl_tot + kl, act_tot + act_loss
        if "lang" in m and model != "mcil":
            clip = clip + clip_loss(sd, seq_feat, goal, d.get("use_for_aux_lang_loss"))
            if bc_z_beta is not None:
                out["lang_pred_loss"] = out.get("lang_pred_loss", 0.0) + bc_z_loss(sd, seq_feat, d["lang"], d.get("use_for_aux_lang_loss"))
            if mia_beta is not None:
                out["lang_contrastive_loss"] = out.get("lang_contrastive_loss", 0.0) + mia_loss(sd, seq_feat, goal, d.get("use_for_aux_lang_loss"))
    total = tot / n_mod
    if bc_z_beta is not None and "lang_pred_loss" in out:  # hulc.py:500-509
        total = total + bc_z_beta * out["lang_pred_loss"]
    if mia_beta is not None and "lang_contrastive_loss" in out:  # hulc.py:510-519
        total = total + mia_beta * out["lang_contrastive_loss"]
    if model != "mcil":
        total = total + clip_beta * clip
        out["lang_clip_loss"] = clip
    out["total_loss"], out["kl_loss"], out["action_loss"] = total, kl_tot / n_mod, act_tot / n_mod
    return out


@torch.no_grad()
def validation_step(sd: SD, batch: Dict[str, Dict], *, model: str = "hulc", rnn_model: str = "rnn_decoder", plan_idx=None, plan_eps=None, sample_u=None,
                    kl_beta: float = 0.01, kl_alpha: float = 0.8) -> Dict[str, torch.Tensor]:
    """Hulc.validation_step / lmp_val (hulc.py:301-388, 739-841), eval mode (no dropout).  For the plan sampled from the proposal ("pp") and
    from the recognition network ("pr"): action loss, actions sampled from the mixture (LogisticDecoderRNN.loss_and_act,
    logistic_decoder_rnn.py:85-102) and their MAE / gripper success rate; plus KL and the CLIP loss of the language modality.
    plan_idx[which][m] (B,32) / plan_eps[which][m] (B,256) and sample_u[which][m] = (u_mix, u_inv) inject the randomness."""
    out: Dict[str, torch.Tensor] = {}
    discrete = model != "mcil"
    dg = model != "mcil"
    for m, d in batch.items():
        emb = perceptual_encoder(sd, d["rgb_obs"]["rgb_static"], d["rgb_obs"]["rgb_gripper"])
        goal = goal_encoder(sd, "language_goal", d["lang"]) if "lang" in m else goal_encoder(sd, "visual_goal", emb[:, -1])
        B, S, _ = emb.shape
        actions, robot_obs = d["actions"], d["state_info"]["robot_obs"]
        if model == "mcil":
            seq_feat = birnn_tanh(sd, "plan_recognition.birnn_model", emb)[:, -1]
            pr_state = linear(sd, "plan_recognition.fc_state.0", seq_feat)
        else:
            pr_state, seq_feat = plan_recognition_transformer(sd, emb, p=0.0, masks=None)
        if model == "gcbc":  # GCBC.validation_step (gcbc.py:183-281): one decoder pass on an empty plan, no KL
            x = torch.cat([emb[..., 64:128], goal.unsqueeze(1).expand(-1, S, -1)], -1)
            h = gru(sd, "action_decoder.rnn", x, 2) if rnn_model == "gru_decoder" else elman_rnn(sd, "action_decoder.rnn", x, 2, "relu")
            lp, ls, mu, grip = decoder_heads(sd, h, 6, 10, -7.0, True)
            u_mix, u_inv = sample_u["pr"][m]
            pred_w = tcp_to_world_frame(logistic_mixture_sample(lp, ls, mu, grip, u_mix, u_inv), robot_obs)
            out[f"action_loss_{m}"] = decoder_loss(lp, ls, mu, grip, world_to_tcp_frame(actions, robot_obs), discrete_gripper=True, gripper_alpha=1.0,
                                                   num_classes=10, log_scale_min=-7.0)
            out[f"sample_act_{m}"] = pred_w
            out[f"mae_{m}"], out[f"gripper_sr_{m}"] = validation_metrics(pred_w, actions)
            if "lang" in m:
                out["val_pred_clip_loss"] = clip_loss(sd, seq_feat, goal, d.get("use_for_aux_lang_loss"))
            continue
        pp_state = plan_proposal(sd, emb[:, 0], goal)
        for which, state in (("pp", pp_state), ("pr", pr_state)):
            if discrete:
                plan = torch.nn.functional.one_hot(plan_idx[which][m].long(), 32).to(emb.dtype).flatten(-2)
            else:
                mean, std = cont_state(state)
                plan = mean + std * plan_eps[which][m]
            percep = emb[..., 64:128] if model != "mcil" else emb
            x = torch.cat([plan.unsqueeze(1).expand(-1, S, -1), percep, goal.unsqueeze(1).expand(-1, S, -1)], -1)
            h = gru(sd, "action_decoder.rnn", x, 2) if rnn_model == "gru_decoder" else elman_rnn(sd, "action_decoder.rnn", x, 2, "relu")
            lp, ls, mu, grip = decoder_heads(sd, h, 6 if dg else 7, 10, -7.0, dg)
            u_mix, u_inv = sample_u[which][m]
            pred = logistic_mixture_sample(lp, ls, mu, grip, u_mix, u_inv)
            act_t = world_to_tcp_frame(actions, robot_obs) if dg else actions
            out[f"action_loss_{which}_{m}"] = decoder_loss(lp, ls, mu, grip, act_t, discrete_gripper=dg, gripper_alpha=1.0, num_classes=10 if dg else 256,
                                                           log_scale_min=-7.0)
            pred_w = tcp_to_world_frame(pred, robot_obs) if dg else pred
            out[f"sample_act_{which}_{m}"] = pred_w
            out[f"mae_{which}_{m}"], out[f"gripper_sr_{which}_{m}"] = validation_metrics(pred_w, actions)
            out[f"sampled_plan_{which}_{m}"] = plan
        out[f"kl_loss_{m}"] = kl_loss_discrete(pp_state, pr_state, 32, 32, kl_beta, kl_alpha) if discrete else kl_loss_continuous(pp_state, pr_state, kl_beta, kl_alpha)
        if "lang" in m and model != "mcil":
            out["val_pred_clip_loss"] = clip_loss(sd, seq_feat, goal, d.get("use_for_aux_lang_loss"))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# inference: Hulc.step / get_pp_plan_{lang,vision} / predict_with_plan (hulc.py:851-957), LogisticDecoderRNN.act (:104-119)
# ----------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def inference_plan(sd: SD, rgb_static: torch.Tensor, rgb_gripper: torch.Tensor, *, lang: Optional[torch.Tensor] = None, plan_idx: Optional[torch.Tensor] = None,
                   plan_eps: Optional[torch.Tensor] = None, model: str = "hulc"):
    """get_pp_plan_lang (T = 1 frame, `lang` (1,384)) / get_pp_plan_vision (T = 2: observation + goal image; the goal is encoded from the last
    frame): latent goal and a plan drawn from the proposal network (class indices `plan_idx` (1,32) or Normal noise `plan_eps` (1,256)
    injected); GCBC (gcbc.py:287-317): the goal only, an empty plan.  rgb_*: (T,3,H,W)."""
    emb = perceptual_encoder(sd, rgb_static[None], rgb_gripper[None])
    goal = goal_encoder(sd, "language_goal", lang) if lang is not None else goal_encoder(sd, "visual_goal", emb[:, -1])
    if model == "gcbc":
        return emb.new_zeros(1, 0), goal, None
    pp_state = plan_proposal(sd, emb[:, 0], goal)
    if model == "mcil":
        mean, std = cont_state(pp_state)
        plan = mean + std * plan_eps
    else:
        plan = torch.nn.functional.one_hot(plan_idx.long(), 32).to(emb.dtype).flatten(-2)
    return plan, goal, pp_state


@torch.no_grad()
def inference_act(sd: SD, rgb_static, rgb_gripper, robot_obs_raw, plan, goal, hidden, u_mix, u_inv, *, model: str = "hulc", rnn_model: str = "rnn_decoder"):
    """predict_with_plan -> LogisticDecoderRNN.act: one step of the 2-layer decoder RNN (ReLU Elman cell or GRU) from the carried hidden state
    (2,1,H), sampled action mapped to the world frame (not for MCIL: no gripper control).  Returns (action (1,1,7), new hidden)."""
    emb = perceptual_encoder(sd, rgb_static[None], rgb_gripper[None])
    dg = model != "mcil"
    x = torch.cat([plan, emb[:, 0, 64:128] if dg else emb[:, 0], goal], -1)
    new_hidden = []
    p = "action_decoder.rnn"
    for l in range(2):
        Wi, Wh, bi, bh = sd[f"{p}.weight_ih_l{l}"], sd[f"{p}.weight_hh_l{l}"], sd[f"{p}.bias_ih_l{l}"], sd[f"{p}.bias_hh_l{l}"]
        if rnn_model == "gru_decoder":
            H = Wh.shape[1]
            gi, gh = x @ Wi.t() + bi, hidden[l] @ Wh.t() + bh
            r = torch.sigmoid(gi[:, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, H : 2 * H] + gh[:, H : 2 * H])
            n = torch.tanh(gi[:, 2 * H :] + r * gh[:, 2 * H :])
            x = (1 - z) * n + z * hidden[l]
        else:
            x = torch.relu(x @ Wi.t() + bi + hidden[l] @ Wh.t() + bh)
        new_hidden.append(x)
    lp, ls, mu, grip = decoder_heads(sd, x[:, None], 6 if dg else 7, 10, -7.0, dg)
    pred = logistic_mixture_sample(lp, ls, mu, grip, u_mix, u_inv)
    return (tcp_to_world_frame(pred, robot_obs_raw.reshape(1, 1, -1)) if dg else pred), torch.stack(new_hidden)
