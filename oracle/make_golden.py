"""Generate `tests/golden/*.npz` by running the UNMODIFIED reference and pin the oracle against it.

Run in the build container only (needs `/root/reference`):

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden hulc_b2s8  # one case

For every case the script
  1. builds the reference `Hulc`/`GCBC` LightningModule from the resolved config tree (`synthetic.model_config`) under
     the dependency shims of `oracle/ref_stubs.py`,
  2. overwrites its parameters with `synthetic.fill_state_dict_`, builds the seeded batch (`synthetic.make_batch`),
  3. injects the run's randomness (categorical sample via a patched `torch.multinomial`, Normal noise via a patched
     `_standard_normal`, dropout keep-masks via patched `F.dropout`/`F.scaled_dot_product_attention`),
  4. runs `training_step` + `backward` of the reference, and the oracle on the same tensors,
  5. asserts oracle == reference (losses, logits, every parameter gradient) and writes the reference's numbers
     to the fixture.  Fixtures hold outputs only; inputs and weights are regenerated from seeds by the tests.
"""
from __future__ import annotations

import contextlib
import copy
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_stubs  # noqa: E402

ref_stubs.install()

from hulc_b200.utils import synthetic  # noqa: E402
from oracle import hulc_oracle as O  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

CASES = {
    # name: (model, rnn_model, B, S, dropout_p)
    "hulc_b2s8": ("hulc", "rnn_decoder", 2, 8, 0.0),  # BASELINE config 1
    "hulc_b2s8_drop": ("hulc", "rnn_decoder", 2, 8, 0.1),
    "hulc_gru_b2s8": ("hulc", "gru_decoder", 2, 8, 0.0),
    "gcbc_b2s8": ("gcbc", "rnn_decoder", 2, 8, 0.0),
    "mcil_b2s8": ("mcil", "rnn_decoder", 2, 8, 0.0),
    "hulc_b4s32": ("hulc", "rnn_decoder", 4, 32, 0.0),  # full window, small batch
    "hulc_b32s32": ("hulc", "rnn_decoder", 32, 32, 0.0),  # BASELINE config 2 shape
    "mcil_b32s32": ("mcil", "rnn_decoder", 32, 32, 0.0),  # BASELINE config 4 shape
    "gcbc_b32s64": ("gcbc", "rnn_decoder", 32, 64, 0.0),  # BASELINE config 5 shape (window 64)
    # ablation blocks: BC-Z + MIA auxiliary losses next to the CLIP loss.  (model/action_decoder=deterministic cannot be
    # pinned: DeterministicDecoder.__init__ raises NameError in the unmodified reference, deterministic_decoder.py:33.)
    "hulc_aux_b4s32": ("hulc", "rnn_decoder", 4, 32, 0.0, dict(bc_z=True, mia=True)),
}


def build_reference(model, rnn_model, dropout_p, max_window, **variant):
    import importlib

    cfg = synthetic.model_config(model, rnn_model, max_window, dropout_p, target_root="hulc", **variant)
    cfg = copy.deepcopy(cfg)
    target = cfg.pop("_target_")
    cfg.pop("_recursive_")
    mod, _, name = target.rpartition(".")
    cls = getattr(importlib.import_module(mod), name)
    torch.manual_seed(0)
    net = cls(**cfg)
    synthetic.fill_state_dict_(net.state_dict())
    net.train()
    return net


@contextlib.contextmanager
def injected_randomness(u_queue, eps_queue, mask_queue, p):
    """Route the reference's three RNG consumers to seeded tensors."""
    import torch.distributions.normal as tdn
    import torch.nn.functional as F

    orig_mn, orig_sn, orig_do, orig_sdpa = torch.multinomial, tdn._standard_normal, F.dropout, F.scaled_dot_product_attention

    def multinomial(probs_2d, n, replacement=False, **kw):
        u = u_queue.pop(0).reshape(-1, 1)
        assert n == 1 and u.shape[0] == probs_2d.shape[0]
        c = torch.cumsum(probs_2d, -1)
        return (c <= u).sum(-1, keepdim=True).clamp(max=probs_2d.shape[-1] - 1)

    def standard_normal(shape, dtype, device):
        e = eps_queue.pop(0)
        assert tuple(e.shape) == tuple(shape), (e.shape, shape)
        return e.to(dtype)

    def dropout(x, p_=0.5, training=True, inplace=False):
        if not training or p_ == 0.0:
            return x
        name, keep = mask_queue.pop(0)
        if keep.dim() == 3 and tuple(keep.shape) != tuple(x.shape):  # (B,S,D) recipe -> seq-first (S,B,D) reference
            keep = keep.permute(1, 0, 2)
        assert tuple(keep.shape) == tuple(x.shape), (name, keep.shape, x.shape)
        return x * keep.to(x.dtype) / (1.0 - p_)

    def sdpa(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, **kw):
        assert attn_mask is None and not is_causal
        a = torch.softmax(q @ k.transpose(-1, -2) / (q.shape[-1] ** 0.5), -1)
        return dropout(a, dropout_p, True) @ v

    torch.multinomial, tdn._standard_normal = multinomial, standard_normal
    F.dropout, F.scaled_dot_product_attention = dropout, sdpa
    try:
        yield
    finally:
        torch.multinomial, tdn._standard_normal = orig_mn, orig_sn
        F.dropout, F.scaled_dot_product_attention = orig_do, orig_sdpa


def run_case(name):
    model, rnn_model, B, S, p = CASES[name][:5]
    variant = CASES[name][5] if len(CASES[name]) > 5 else {}
    t0 = time.time()
    net = build_reference(model, rnn_model, p, max(32, S), **variant)
    batch = synthetic.make_batch(B, S, seed=1)
    noise = {m: synthetic.plan_noise(B, S, m) for m in batch}
    masks = {m: synthetic.dropout_masks(B, S, m, p) for m in batch} if p > 0 else None

    # --- reference -------------------------------------------------------------------------------------------------
    u_q, eps_q, m_q = [], [], []
    for m in batch:
        if model == "hulc":
            u_q.append(noise[m]["u"])
        if model == "mcil":
            eps_q.append(noise[m]["eps"])
        if masks is not None:
            order = ["in"] + [f"l{i}.{s}" for i in range(2) for s in ("attn", "drop1", "ffn", "drop2")]
            m_q += [(k, masks[m][k]) for k in order]

    captured = {}
    dec = net.action_decoder
    orig_loss = dec._loss

    def spy_loss(logit_probs, log_scales, means, gripper_act, actions):
        captured[net.modality_scope] = (logit_probs, log_scales, means, gripper_act, actions)
        return orig_loss(logit_probs, log_scales, means, gripper_act, actions)

    dec._loss = spy_loss
    with injected_randomness(u_q, eps_q, m_q, p):
        loss_ref = net.training_step(batch, 0)
    assert not u_q and not eps_q and not m_q, "injected randomness not fully consumed"
    loss_ref.backward()
    ref_grads = {k: (v.grad.detach().clone() if v.grad is not None else None) for k, v in net.named_parameters()}
    logged = net.logged

    # --- oracle ----------------------------------------------------------------------------------------------------
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and k in ref_grads) for k, v in net.state_dict().items()}
    out = O.training_step(
        sd, batch, model=model, rnn_model=rnn_model, dropout_p=p,
        plan_u={m: noise[m]["u"] for m in batch}, plan_eps={m: noise[m]["eps"] for m in batch}, dropout_masks=masks,
        bc_z_beta=1.0 if variant.get("bc_z") else None, mia_beta=1.0 if variant.get("mia") else None,
    )
    out["total_loss"].backward()

    def close(a, b, what, rtol=1e-4, atol=1e-5):
        if not torch.allclose(a, b, rtol=rtol, atol=atol):
            raise AssertionError(f"{name}: oracle != reference for {what}: max|d|={float((a - b).abs().max()):.3e}")

    close(out["total_loss"].detach(), loss_ref.detach(), "total_loss", 1e-5, 1e-6)
    fx = {"total_loss": loss_ref.detach()}
    for k in ("train/kl_loss", "train/action_loss", "train/lang_clip_loss", "train/kl_loss_scaled_vis", "train/kl_loss_scaled_lang",
              "train/action_loss_vis", "train/action_loss_lang", "train/pred_lang", "train/lang_contrastive"):
        if k in logged:
            fx[k.replace("train/", "")] = logged[k]
    close(out["action_loss"].detach(), logged["train/action_loss"], "action_loss", 1e-5, 1e-6)
    if model == "hulc":
        close(out["kl_loss"].detach(), logged["train/kl_loss"], "kl_loss", 1e-5, 1e-7)
    if model != "mcil":
        close(3.0 * out["lang_clip_loss"].detach(), logged["train/lang_clip_loss"], "clip", 1e-5, 1e-6)
    if variant.get("bc_z"):
        close(out["lang_pred_loss"].detach(), logged["train/pred_lang"], "bc-z loss", 1e-5, 1e-6)
    if variant.get("mia"):
        close(out["lang_contrastive_loss"].detach(), logged["train/lang_contrastive"], "mia loss", 1e-5, 1e-6)
    nseq = min(B, 2)
    for m in batch:
        lp, ls, mu, grip, act = captured[m]
        close(out[f"logit_probs_{m}"].detach(), lp.detach(), f"logit_probs_{m}")
        close(out[f"log_scales_{m}"].detach(), ls.detach(), f"log_scales_{m}")
        close(out[f"means_{m}"].detach(), mu.detach(), f"means_{m}")
        fx[f"logit_probs_{m}"], fx[f"log_scales_{m}"], fx[f"means_{m}"] = lp[:nseq].detach(), ls[:nseq].detach(), mu[:nseq].detach()
        if grip is not None:
            close(out[f"gripper_act_{m}"].detach(), grip.detach(), f"gripper_act_{m}")
            close(out[f"actions_tcp_{m}"], act, f"actions_tcp_{m}", 1e-5, 2e-5)
            fx[f"gripper_act_{m}"], fx[f"actions_tcp_{m}"] = grip[:nseq].detach(), act[:nseq].detach()
        if f"plan_idx_{m}" in out:
            fx[f"plan_idx_{m}"] = out[f"plan_idx_{m}"]
    worst = 0.0
    for k, g in ref_grads.items():
        og = sd[k].grad
        if g is None:
            assert og is None or float(og.abs().max()) == 0.0, f"{name}: oracle has a gradient for unused parameter {k}"
            fx[f"gradnorm/{k}"] = torch.tensor(-1.0)
            continue
        denom = float(g.norm()) + 1e-12
        rel = float((og - g).norm()) / denom
        worst = max(worst, rel)
        assert rel < 1e-3, f"{name}: oracle grad != reference grad for {k}: rel {rel:.3e}"
        fx[f"gradnorm/{k}"] = g.norm()
        fx[f"gradhead/{k}"] = g.reshape(-1)[:16].clone()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"{name}.npz", **{k: np.asarray(v.detach().cpu().numpy()) for k, v in fx.items()})
    print(f"[golden] {name}: total_loss={float(loss_ref):.6f} params={sum(p.numel() for p in net.parameters())} "
          f"worst oracle-vs-reference grad rel err={worst:.2e}  ({time.time() - t0:.1f}s)")


AUTOCAST_CASES = {
    # name: fp32 case whose inputs / sampled plan classes it shares.  The UNMODIFIED reference under torch.autocast(bfloat16) — the
    # yardstick of the bf16 path (BASELINE config 3; the reference trains at 16-bit precision, conf/trainer/play_trainer.yaml:3).
    "hulc_b2s8_bf16": "hulc_b2s8",
    "hulc_b4s32_bf16": "hulc_b4s32",
    "hulc_b32s32_bf16": "hulc_b32s32",  # BASELINE config 3's per-GPU batch
}


def run_autocast_case(name):
    """The reference's training_step under bf16 autocast on the same inputs, weights and sampled plan classes as the fp32 fixture.
    As on the GPU: the distribution objects get fp32 logits (bf16 logits fail OneHotCategorical's simplex check) and
    world_to_tcp_frame runs in fp32 (the reference forces that with its own autocast(dtype=float32) context, gripper_control.py:17).
    CPU autocast keeps softmax / layer_norm / loss arithmetic in bf16 where CUDA autocast would promote them to fp32, so the deviation
    recorded here is an upper bound of what the reference shows on a GPU."""
    import hulc.models.decoders.logistic_decoder_rnn as ldr

    base = AUTOCAST_CASES[name]
    model, rnn_model, B, S, p = CASES[base]
    assert p == 0.0 and model == "hulc"
    t0 = time.time()
    fp32 = dict(np.load(GOLDEN / f"{base}.npz"))
    net = build_reference(model, rnn_model, p, 32)
    batch = synthetic.make_batch(B, S, seed=1)
    idx_q = [torch.from_numpy(fp32[f"plan_idx_{m}"]).long() for m in batch]

    orig_mn, orig_get, orig_w2t = torch.multinomial, net.dist.get_dist, ldr.world_to_tcp_frame

    def multinomial(probs_2d, n, replacement=False, **kw):  # the fp32 run's sampled classes: the deviation measured is arithmetic, not a re-draw
        return idx_q.pop(0).reshape(-1, 1)

    def get_dist(state):
        return orig_get(type(state)(*[x.float() for x in state]))

    def w2t(action, robot_obs):
        with torch.autocast("cpu", enabled=False):
            return orig_w2t(action.float(), robot_obs.float())

    captured = {}
    dec = net.action_decoder
    orig_loss = dec._loss

    def spy_loss(logit_probs, log_scales, means, gripper_act, actions):
        captured[net.modality_scope] = (logit_probs, log_scales, means, gripper_act, actions)
        return orig_loss(logit_probs, log_scales, means, gripper_act, actions)

    dec._loss = spy_loss
    torch.multinomial, net.dist.get_dist, ldr.world_to_tcp_frame = multinomial, get_dist, w2t
    try:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss = net.training_step(batch, 0)
    finally:
        torch.multinomial, ldr.world_to_tcp_frame = orig_mn, orig_w2t
    assert not idx_q
    logged = net.logged
    fx = {"total_loss": loss.detach().float()}
    for k in ("train/kl_loss", "train/action_loss", "train/lang_clip_loss", "train/action_loss_vis", "train/action_loss_lang"):
        fx[k.replace("train/", "")] = logged[k].float()
    nseq = min(B, 2)
    for m in batch:
        lp, ls, mu, grip, act = captured[m]
        fx[f"logit_probs_{m}"], fx[f"log_scales_{m}"], fx[f"means_{m}"] = lp[:nseq].detach().float(), ls[:nseq].detach().float(), mu[:nseq].detach().float()
        fx[f"gripper_act_{m}"] = grip[:nseq].detach().float()
    np.savez_compressed(GOLDEN / f"{name}.npz", **{k: np.asarray(v.detach().cpu().numpy()) for k, v in fx.items()})
    dev = {k: abs(float(fx[k]) - float(fp32[k])) / abs(float(fp32[k])) for k in ("total_loss", "action_loss", "kl_loss", "lang_clip_loss")}
    rms = lambda a: float(np.sqrt(np.mean(np.square(a))))
    ldev = {k: rms(fx[f"{k}_vis"].numpy() - fp32[f"{k}_vis"]) / rms(fp32[f"{k}_vis"]) for k in ("logit_probs", "means", "log_scales")}
    print(f"[golden] {name}: reference under bf16 autocast vs its fp32 run: relative loss deviations {dev}; logits rms deviation / rms {ldev}  ({time.time() - t0:.1f}s)")


VAL_CASES = {
    # name: (model, rnn_model, B, S) — validation_step / lmp_val (hulc.py:301-388, 739-841), eval mode
    "val_hulc_b2s8": ("hulc", "rnn_decoder", 2, 8),
    "val_hulc_b4s32": ("hulc", "rnn_decoder", 4, 32),
    "val_mcil_b2s8": ("mcil", "rnn_decoder", 2, 8),
    "val_gcbc_b2s8": ("gcbc", "rnn_decoder", 2, 8),  # GCBC.validation_step (gcbc.py:183-281)
}


def run_val_case(name):
    """The reference's validation_step with its randomness injected (torch.multinomial / _standard_normal for the latent plan, torch.rand for
    LogisticDecoderRNN._sample), the oracle's restatement on the same tensors, and the fixture with the reference's numbers."""
    import types

    model, rnn_model, B, S = VAL_CASES[name]
    t0 = time.time()
    net = build_reference(model, rnn_model, 0.1, 32)
    net.eval()
    batch = synthetic.make_batch(B, S, seed=1)
    n_dims = 6 if model != "mcil" else 7
    noise = {w: {m: synthetic.validation_noise(B, S, m, w, n_dims=n_dims) for m in batch} for w in ("pp", "pr")}
    u_q, eps_q, rand_q = [], [], []
    for m in batch:
        for w in (("pp", "pr") if model != "gcbc" else ("pr",)):  # lmp_val samples from the proposal first (hulc.py:337-343), then from the recognition network (:360-366)
            if model == "hulc":
                u_q.append(noise[w][m]["u"])
            elif model == "mcil":
                eps_q.append(noise[w][m]["eps"])
            rand_q += [noise[w][m]["u_mix"], noise[w][m]["u_inv"]]
    net.trainer = types.SimpleNamespace(datamodule=types.SimpleNamespace(modalities=list(batch)))
    net.clip_groundtruth = lambda *a, **k: None  # needs the dataset's task annotations; a logging-only metric
    captured = {}
    orig_val = net.lmp_val

    def spy_val(*a, **k):
        r = orig_val(*a, **k)
        captured[net.modality_scope] = r
        return r

    if model != "gcbc":
        net.lmp_val = spy_val
    else:  # GCBC has no lmp_val: capture (loss, sampled actions) at the decoder
        orig_la = net.action_decoder.loss_and_act

        def spy_la(*a, **k):
            r = orig_la(*a, **k)
            captured[net.modality_scope] = (r[0].detach().clone(), r[1].detach().clone())
            return r

        net.action_decoder.loss_and_act = spy_la
    orig_rand = torch.rand

    def rand(*size, **kw):
        u = rand_q.pop(0)
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        assert tuple(u.shape) == shape, (u.shape, shape)
        return u.clone()

    orig_normal = torch.normal

    def normal(mean, std, *a, **kw):  # Normal.sample() (continuous latent) draws through torch.normal, not _standard_normal
        e = eps_q.pop(0)
        assert tuple(e.shape) == tuple(mean.shape), (e.shape, mean.shape)
        return mean + std * e

    torch.rand, torch.normal = rand, normal
    try:
        with injected_randomness(u_q, eps_q, [], 0.0), torch.no_grad():
            net.validation_step(batch, 0)
    finally:
        torch.rand, torch.normal = orig_rand, orig_normal
    assert not u_q and not eps_q and not rand_q, "injected randomness not fully consumed"
    logged = net.logged

    # --- oracle ----------------------------------------------------------------------------------------------------
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    plan_idx = {w: {} for w in ("pp", "pr")}
    fx = {}
    if model == "gcbc":
        out = O.validation_step(sd, batch, model=model, rnn_model=rnn_model, sample_u={"pr": {m: (noise["pr"][m]["u_mix"], noise["pr"][m]["u_inv"]) for m in batch}})
        for m in batch:
            loss, act = captured[m]
            assert torch.allclose(out[f"action_loss_{m}"], loss, rtol=1e-4, atol=1e-5), f"{name}: action_loss_{m}"
            assert torch.allclose(out[f"sample_act_{m}"], act, rtol=1e-4, atol=1e-4), f"{name}: sample_act_{m}"
            for key, okey in ((f"val_total_mae/{m}_total_mae", None), (f"val_grip/{m}_grip_sr", f"gripper_sr_{m}")):
                ov = out[f"mae_{m}"].mean() if okey is None else out[okey]
                assert torch.allclose(ov, logged[key], rtol=1e-4, atol=1e-5), f"{name}: {key}"
            fx[f"action_loss_{m}"], fx[f"mae_{m}"], fx[f"gripper_sr_{m}"] = loss, out[f"mae_{m}"], logged[f"val_grip/{m}_grip_sr"]
        assert torch.allclose(out["val_pred_clip_loss"], logged["val/val_pred_clip_loss"], rtol=1e-5, atol=1e-6)
        fx["val_pred_clip_loss"] = logged["val/val_pred_clip_loss"]
        for k, v in logged.items():
            if k.startswith("val"):
                fx["logged/" + k] = torch.as_tensor(v)
        np.savez_compressed(GOLDEN / f"{name}.npz", **{k: np.asarray(torch.as_tensor(v).detach().cpu().numpy()) for k, v in fx.items()})
        print(f"[golden] {name}: act_loss_vis={float(captured[list(batch)[0]][0]):.6f} keys={len(fx)}  ({time.time() - t0:.1f}s)")
        return
    for m in batch:
        pp_plan, _, pr_plan = captured[m][0], captured[m][1], captured[m][2]
        if model == "hulc":
            plan_idx["pp"][m] = pp_plan.view(B, 32, 32).argmax(-1)
            plan_idx["pr"][m] = pr_plan.view(B, 32, 32).argmax(-1)
            fx[f"plan_idx_pp_{m}"], fx[f"plan_idx_pr_{m}"] = plan_idx["pp"][m], plan_idx["pr"][m]
    out = O.validation_step(sd, batch, model=model, rnn_model=rnn_model, plan_idx=plan_idx,
                            plan_eps={w: {m: noise[w][m]["eps"] for m in batch} for w in ("pp", "pr")},
                            sample_u={w: {m: (noise[w][m]["u_mix"], noise[w][m]["u_inv"]) for m in batch} for w in ("pp", "pr")})

    def close(a, b, what, rtol=1e-4, atol=1e-5):
        if not torch.allclose(a, b, rtol=rtol, atol=atol):
            raise AssertionError(f"{name}: oracle != reference for {what}: max|d|={float((a - b).abs().max()):.3e}")

    for m in batch:
        (pp_plan, loss_pp, pr_plan, loss_pr, kl, mae_pp, mae_pr, sr_pp, sr_pr, _seq) = captured[m]
        close(out[f"action_loss_pp_{m}"], loss_pp, f"action_loss_pp_{m}")
        close(out[f"action_loss_pr_{m}"], loss_pr, f"action_loss_pr_{m}")
        close(out[f"kl_loss_{m}"], kl, f"kl_loss_{m}", 1e-5, 1e-7)
        close(out[f"mae_pp_{m}"], mae_pp, f"mae_pp_{m}", 1e-4, 1e-4)
        close(out[f"mae_pr_{m}"], mae_pr, f"mae_pr_{m}", 1e-4, 1e-4)
        close(out[f"gripper_sr_pp_{m}"], sr_pp, f"gripper_sr_pp_{m}")
        close(out[f"gripper_sr_pr_{m}"], sr_pr, f"gripper_sr_pr_{m}")
        close(out[f"sampled_plan_pp_{m}"], pp_plan, f"sampled_plan_pp_{m}")
        close(out[f"sampled_plan_pr_{m}"], pr_plan, f"sampled_plan_pr_{m}")
        fx.update({f"action_loss_pp_{m}": loss_pp, f"action_loss_pr_{m}": loss_pr, f"kl_loss_{m}": kl, f"mae_pp_{m}": mae_pp, f"mae_pr_{m}": mae_pr,
                   f"gripper_sr_pp_{m}": sr_pp, f"gripper_sr_pr_{m}": sr_pr})
    if "val/val_pred_clip_loss" in logged:
        close(out["val_pred_clip_loss"], logged["val/val_pred_clip_loss"], "val_pred_clip_loss", 1e-5, 1e-6)
        fx["val_pred_clip_loss"] = logged["val/val_pred_clip_loss"]
    for k, v in logged.items():
        if k.startswith("val"):
            fx["logged/" + k] = torch.as_tensor(v)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"{name}.npz", **{k: np.asarray(torch.as_tensor(v).detach().cpu().numpy()) for k, v in fx.items()})
    print(f"[golden] {name}: act_loss_pp_vis={float(captured[list(batch)[0]][1]):.6f} keys={len(fx)}  ({time.time() - t0:.1f}s)")


INFER_CASES = {
    # name: (goal kind, control steps, replan_freq) — Hulc.step (hulc.py:851-870): re-plans (and clears the decoder's hidden state) every replan_freq steps
    "infer_hulc_lang": ("lang", 7, 3),
    "infer_hulc_vision": ("vision", 5, 4),
}


def run_infer_case(name):
    """A rollout of the reference's Hulc.step with its randomness injected, the oracle's restatement on the same tensors, the fixture."""
    kind, T, replan = INFER_CASES[name]
    t0 = time.time()
    net = build_reference("hulc", "rnn_decoder", 0.1, 32)
    net.eval()
    net.replan_freq = replan
    r = synthetic.rollout_inputs(T, kind)
    net.lang_embeddings = {"the task": r["lang"].numpy()[None]}  # load_lang_embeddings stores (1, 1, 384) arrays per annotation (hulc.py:872-882)
    n_replans = (T + replan - 1) // replan
    u_q = [r["plan_u"][i : i + 1] for i in range(n_replans)]
    rand_q = []
    for t in range(T):
        rand_q += [r["u_mix"][t], r["u_inv"][t]]
    orig_rand = torch.rand

    def rand(*size, **kw):
        u = rand_q.pop(0)
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        assert tuple(u.shape) == shape, (u.shape, shape)
        return u.clone()

    def obs(t):
        return {"rgb_obs": {"rgb_static": r["rgb_static"][t][None, None], "rgb_gripper": r["rgb_gripper"][t][None, None]}, "depth_obs": {},
                "robot_obs": r["robot_obs"][t][None, None], "robot_obs_raw": r["robot_obs_raw"][t][None, None]}

    goal = "the task" if kind == "lang" else {"rgb_obs": {"rgb_static": r["goal_static"][None], "rgb_gripper": r["goal_gripper"][None]}, "depth_obs": {},
                                              "robot_obs": r["goal_robot_obs"][None]}
    plans, actions = [], []
    net.reset()
    torch.rand = rand
    try:
        with injected_randomness(u_q, [], [], 0.0), torch.no_grad():
            for t in range(T):
                a = net.step(obs(t), goal)
                actions.append(a.detach().clone().reshape(7))
                if t % replan == 0:
                    plans.append(net.plan.detach().clone().view(32, 32).argmax(-1))
    finally:
        torch.rand = orig_rand
    assert not u_q and not rand_q, "injected randomness not fully consumed"
    actions, plan_idx = torch.stack(actions), torch.stack(plans)

    # --- oracle ----------------------------------------------------------------------------------------------------
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    hidden = plan = goal_t = None
    for t in range(T):
        if t % replan == 0:
            if kind == "lang":
                plan, goal_t, _ = O.inference_plan(sd, r["rgb_static"][t : t + 1], r["rgb_gripper"][t : t + 1], lang=r["lang"], plan_idx=plan_idx[t // replan][None])
            else:
                plan, goal_t, _ = O.inference_plan(sd, torch.cat([r["rgb_static"][t : t + 1], r["goal_static"]]), torch.cat([r["rgb_gripper"][t : t + 1], r["goal_gripper"]]),
                                                   plan_idx=plan_idx[t // replan][None])
            hidden = torch.zeros(2, 1, 2048)
        a, hidden = O.inference_act(sd, r["rgb_static"][t : t + 1], r["rgb_gripper"][t : t + 1], r["robot_obs_raw"][t : t + 1], plan, goal_t, hidden, r["u_mix"][t], r["u_inv"][t])
        if not torch.allclose(a.reshape(7), actions[t], rtol=1e-4, atol=1e-4):
            raise AssertionError(f"{name}: oracle != reference at step {t}: max|d|={float((a.reshape(7) - actions[t]).abs().max()):.3e}")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"{name}.npz", actions=actions.numpy(), plan_idx=plan_idx.numpy())
    print(f"[golden] {name}: {T} steps, {len(plans)} plans, |a|max={float(actions.abs().max()):.3f}  ({time.time() - t0:.1f}s)")


def run_aug_case(name="aug_random_shifts"):
    """RandomShiftsAug of the UNMODIFIED reference (hulc/utils/transforms.py:8-29; drawn shifts injected through a patched torch.randint), followed
    by the yaml's ScaleImageTensor + Normalize(0.5, 0.5), on seeded uint8 frames of both cameras; pins oracle.random_shifts_aug (an exact crop of
    the replicate-padded frame; the reference's bilinear grid_sample lands within 6e-5 of it) and writes a sub-sampled fixture."""
    from hulc.utils.transforms import RandomShiftsAug

    fx = {}
    for cam, (h, pad, n) in {"static": (200, 10, 3), "gripper": (84, 4, 4)}.items():
        g = torch.Generator().manual_seed(17 + h)
        x = torch.randint(0, 256, (n, 3, h, h), generator=g, dtype=torch.uint8)
        sh = torch.randint(0, 2 * pad + 1, (n, 2), generator=g)
        orig = torch.randint
        torch.randint = lambda *a, **k: sh.view(n, 1, 1, 2).to(k.get("dtype", torch.float32))
        try:
            ref = RandomShiftsAug(pad)(x)
        finally:
            torch.randint = orig
        ref = ((ref / 255.0) - 0.5) / 0.5
        mine = O.random_shifts_aug(x, sh, pad)
        err = float((mine - ref).abs().max())
        assert err < 1e-4, f"{name}/{cam}: oracle != reference ({err:.2e})"
        fx[f"{cam}_shifts"], fx[f"{cam}_out_sub"] = sh.numpy().astype(np.int32), ref[:, :, ::7, ::5].numpy()
        print(f"[golden] {name}/{cam}: {n} frames {h}x{h}, pad {pad}: max |oracle - reference| = {err:.2e}")
    np.savez_compressed(GOLDEN / f"{name}.npz", **fx)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES) + list(VAL_CASES) + list(INFER_CASES) + list(AUTOCAST_CASES) + ["aug_random_shifts"]
    for n in names:
        if n == "aug_random_shifts":
            run_aug_case(n)
            continue
        (run_val_case if n in VAL_CASES else run_infer_case if n in INFER_CASES else run_autocast_case if n in AUTOCAST_CASES else run_case)(n)
