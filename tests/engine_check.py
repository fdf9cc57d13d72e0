"""Shared checker: run hulc_b200.engine and the oracle on the same seeded inputs and compare losses, action logits,
intermediates and every parameter gradient.  Used by the emulator tests (CPU) and the GPU parity tests."""
from __future__ import annotations

import numpy as np
import torch

from hulc_b200.utils import synthetic
from oracle import hulc_oracle as O


def _cat_masks(masks, mods):
    """per-modality bool keep-masks -> whole-batch uint8 masks in batch order"""
    return {k: torch.cat([masks[m][k] for m in mods], 0).to(torch.uint8).contiguous() for k in masks[mods[0]]}


def run_pair(model, rnn_model, B, S, p, device, hw=(200, 84), use_idx=False, seed=1, max_window=32, precision="fp32", aux=False, aux_mask=None):
    """aux: the BC-Z and MIA auxiliary heads next to the CLIP loss (ablation configs, hulc.py:567-648); aux_mask: use_for_aux_lang_loss of the
    language modality (None keeps the batch's all-true mask)."""
    from hulc_b200.engine import HulcEngine
    from hulc_b200.spec import ModelDims

    dims = ModelDims.shipped(model, rnn_model, max_window, bc_z=True, mia=True, **({} if model == "mcil" else {"dropout_p": float(p)})) if aux else None
    sd = synthetic.make_state_dict(model, rnn_model, max_window=max_window, dims=dims)
    if hw != (200, 84):  # reduced frames (emulator speed): the gripper flatten-FC shrinks with them
        k = ((((hw[1] - 8) // 4 + 1) - 4) // 2 + 1) - 2
        sd["perceptual_encoder.rgb_gripper_encoder.conv_model.7.weight"] = sd["perceptual_encoder.rgb_gripper_encoder.conv_model.7.weight"][:, : 64 * k * k].contiguous()
    batch = synthetic.make_batch(B, S, seed=seed, static_hw=hw[0], gripper_hw=hw[1])
    if aux_mask is not None:
        batch["lang"]["use_for_aux_lang_loss"] = aux_mask
    mods = list(batch)
    noise = {m: synthetic.plan_noise(B, S, m) for m in mods}
    masks = {m: synthetic.dropout_masks(B, S, m, p) for m in mods} if p > 0 else None

    # oracle (CPU autograd)
    sd_o = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.training_step(sd_o, batch, model=model, rnn_model=rnn_model, dropout_p=p, plan_u={m: noise[m]["u"] for m in mods},
                          plan_eps={m: noise[m]["eps"] for m in mods}, dropout_masks=masks, bc_z_beta=1.0 if aux else None, mia_beta=1.0 if aux else None)
    ref["total_loss"].backward()

    eng = HulcEngine(model, rnn_model, max_window=max_window, device=device, dropout_p=p, precision=precision, dims=dims)
    if hw != (200, 84):
        from hulc_b200.engine import ParamStore
        eng.spec["perceptual_encoder.rgb_gripper_encoder.conv_model.7.weight"] = tuple(sd["perceptual_encoder.rgb_gripper_encoder.conv_model.7.weight"].shape)
        eng.ps = ParamStore(eng.spec, device)
    eng.load_state_dict(sd)
    dbatch = synthetic._to(batch, device)
    kw = {}
    if model != "gcbc":
        if model == "mcil":
            kw["plan_eps"] = {m: noise[m]["eps"].to(device) for m in mods}
        elif use_idx:
            kw["plan_idx"] = {m: ref[f"plan_idx_{m}"].to(device) for m in mods}
        else:
            kw["plan_u"] = {m: noise[m]["u"].to(device) for m in mods}
    if masks is not None:
        kw["dropout_masks"] = {k: v.to(device) for k, v in _cat_masks(masks, mods).items()}
    out = eng.step(dbatch, **kw)
    return dict(ref=ref, out=out, eng=eng, sd_o=sd_o, mods=mods, B=B, S=S, model=model)


def compare(res, rtol=1e-3, atol=1e-4, grad_rtol=2e-3, inter_rtol=None, inter_atol=None, skip_grads=()):
    """rtol/atol: losses and action logits (the north-star tolerance).  inter_*: intermediates (embeddings, latent
    states); default the same.  grad_rtol: relative L2 error of every parameter gradient."""
    ref, out, eng, sd_o, mods, B, S, model = (res[k] for k in ("ref", "out", "eng", "sd_o", "mods", "B", "S", "model"))
    cpu = lambda t: t.detach().float().cpu()
    report = {}
    for k in ("total_loss", "action_loss", "kl_loss", "lang_clip_loss", "lang_pred_loss", "lang_contrastive_loss"):
        if k in ref or k in out:
            assert k in ref and k in out, f"{k}: reported by only one side"
            a, b = float(cpu(out[k])), float(ref[k])
            report[k] = (a, b)
            np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=k)
    # intermediates
    irt, iat = (inter_rtol if inter_rtol is not None else rtol), (inter_atol if inter_atol is not None else atol)
    emb_ref = torch.cat([ref[f"emb_{m}"] for m in mods], 0)
    torch.testing.assert_close(cpu(out["perceptual_emb"]), emb_ref.detach(), rtol=irt, atol=iat)
    torch.testing.assert_close(cpu(out["latent_goal"]), torch.cat([ref[f"goal_{m}"] for m in mods], 0).detach(), rtol=irt, atol=iat)
    torch.testing.assert_close(cpu(out["pr_state"]), torch.cat([ref[f"pr_state_{m}"] for m in mods], 0).detach(), rtol=irt, atol=iat)
    if "pp_state" in out:
        torch.testing.assert_close(cpu(out["pp_state"]), torch.cat([ref[f"pp_state_{m}"] for m in mods], 0).detach(), rtol=irt, atol=iat)
    if "plan_idx" in out:
        assert torch.equal(cpu(out["plan_idx"]).long(), torch.cat([ref[f"plan_idx_{m}"] for m in mods], 0))
    if f"actions_tcp_{mods[0]}" in ref:
        torch.testing.assert_close(cpu(out["actions_tcp"]), torch.cat([ref[f"actions_tcp_{m}"] for m in mods], 0), rtol=1e-4, atol=2e-4)
    # action logits: engine heads are time-major rows [S, nB, logit_probs | means | log_scales | gripper]
    heads = cpu(out["heads_tm"]).transpose(0, 1)  # (nB, S, n)
    nm = eng.n_dims * eng.n_mix
    lp = torch.cat([ref[f"logit_probs_{m}"] for m in mods], 0).detach().reshape(len(mods) * B, S, nm)
    mu = torch.cat([ref[f"means_{m}"] for m in mods], 0).detach().reshape(len(mods) * B, S, nm)
    ls = torch.cat([ref[f"log_scales_{m}"] for m in mods], 0).detach().reshape(len(mods) * B, S, nm)
    torch.testing.assert_close(heads[..., :nm], lp, rtol=rtol, atol=atol)
    torch.testing.assert_close(heads[..., nm : 2 * nm], mu, rtol=rtol, atol=atol)
    torch.testing.assert_close(heads[..., 2 * nm : 3 * nm].clamp(min=-7.0), ls, rtol=rtol, atol=atol)
    if f"gripper_act_{mods[0]}" in ref:
        torch.testing.assert_close(heads[..., 3 * nm :], torch.cat([ref[f"gripper_act_{m}"] for m in mods], 0).detach(), rtol=rtol, atol=atol)
    # every parameter gradient, relative to its norm
    worst = ("", 0.0)
    for k, v in sd_o.items():
        g = cpu(eng.ps.g[k])
        if k in skip_grads:
            continue
        if v.grad is None:
            assert float(g.abs().max()) == 0.0, f"{k}: reference has no gradient"
            continue
        gn = float(v.grad.norm())
        err = float((g - v.grad).norm()) / max(gn, 1e-12)
        if err > worst[1]:
            worst = (k, err)
        assert err < grad_rtol or float((g - v.grad).abs().max()) < 1e-7, f"grad {k}: rel err {err:.3e} (norm {gn:.3e})"
    report["worst_grad"] = worst
    return report
