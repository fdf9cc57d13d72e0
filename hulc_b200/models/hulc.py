"""`Hulc` — drop-in for the training surface of the reference LightningModule `hulc.models.hulc.Hulc`
(reference hulc/models/hulc.py:27-695): same constructor signature (:58-84), `training_step` (:390), `lmp_train` (:254),
`compute_kl_loss` (:539), `clip_auxiliary_loss` (:650), `configure_optimizers` (:239), `set_kl_beta` (:563), the same
`self.log` keys (:470-536) and the same `state_dict` keys/shapes, so `hulc/training.py` + the Hydra
configs drive it by re-pointing `model._target_`.

All arithmetic runs in the sm_100a kernels of libhulc_b200.so through `hulc_b200.engine.HulcEngine`; parameters are
`nn.Parameter` views into the engine's flat fp32 buffer (so Lightning, DDP, checkpoints and optimizers see ordinary
parameters) and gradients are produced by the engine's hand-written backward pass, not by autograd.  There is no CPU
path: calling `training_step` with the model on the CPU raises.
"""
from __future__ import annotations

import logging
import math
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from ..engine import HulcEngine

try:  # Lightning is optional: the reference's trainer needs it, the kernels do not
    import pytorch_lightning as pl

    _Base = pl.LightningModule
except Exception:  # pragma: no cover - exercised in this repo's environment (no lightning installed)
    pl = None

    class _Base(nn.Module):
        """Minimal stand-in for LightningModule when pytorch_lightning is not installed."""

        def __init__(self):
            super().__init__()
            self.logged: Dict[str, Any] = {}

        def log(self, name, value, **kw):
            self.logged[name] = value.detach() if torch.is_tensor(value) else value

        def save_hyperparameters(self, *a, **k):
            pass

        @property
        def device(self):
            return next(self.parameters()).device


logger = logging.getLogger(__name__)


def _get(cfg, key, default=None):
    if cfg is None:
        return default
    try:
        return cfg[key] if key in cfg else default
    except TypeError:
        return getattr(cfg, key, default)


def _tensors(d):
    for v in d.values():
        if isinstance(v, dict):
            yield from _tensors(v)
        elif torch.is_tensor(v):
            yield v


class _Block(nn.Module):
    """Name-space node so that `state_dict()` reproduces the reference's dotted keys."""


class _StepLoss(torch.autograd.Function):
    """Connects the engine's fused forward+backward to autograd: forward returns the loss the kernels computed,
    backward hands out the parameter gradients the kernels already produced (scaled by the incoming gradient)."""

    @staticmethod
    def forward(ctx, loss, engine, *params):
        ctx.engine = engine
        return loss.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        eng = ctx.engine
        grads = tuple(eng.ps.g[k] * grad_out for k in eng.ps.keys)
        return (None, None) + grads


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (hulc.py:240 instantiates `torch.optim.Adam(lr=2e-4)`) as ONE kernel launch over the
    engine's flat parameter / gradient / moment buffers.  Uses the gradients the engine wrote (or, when autograd / DDP
    populated `.grad` on the parameters, those — they are views of / copies into the same layout)."""

    def __init__(self, module: "Hulc", lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if weight_decay:
            raise NotImplementedError("the reference trains with weight_decay=0 (conf/model/optimizer/adam.yaml)")
        self._module = module
        super().__init__(list(module.parameters()), dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        mod, eng = self._module, self._module.engine
        g = self.param_groups[0]
        # gradients that reached the parameters through autograd/DDP (already averaged across ranks) win over the raw
        # local ones in the flat buffer
        for k, p in mod._param_by_key.items():
            if p.grad is not None and p.grad.data_ptr() != eng.ps.g[k].data_ptr():
                eng.ps.g[k].copy_(p.grad)
        eng.ps.adam_step(lr=g["lr"], betas=g["betas"], eps=g["eps"])
        return loss

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)


class Hulc(_Base):
    MODEL = "hulc"

    def __init__(
        self,
        perceptual_encoder,
        plan_proposal,
        plan_recognition,
        language_goal,
        visual_goal,
        action_decoder,
        kl_beta: float,
        kl_balancing_mix: float,
        state_recons: bool,
        state_recon_beta: float,
        use_bc_z_auxiliary_loss: bool,
        bc_z_auxiliary_loss_beta: float,
        use_mia_auxiliary_loss: bool,
        mia_auxiliary_loss_beta: float,
        optimizer,
        lr_scheduler,
        distribution,
        val_instructions,
        use_clip_auxiliary_loss: bool,
        clip_auxiliary_loss_beta: float,
        replan_freq: int = 30,
        bc_z_lang_decoder=None,
        mia_lang_discriminator=None,
        proj_vis_lang=None,
        device: Optional[str] = None,
        precision: str = "tf32",
    ):
        super().__init__()
        if state_recons or use_bc_z_auxiliary_loss or use_mia_auxiliary_loss:
            raise NotImplementedError("state reconstruction / BC-Z / MIA auxiliary losses are off in every shipped model yaml and are not built (DESIGN.md, out of scope)")
        for name, enc in (("depth_static", None), ("depth_gripper", None), ("proprio", None), ("tactile", None)):
            if _get(perceptual_encoder, name) not in (None, {}, "none"):
                raise NotImplementedError(f"perceptual_encoder.{name} is disabled in conf/model/perceptual_encoder/gripper_cam.yaml and not built")
        continuous = _get(distribution, "dist", "discrete") == "continuous"
        birnn = str(_get(plan_recognition, "_target_", "")).endswith("PlanRecognitionBiRNNNetwork")
        if continuous != birnn:
            raise NotImplementedError("supported latent plans: transformer posterior + discrete latent (hulc/gcbc) or BiRNN posterior + continuous latent (mcil)")
        model = "mcil" if birnn else self.MODEL
        if model != "mcil" and (_get(distribution, "category_size", 32), _get(distribution, "class_size", 32)) != (32, 32):
            raise NotImplementedError("the plan kernels are specialised for 32 categoricals x 32 classes (conf/model/distribution/discrete.yaml)")
        if bool(use_clip_auxiliary_loss) != (model != "mcil"):
            raise NotImplementedError("CLIP auxiliary loss is on for hulc/gcbc and off for mcil in the shipped configs")
        rnn_model = _get(action_decoder, "rnn_model", "rnn_decoder")
        max_window = int(_get(plan_recognition, "max_position_embeddings", 32))
        dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.engine = HulcEngine(
            model, rnn_model, max_window=max_window, device=dev, dropout_p=float(_get(plan_recognition, "dropout_p", 0.0) or 0.0),
            kl_beta=kl_beta, kl_balancing_mix=kl_balancing_mix, clip_beta=clip_auxiliary_loss_beta,
            gripper_alpha=float(_get(action_decoder, "gripper_alpha", 1.0)), nhead=int(_get(plan_recognition, "num_heads", 8)),
            nlayers=int(_get(plan_recognition, "num_layers", 2)), lr=float(_get(optimizer, "lr", 2e-4)), precision=precision,
        )
        self._param_by_key: Dict[str, nn.Parameter] = {}
        for k in self.engine.ps.keys:
            self._register(k, nn.Parameter(self.engine.ps.p[k]))
        self._register_reference_buffers(model)
        self._init_parameters()
        self.use_clip_auxiliary_loss, self.clip_auxiliary_loss_beta = use_clip_auxiliary_loss, clip_auxiliary_loss_beta
        self.kl_beta, self.kl_balancing_mix = kl_beta, kl_balancing_mix
        self.modality_scope = "vis"
        self.optimizer_config, self.lr_scheduler = optimizer, lr_scheduler
        self.replan_freq = replan_freq
        self.val_instructions = val_instructions
        self._graphs = None
        self.save_hyperparameters()

    # ---- parameter plumbing ---------------------------------------------------------------------------------------------
    def _register(self, key: str, param: nn.Parameter):
        node = self
        parts = key.split(".")
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Block())
            node = node._modules[p]
        node.register_parameter(parts[-1], param)
        self._param_by_key[key] = param

    def _buffer(self, key: str, value: torch.Tensor):
        node = self
        parts = key.split(".")
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Block())
            node = node._modules[p]
        node.register_buffer(parts[-1], value.to(self.engine.device))

    def _register_reference_buffers(self, model):
        """Buffers the reference modules carry in their state_dict (vision_network.py:88-98, logistic_decoder_rnn.py:50-80)."""
        lin = torch.linspace(-1.0, 1.0, 21)
        pe = "perceptual_encoder.rgb_static_encoder.spatial_softmax"
        self._buffer(f"{pe}.x_map", lin.view(21, 1).expand(21, 21).reshape(-1).clone())
        self._buffer(f"{pe}.y_map", lin.view(1, 21).expand(21, 21).reshape(-1).clone())
        self._buffer(f"{pe}.temperature", torch.ones(1))
        n_out, n_mix = self.engine.n_dims, self.engine.n_mix
        self._buffer("action_decoder.one_hot_embedding_eye", torch.eye(n_mix))
        self._buffer("action_decoder.ones", torch.ones(1, 1, n_mix))
        self._buffer("action_decoder.action_max_bound", torch.ones(1, 1, n_out, n_mix))
        self._buffer("action_decoder.action_min_bound", -torch.ones(1, 1, n_out, n_mix))
        if model != "mcil":
            self._buffer("action_decoder.gripper_bounds", torch.tensor([-1.0, 1.0]))

    @torch.no_grad()
    def _init_parameters(self):
        """PyTorch's default initialisers for the reference's module types (Linear/Conv: kaiming-uniform(a=sqrt 5) ==
        U(+-1/sqrt(fan_in)) for weight and bias; RNN/GRU: U(+-1/sqrt(hidden)); LayerNorm: 1/0; Embedding: N(0,1);
        logit_scale = ln(1/0.07), hulc.py:115)."""
        for k, p in self._param_by_key.items():
            if k == "logit_scale":
                p.fill_(math.log(1 / 0.07))
            elif ".ln." in k or ".norm1." in k or ".norm2." in k:
                p.fill_(1.0 if k.endswith("weight") else 0.0)
            elif "position_embeddings" in k:
                p.normal_()
            elif ".rnn." in k or "birnn_model" in k:
                p.uniform_(-1.0 / math.sqrt(self.engine.H), 1.0 / math.sqrt(self.engine.H))
            else:
                wkey = k[: -len("bias")] + "weight" if k.endswith("bias") else k
                wkey = wkey.replace("in_proj_weight", "in_proj_weight")
                w = self._param_by_key.get(wkey, p)
                fan_in = int(math.prod(w.shape[1:])) if w.dim() > 1 else w.shape[0]
                if k.endswith("in_proj_bias") or k.endswith("out_proj.bias"):
                    p.zero_()
                else:
                    p.uniform_(-1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))

    def _apply(self, fn, *a, **kw):
        """`.to(device)` / `.cuda()`: move the flat buffers and re-point every parameter at its view."""
        ps = self.engine.ps
        new_flat = fn(ps.flat)
        if new_flat.device != ps.flat.device or new_flat.dtype != ps.flat.dtype:
            if new_flat.dtype != torch.float32:
                raise NotImplementedError("parameters are kept in fp32 (the kernels compute in fp32)")
            ps.flat, ps.grad, ps.exp_avg, ps.exp_avg_sq = new_flat, fn(ps.grad), fn(ps.exp_avg), fn(ps.exp_avg_sq)
            ps.device = new_flat.device
            ps.rebuild_views()
            eng = self.engine
            eng.device = new_flat.device
            eng._bufs.clear()
            eng.nan_flag = fn(eng.nan_flag)
            for k, p in self._param_by_key.items():
                p.data = ps.p[k]
                p.grad = None
            for mod in self.modules():
                for name, b in list(mod._buffers.items()):
                    if b is not None:
                        mod._buffers[name] = fn(b)
        return self

    # ---- reference surface ----------------------------------------------------------------------------------------------------
    def configure_optimizers(self):
        """hulc.py:239-252: Adam + constant schedule stepped every iteration."""
        opt = FusedAdam(self, lr=float(_get(self.optimizer_config, "lr", 2e-4)))
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda _: 1.0)  # transformers.get_constant_schedule
        return {"optimizer": opt, "lr_scheduler": {"scheduler": sched, "interval": "step", "frequency": 1}}

    def set_kl_beta(self, kl_beta):
        """Called by the KL-annealing callbacks (hulc/utils/kl_callbacks.py:22)."""
        self.kl_beta = kl_beta
        self.engine.kl_beta = float(kl_beta)

    def enable_cuda_graphs(self, flag: bool = True):
        """Replay the training step from a CUDA graph captured per distinct set of input buffers (the batch tensors are the
        graph's static inputs: a loader that re-uses its device staging buffers hits the same graph every step)."""
        self._graphs = {} if flag else None

    def fused_step(self, batch, **inject) -> Dict[str, torch.Tensor]:
        """Forward + backward in the kernels; gradients land in the flat gradient buffer.  No autograd graph."""
        if self._graphs is not None and not inject:
            key = tuple((t.data_ptr(), tuple(t.shape)) for t in _tensors(batch))
            sg = self._graphs.get(key)
            if sg is None:
                sg = self._graphs[key] = self.engine.capture(batch, optimizer=False)
            return sg.replay()
        return self.engine.step(batch, **inject)

    def training_step(self, batch: Dict[str, Dict], batch_idx: int = 0, **inject) -> torch.Tensor:
        """hulc.py:390-537.  Returns total_loss with an autograd edge to every parameter, so Lightning's
        `loss.backward()` (and DDP's gradient hooks) work unchanged; the gradients themselves were computed by the
        kernels' backward pass during this call."""
        out = self.fused_step(batch, **inject)
        self.last_outputs = out
        mods = list(batch.keys())
        kl, act = out["kl_loss"], out["action_loss"]
        for m in mods:
            # key names as logged by the reference (hulc.py:470-490); kl is logged once unscaled-by-beta there ("kl_loss")
            # and once scaled per modality
            self.log(f"train/kl_loss_scaled_{m}", out[f"kl_loss_{m}"], on_step=False, on_epoch=True)
            self.log(f"train/action_loss_{m}", out[f"action_loss_{m}"], on_step=False, on_epoch=True)
        self.log("train/kl_loss", kl, on_step=False, on_epoch=True, sync_dist=True)
        self.log("train/action_loss", act, on_step=False, on_epoch=True, sync_dist=True)
        if "lang_clip_loss" in out:
            self.log("train/lang_clip_loss", self.clip_auxiliary_loss_beta * out["lang_clip_loss"], on_step=False, on_epoch=True, sync_dist=True)
        self.log("train/total_loss", out["total_loss"], on_step=False, on_epoch=True, sync_dist=True)
        params = [self._param_by_key[k] for k in self.engine.ps.keys]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _StepLoss.apply(out["total_loss"], self.engine, *params)
        return out["total_loss"]

    @torch.no_grad()
    def validation_step(self, batch: Dict[str, Dict], batch_idx: int = 0, **inject) -> Dict[str, torch.Tensor]:
        """hulc.py:739-841 (+ lmp_val, :301-388): logs the reference's validation keys and returns the sampled plans and episode indices.
        Dropout is off as under `model.eval()`.  `clip_groundtruth` (a logging-only metric over the dataset's task annotations) is not
        computed.  `inject`: see HulcEngine.validation_step."""
        out = self.engine.validation_step(batch, **inject)
        self.last_val_outputs = out
        mods = list(batch.keys())
        output: Dict[str, torch.Tensor] = {}
        total_pp = None
        gcbc = self.engine.model == "gcbc"
        for m in mods:
            if gcbc:  # gcbc.py:183-281: one decoder pass, no latent plan
                if "lang" in m and "val_pred_clip_loss" in out:
                    self.log("val/val_pred_clip_loss", out["val_pred_clip_loss"], sync_dist=True)
                total_pp = out[f"action_loss_{m}"] if total_pp is None else total_pp + out[f"action_loss_{m}"]
                self.log("val_act/action_loss", total_pp / len(mods), sync_dist=True)
                self.log(f"val_act/{m}_act_loss", out[f"action_loss_{m}"], sync_dist=True)
                self.log(f"val_total_mae/{m}_total_mae", out[f"mae_{m}"].mean(), sync_dist=True)
                self.log(f"val_pos_mae/{m}_pos_mae", out[f"mae_{m}"][..., :3].mean(), sync_dist=True)
                self.log(f"val_orn_mae/{m}_orn_mae", out[f"mae_{m}"][..., 3:6].mean(), sync_dist=True)
                self.log(f"val_grip/{m}_grip_sr", out[f"gripper_sr_{m}"], sync_dist=True)
                output[f"idx_{m}"] = batch[m]["idx"]
                continue
            if "lang" in m and "val_pred_clip_loss" in out:
                self.log("val/val_pred_clip_loss", out["val_pred_clip_loss"], sync_dist=True)
            total_pp = out[f"action_loss_pp_{m}"] if total_pp is None else total_pp + out[f"action_loss_pp_{m}"]
            for w in ("pr", "pp"):
                mae = out[f"mae_{w}_{m}"]
                self.log(f"val_total_mae/{m}_total_mae_{w}", mae.mean(), sync_dist=True)
                self.log(f"val_pos_mae/{m}_pos_mae_{w}", mae[..., :3].mean(), sync_dist=True)
                self.log(f"val_orn_mae/{m}_orn_mae_{w}", mae[..., 3:6].mean(), sync_dist=True)
                self.log(f"val_act/{m}_act_loss_{w}", out[f"action_loss_{w}_{m}"], sync_dist=True)
                self.log(f"val_grip/{m}_grip_sr_{w}", out[f"gripper_sr_{w}_{m}"], sync_dist=True)
                output[f"sampled_plan_{w}_{m}"] = out[f"sampled_plan_{w}_{m}"]
            self.log(f"val_kl/{m}_kl_loss", out[f"kl_loss_{m}"], sync_dist=True)
            # the reference logs the running sum over the modalities seen so far, divided by their total number (hulc.py:830-834)
            self.log("val_act/action_loss_pp", total_pp / len(mods), sync_dist=True)
            output[f"idx_{m}"] = batch[m]["idx"]
        return output

    # ---- inference (hulc.py:843-957) ------------------------------------------------------------------------------------------
    def reset(self):
        """Call at the beginning of a rollout (hulc.py:843-849)."""
        self.plan = None
        self.latent_goal = None
        self.rollout_step_counter = 0
        self.engine._infer_state = None

    def load_lang_embeddings(self, embeddings_path):
        """hulc.py:872-882: <dataset>/validation/embeddings.npy -> {annotation: embedding}."""
        import numpy as np

        embeddings = np.load(embeddings_path, allow_pickle=True).item()
        self.lang_embeddings = {v["ann"][0]: v["emb"] for k, v in embeddings.items()}

    @torch.no_grad()
    def step(self, obs, goal, *, plan_idx=None, plan_u=None, sample_u=None):
        """One control step (hulc.py:851-870).  obs: {"rgb_obs": {"rgb_static": (1,1,3,H,W), "rgb_gripper": (1,1,3,h,w)}, "robot_obs_raw":
        (1,1,15), ...}; goal: an annotation string (key of `lang_embeddings`) or a goal observation dict.  Every `replan_freq` steps the goal
        is encoded and a plan sampled from the proposal network (the decoder's hidden state is cleared); then the decoder advances one step and
        the sampled action (world frame, (1,1,7)) is returned.  plan_idx / plan_u / sample_u inject the randomness for parity runs."""
        if not hasattr(self, "rollout_step_counter"):
            self.reset()
        dev = self.engine.device
        st, gr = obs["rgb_obs"]["rgb_static"].to(dev).float(), obs["rgb_obs"]["rgb_gripper"].to(dev).float()
        if self.rollout_step_counter % self.replan_freq == 0:
            if isinstance(goal, str):
                lang = torch.from_numpy(self.lang_embeddings[goal]).to(dev).squeeze(0).float()
                self.plan, self.latent_goal = self.engine.infer_plan(st[0].contiguous(), gr[0].contiguous(), lang=lang.reshape(1, -1).contiguous(), plan_idx=plan_idx, plan_u=plan_u)
            else:
                gs, gg = goal["rgb_obs"]["rgb_static"].to(dev).float(), goal["rgb_obs"]["rgb_gripper"].to(dev).float()
                self.plan, self.latent_goal = self.engine.infer_plan(torch.cat([st, gs], 1)[0].contiguous(), torch.cat([gr, gg], 1)[0].contiguous(), plan_idx=plan_idx,
                                                                     plan_u=plan_u)
        action = self.engine.infer_act(st[0].contiguous(), gr[0].contiguous(), obs["robot_obs_raw"].to(dev).float().reshape(1, -1), sample_u=sample_u)
        self.rollout_step_counter += 1
        return action

    def lmp_train(self, perceptual_emb, latent_goal, train_acts, robot_obs):
        raise NotImplementedError(
            "lmp_train is fused into training_step here (one batched pass over both modalities); the per-block values it "
            "returned in the reference are in `self.last_outputs` (pp_state, pr_state, seq_feat, kl/action losses)")

    def compute_kl_loss(self, pp_state, pr_state):
        """hulc.py:539-561 on device tensors of logits (discrete) / [mean|raw_std] (continuous): returns the scaled,
        balanced KL the training step uses (forward value only; its gradient is part of the fused backward)."""
        from .. import ops

        pp, pr = pp_state.contiguous().float(), pr_state.contiguous().float()
        Bn = pp.shape[0]
        out = torch.empty(1, device=pp.device)
        if self.engine.discrete:
            kl_rows = torch.empty(Bn * 32, device=pp.device)
            ops.plan_discrete_fwd(pr, pp, None, kl_rows, idx_in=torch.zeros(Bn * 32, dtype=torch.int32, device=pp.device))
            ops.sum_to(kl_rows, out, float(self.kl_beta) / Bn)
        else:
            P = pp.shape[1] // 2
            kl_el, plan = torch.empty(Bn, P, device=pp.device), torch.empty(Bn, P, device=pp.device)
            ops.plan_cont_fwd(pr, pp, plan, kl_el, eps=torch.zeros(Bn, P, device=pp.device))
            ops.sum_to(kl_el, out, float(self.kl_beta) / Bn)
        return out[0]

    def on_load_checkpoint(self, checkpoint):  # state_dict keys are the reference's; nothing to translate
        pass
