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
from ..spec import dims_from_configs

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
    """Connects the engine's fused forward+backward to autograd: forward returns the loss the kernels computed, backward hands out
    the parameter gradients the kernels already produced.  Zero-copy: the gradients handed to autograd are fresh VIEWS of the flat
    gradient buffer (so `p.grad` aliases it and `FusedAdam.step` has nothing to copy back); the incoming gradient (1 for a plain
    `loss.backward()`, 1/k under gradient accumulation, the loss scale under AMP) is applied to the flat buffer in place by one
    launch that reads the factor from the device and returns immediately when it is 1."""

    @staticmethod
    def forward(ctx, loss, engine, *params):
        ctx.engine = engine
        return loss.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        from .. import ops

        ps = ctx.engine.ps
        ops.scale_dev_(ps.grad, grad_out.reshape(1).to(torch.float32))
        return (None, None) + tuple(ps._view(ps.grad, k) for k in ps.keys)


_ADAM_DEFAULTS = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                      differentiable=False, fused=None)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (hulc.py:240 instantiates `torch.optim.Adam(lr=2e-4)`) as ONE kernel launch over the
    engine's flat parameter / gradient / moment buffers.  Gradients: the ones the engine wrote; when autograd / DDP put other
    tensors into `p.grad` (DDP's bucket views, already averaged across ranks) those are copied into the flat layout first —
    gradients that already alias the flat buffer (the normal `loss.backward()` path, see _StepLoss) cost nothing.
    `state_dict()` / `load_state_dict()` use torch.optim.Adam's own layout ({"state": {i: {"step", "exp_avg", "exp_avg_sq"}},
    "param_groups": [...]}, parameters numbered in `module.parameters()` order), so optimizer checkpoints interchange with the
    reference's and resuming restores the moments and the bias-correction step (hulc/training.py resumes via trainer.fit(ckpt_path))."""

    def __init__(self, module: "Hulc", lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if weight_decay:
            raise NotImplementedError("the reference trains with weight_decay=0 (conf/model/optimizer/adam.yaml)")
        self._module = module
        self._gptrs = None
        super().__init__(list(module.parameters()), dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        mod, eng = self._module, self._module.engine
        g = self.param_groups[0]
        ps = eng.ps
        if self._gptrs is None or self._gptrs[0] is not ps.grad:  # (re)build after the flat buffers moved
            self._gptrs = (ps.grad, [(k, p, ps.g[k].data_ptr()) for k, p in mod._param_by_key.items()])
        for k, p, ptr in self._gptrs[1]:
            pg = p.grad
            if pg is not None and pg.data_ptr() != ptr:
                ps.g[k].copy_(pg)
        ps.adam_step(lr=g["lr"], betas=g["betas"], eps=g["eps"])
        return loss

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)

    # ---- checkpointing: torch.optim.Adam's layout --------------------------------------------------------------------------
    def _keys_in_param_order(self):
        by_id = {id(p): k for k, p in self._module._param_by_key.items()}
        return [by_id[id(p)] for p in self.param_groups[0]["params"]]

    def state_dict(self):
        ps = self._module.engine.ps
        keys = self._keys_in_param_order()
        state = {}
        if ps.step_count > 0:
            for i, k in enumerate(keys):
                state[i] = {"step": torch.tensor(float(ps.step_count)), "exp_avg": ps._view(ps.exp_avg, k).detach().clone(),
                            "exp_avg_sq": ps._view(ps.exp_avg_sq, k).detach().clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        for k, v in _ADAM_DEFAULTS.items():
            group.setdefault(k, v)
        group["params"] = list(range(len(keys)))
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        ps = self._module.engine.ps
        keys = self._keys_in_param_order()
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(keys):
            raise ValueError("optimizer state does not match this model: one parameter group over all parameters expected")
        for name in ("lr", "betas", "eps"):
            if name in groups[0]:
                self.param_groups[0][name] = tuple(groups[0][name]) if name == "betas" else groups[0][name]
        if "initial_lr" in groups[0]:
            self.param_groups[0]["initial_lr"] = groups[0]["initial_lr"]
        if groups[0].get("weight_decay", 0) or groups[0].get("amsgrad", False):
            raise NotImplementedError("weight_decay / amsgrad optimizer state is not supported")
        state = state_dict["state"]
        steps = set()
        ps.exp_avg.zero_(); ps.exp_avg_sq.zero_()
        for i, pid in enumerate(groups[0]["params"]):
            st = state.get(pid, state.get(str(pid)))
            if st is None:
                continue
            k = keys[i]
            ps._view(ps.exp_avg, k).copy_(st["exp_avg"].to(ps.device, torch.float32))
            ps._view(ps.exp_avg_sq, k).copy_(st["exp_avg_sq"].to(ps.device, torch.float32))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise NotImplementedError(f"per-parameter step counts differ ({sorted(steps)}): the fused update keeps one step count")
        ps.step_count = steps.pop() if steps else 0
        ps.step_dev.fill_(ps.step_count)


def _lr_lambda(cfg, num_training_steps_fn):
    """The three schedules the reference ships (conf/model/lr_scheduler/*.yaml -> transformers.get_*_schedule*), as LambdaLR factors."""
    target = str(_get(cfg, "_target_", "transformers.get_constant_schedule"))
    name = target.rsplit(".", 1)[-1]
    if name == "get_constant_schedule":
        return lambda step: 1.0
    if name not in ("get_cosine_schedule_with_warmup", "get_linear_schedule_with_warmup"):
        raise NotImplementedError(f"lr_scheduler._target_={target!r}: constant, linear-with-warmup and cosine-with-warmup are supported")
    total, warm = _get(cfg, "num_training_steps", -1), _get(cfg, "num_warmup_steps", 0)
    if total is None or total < 0:  # hulc.py:218-237 (compute_warmup): infer from the trainer
        total = num_training_steps_fn()
    if isinstance(warm, float):
        warm = warm * total
    total, warm = int(total), int(warm)
    if name == "get_linear_schedule_with_warmup":
        return lambda step: step / max(1, warm) if step < warm else max(0.0, (total - step) / max(1, total - warm))
    cycles = float(_get(cfg, "num_cycles", 0.5))

    def cosine(step):
        if step < warm:
            return step / max(1, warm)
        progress = (step - warm) / max(1, total - warm)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * cycles * 2.0 * progress)))

    return cosine


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
        if state_recons:
            raise NotImplementedError("state_recons needs the proprio encoder that conf/model/perceptual_encoder/gripper_cam.yaml disables (concat_encoders.py:46-50); not built")
        if bool(use_bc_z_auxiliary_loss) != (bc_z_lang_decoder not in (None, {}, "none")) or bool(use_mia_auxiliary_loss) != (mia_lang_discriminator not in (None, {}, "none")):
            raise NotImplementedError("use_bc_z_auxiliary_loss / use_mia_auxiliary_loss go together with their networks (model/bc_z_lang_decoder, model/mia_lang_discriminator)")
        birnn = str(_get(plan_recognition, "_target_", "")).endswith("PlanRecognitionBiRNNNetwork")
        model = "mcil" if birnn else self.MODEL
        if bool(use_clip_auxiliary_loss) != (model != "mcil"):
            raise NotImplementedError("CLIP auxiliary loss is on for hulc/gcbc and off for mcil in the shipped configs")
        # every size of the network comes from the config tree (hulc.py:86-187 wires the sub-configs the same way); values the kernels
        # cannot honour raise NotImplementedError naming the key
        dims = dims_from_configs(model, perceptual_encoder, plan_proposal, plan_recognition, language_goal, visual_goal, action_decoder, distribution, proj_vis_lang,
                                 bc_z_lang_decoder=bc_z_lang_decoder, mia_lang_discriminator=mia_lang_discriminator)
        self._check_optimizer_config(optimizer)
        dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.engine = HulcEngine(device=dev, kl_beta=kl_beta, kl_balancing_mix=kl_balancing_mix, clip_beta=clip_auxiliary_loss_beta,
                                 lr=float(_get(optimizer, "lr", 2e-4)), precision=precision, dims=dims)
        self._param_by_key: Dict[str, nn.Parameter] = {}
        for k in self.engine.ps.keys:
            self._register(k, nn.Parameter(self.engine.ps.p[k]))
        self._register_reference_buffers(model)
        self._init_parameters()
        self.use_clip_auxiliary_loss, self.clip_auxiliary_loss_beta = use_clip_auxiliary_loss, clip_auxiliary_loss_beta
        self.use_bc_z_auxiliary_loss, self.bc_z_auxiliary_loss_beta = bool(use_bc_z_auxiliary_loss), bc_z_auxiliary_loss_beta
        self.use_mia_auxiliary_loss, self.mia_auxiliary_loss_beta = bool(use_mia_auxiliary_loss), mia_auxiliary_loss_beta
        self.engine.bc_z_beta, self.engine.mia_beta = float(bc_z_auxiliary_loss_beta), float(mia_auxiliary_loss_beta)
        self.kl_beta, self.kl_balancing_mix = kl_beta, kl_balancing_mix
        self.modality_scope = "vis"
        self.optimizer_config, self.lr_scheduler = optimizer, lr_scheduler
        self.replan_freq = replan_freq
        self.val_instructions = val_instructions
        self._graphs = None
        self._graph_hyper = None
        self.max_graphs = 4  # captured steps kept (least recently used goes first)
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
        U(+-1/sqrt(fan_in)) for weight and bias; RNN/GRU: U(+-1/sqrt(hidden)); LayerNorm: 1/0; Embedding: N(0,1); the attention
        in-projection: xavier-uniform weight, zero bias, zero out-projection bias (nn.MultiheadAttention._reset_parameters);
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
                w = self._param_by_key.get(wkey, p)
                fan_in = int(math.prod(w.shape[1:])) if w.dim() > 1 else w.shape[0]
                if k.endswith("in_proj_bias") or k.endswith("out_proj.bias"):
                    p.zero_()
                elif k.endswith("in_proj_weight"):  # nn.MultiheadAttention._reset_parameters: xavier_uniform_
                    bound = math.sqrt(6.0 / (p.shape[0] + p.shape[1]))
                    p.uniform_(-bound, bound)
                else:
                    p.uniform_(-1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))

    def _apply(self, fn, *a, **kw):
        """`.to(device)` / `.cuda()`: move the flat buffers AND every other piece of device state the kernels dereference (the Adam
        step counter, the RNG seed, the NaN flag), re-point every parameter at its view, and forget everything that was bound to the old
        device (activation buffers, captured graphs, a rollout's carried state).  Lightning builds the module before it picks the device
        (hulc/training.py:44 then trainer.fit), so every DDP rank goes through here."""
        ps = self.engine.ps
        new_flat = fn(ps.flat)
        if new_flat.device != ps.flat.device or new_flat.dtype != ps.flat.dtype:
            if new_flat.dtype != torch.float32:
                raise NotImplementedError("parameters are kept in fp32 (the kernels compute in fp32)")
            ps.flat, ps.grad, ps.exp_avg, ps.exp_avg_sq = new_flat, fn(ps.grad), fn(ps.exp_avg), fn(ps.exp_avg_sq)
            ps.step_dev = fn(ps.step_dev)
            ps.device = new_flat.device
            ps.rebuild_views()
            eng = self.engine
            eng.device = new_flat.device
            eng.rng_dev = fn(eng.rng_dev)
            eng.nan_flag = fn(eng.nan_flag)
            eng._bufs.clear()
            eng._buf_namespaces.clear()
            eng._twins.clear()
            if eng.bf16:
                ps.flat_bf16 = None
                ps.enable_bf16()
            eng._infer_state, eng._infer_planned = None, False
            if eng._infer_graph is not None:
                eng._infer_graph = {}
            if self._graphs is not None:
                self._graphs = {}
            for k, p in self._param_by_key.items():
                p.data = ps.p[k]
                p.grad = None
            for mod in self.modules():
                for name, b in list(mod._buffers.items()):
                    if b is not None:
                        mod._buffers[name] = fn(b)
        return self

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """nn.Module.load_state_dict copies into the parameter views directly: the bf16 operand copy of the parameters must follow."""
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self.engine.ps.bf16_stale = True
        return res

    # ---- reference surface ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _check_optimizer_config(optimizer):
        target = str(_get(optimizer, "_target_", "torch.optim.Adam"))
        if target != "torch.optim.Adam":
            raise NotImplementedError(f"optimizer._target_={target!r}: the fused update implements torch.optim.Adam (conf/model/optimizer/adam.yaml)")
        for k in (optimizer.keys() if hasattr(optimizer, "keys") else []):
            if k in ("_target_", "_recursive_", "_partial_", "lr", "betas", "eps"):
                continue
            if k in _ADAM_DEFAULTS and (_get(optimizer, k) == _ADAM_DEFAULTS[k] or not _get(optimizer, k)):
                continue
            raise NotImplementedError(f"optimizer.{k}={_get(optimizer, k)!r} is not supported by the fused Adam update")

    @property
    def num_training_steps(self) -> int:
        """Total optimizer steps, inferred from the trainer like the reference (hulc.py:189-216)."""
        tr = getattr(self, "trainer", None)
        if tr is None:
            raise RuntimeError("lr_scheduler.num_training_steps=-1 needs an attached Lightning trainer to infer the number of steps")
        if getattr(tr, "max_steps", None) and tr.max_steps > 0:
            est = getattr(tr, "estimated_stepping_batches", None)
            return int(min(tr.max_steps, est)) if est else int(tr.max_steps)
        return int(tr.estimated_stepping_batches)

    def configure_optimizers(self):
        """hulc.py:239-252: the configured Adam + the configured schedule (constant by default), stepped every iteration."""
        cfg = self.optimizer_config
        betas = tuple(_get(cfg, "betas", (0.9, 0.999)))
        opt = FusedAdam(self, lr=float(_get(cfg, "lr", 2e-4)), betas=(float(betas[0]), float(betas[1])), eps=float(_get(cfg, "eps", 1e-8)))
        sched = torch.optim.lr_scheduler.LambdaLR(opt, _lr_lambda(self.lr_scheduler, lambda: self.num_training_steps))
        return {"optimizer": opt, "lr_scheduler": {"scheduler": sched, "interval": "step", "frequency": 1}}

    def set_kl_beta(self, kl_beta):
        """Called by the KL-annealing callbacks once per epoch (hulc/utils/kl_callbacks.py:19-22).  The coefficient is a by-value kernel
        argument, i.e. baked into captured graphs: graphs captured under another value are dropped and re-captured on their next use."""
        self.kl_beta = kl_beta
        self.engine.kl_beta = float(kl_beta)

    def enable_cuda_graphs(self, flag: bool = True):
        """Replay the training step from a CUDA graph captured per distinct set of input buffers (the batch tensors are the
        graph's static inputs: a loader that re-uses its device staging buffers hits the same graph every step).  At most
        `max_graphs` graphs are kept; every batch shape has its own activation buffers (HulcEngine.step).  With the BC-Z / MIA heads on the
        step stays eager: they inspect `use_for_aux_lang_loss` on the host every step, like the reference (hulc.py:581, 624)."""
        if flag and (self.engine.dims.bc_z or self.engine.dims.mia):
            flag = False
        self._graphs = {} if flag else None

    def fused_step(self, batch, **inject) -> Dict[str, torch.Tensor]:
        """Forward + backward in the kernels; gradients land in the flat gradient buffer.  No autograd graph."""
        if self._graphs is not None and not inject:
            eng = self.engine
            hyper = (eng.kl_beta, eng.kl_alpha, eng.clip_beta, eng.dropout_p)
            if hyper != self._graph_hyper:  # by-value kernel arguments changed (KL schedule): captured graphs are stale
                self._graphs.clear()
                self._graph_hyper = hyper
            key = tuple((t.data_ptr(), tuple(t.shape), str(t.dtype)) for t in _tensors(batch))
            sg = self._graphs.pop(key, None)
            if sg is None:
                while len(self._graphs) >= self.max_graphs:  # least recently used first
                    old_key = next(iter(self._graphs))
                    old = self._graphs.pop(old_key)
                    if not any(g.namespace == old.namespace for g in self._graphs.values()):
                        eng.release_buffers(old.namespace)
                sg = eng.capture(batch, optimizer=False)
            self._graphs[key] = sg  # most recently used last
            return sg.replay()
        return self.engine.step(batch, **inject)

    def _unalias_accumulated_grads(self):
        """Gradient accumulation (`accumulate_grad_batches > 1`): `p.grad` of the previous micro-batch aliases the flat gradient buffer this
        step is about to overwrite — move the accumulated values into tensors of their own first (autograd then adds this step's views)."""
        ps = self.engine.ps
        first = self._param_by_key[ps.keys[0]]
        if first.grad is None or first.grad.data_ptr() != ps.g[ps.keys[0]].data_ptr():
            return
        for k, p in self._param_by_key.items():
            if p.grad is not None and p.grad.data_ptr() == ps.g[k].data_ptr():
                p.grad = p.grad.clone()

    def training_step(self, batch: Dict[str, Dict], batch_idx: int = 0, **inject) -> torch.Tensor:
        """hulc.py:390-537.  Returns total_loss with an autograd edge to every parameter, so Lightning's
        `loss.backward()` (and DDP's gradient hooks) work unchanged; the gradients themselves were computed by the
        kernels' backward pass during this call."""
        self._unalias_accumulated_grads()
        out = self.fused_step(batch, **inject)
        self.last_outputs = out
        mods = list(batch.keys())
        kl, act = out["kl_loss"], out["action_loss"]
        total_bs = 0
        for m in mods:
            # key names and batch sizes as logged by the reference (hulc.py:470-490)
            bs = batch[m]["actions"].shape[0]
            total_bs += bs
            self.log(f"train/kl_loss_scaled_{m}", out[f"kl_loss_{m}"], on_step=False, on_epoch=True, batch_size=bs)
            self.log(f"train/action_loss_{m}", out[f"action_loss_{m}"], on_step=False, on_epoch=True, batch_size=bs)
            self.log(f"train/total_loss_{m}", out[f"action_loss_{m}"] + out[f"kl_loss_{m}"], on_step=False, on_epoch=True, batch_size=bs)
        if "lang_pred_loss" in out:  # hulc.py:500-509
            self.log("train/pred_lang", self.bc_z_auxiliary_loss_beta * out["lang_pred_loss"], on_step=False, on_epoch=True, sync_dist=True)
        if "lang_contrastive_loss" in out:  # hulc.py:510-519
            self.log("train/lang_contrastive", self.mia_auxiliary_loss_beta * out["lang_contrastive_loss"], on_step=False, on_epoch=True, sync_dist=True)
        if "lang_clip_loss" in out:
            self.log("train/lang_clip_loss", self.clip_auxiliary_loss_beta * out["lang_clip_loss"], on_step=False, on_epoch=True, sync_dist=True)
        self.log("train/kl_loss", kl, on_step=False, on_epoch=True, batch_size=total_bs)
        self.log("train/action_loss", act, on_step=False, on_epoch=True, batch_size=total_bs)
        self.log("train/total_loss", out["total_loss"], on_step=False, on_epoch=True, batch_size=total_bs)
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
        self.engine.infer_reset()

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
        if self.engine._infer_graph is not None and sample_u is None:
            # graph replay: the observation goes straight from where it is (host memory in a rollout) into the graph's static input buffers
            ro = obs["robot_obs_raw"]
            action = self.engine.infer_act(obs["rgb_obs"]["rgb_static"][0], obs["rgb_obs"]["rgb_gripper"][0], ro if ro.dtype == torch.float32 else ro.float())
        else:
            action = self.engine.infer_act(st[0].contiguous(), gr[0].contiguous(), obs["robot_obs_raw"].to(dev).float().reshape(1, -1), sample_u=sample_u)
        self.rollout_step_counter += 1
        return action

    def enable_device_augmentation(self, static_pad: int = 10, gripper_pad: int = 4):
        """Run the training transforms of conf/datamodule/transforms/rand_shift.yaml on the device for uint8 frames: RandomShiftsAug (pad 10 / 4)
        fused with scale + normalise — the datamodule then only has to hand over the stored uint8 frames."""
        self.engine.set_augmentation(static_pad, gripper_pad)

    def enable_cuda_graph_inference(self, flag: bool = True):
        """Replay `step`'s per-control-step work (encoders, one decoder step, sampling, frame change) from one CUDA graph."""
        self.engine.enable_infer_graph(flag)

    def _dist(self, state: torch.Tensor):
        """Distribution.get_dist (hulc/utils/distributions.py:38-47) on a state tensor: logits [B, 1024] or [mean | raw std] [B, 512]."""
        import torch.distributions as D

        if self.engine.discrete:
            d = self.engine.dims
            logits = state.view(*state.shape[:-1], d.category_size, d.class_size)
            return D.Independent(D.OneHotCategoricalStraightThrough(logits=logits), 1)
        mean, raw = state.chunk(2, dim=-1)
        return D.Independent(D.Normal(mean, torch.nn.functional.softplus(raw) + 1e-4), 1)

    @torch.no_grad()
    def lmp_train(self, perceptual_emb, latent_goal, train_acts, robot_obs, **inject):
        """hulc.py:254-299 for one modality, from already-encoded inputs: returns `(kl_loss, action_loss, total_loss, pp_dist, pr_dist,
        seq_feat)` — prior, posterior, a plan sampled from the posterior, the decoder's mixture NLL (+ gripper CE) and the balanced, scaled
        KL, all on the same kernels as `training_step`.  Forward values: `training_step` does not call this method, it runs the same block
        fused with the encoders, both modalities batched, and with the hand-written backward; the intermediate tensors of this call are kept
        in `self.last_lmp_outputs`.  `inject` (plan_idx / plan_u / plan_eps / dropout_masks, keyed by "vis") pins the randomness for parity runs."""
        if self.engine.model == "gcbc":
            raise NotImplementedError("GCBC has no latent plan: its training_step calls the decoder directly (gcbc.py:50-181)")
        out = self.engine.lmp_forward(perceptual_emb, latent_goal, train_acts, robot_obs, **inject)
        self.last_lmp_outputs = out
        kl, act = out["kl_loss_vis"], out["action_loss_vis"]
        return kl, act, act + kl, self._dist(out["pp_state"]), self._dist(out["pr_state"]), out["seq_feat"]

    @torch.no_grad()
    def clip_auxiliary_loss(self, seq_vis_feat, encoded_lang, use_for_aux_loss=None):
        """hulc.py:650-695: symmetric cross-entropy over the cosine-similarity logits of the projected sequence features and language
        goals of the rows `use_for_aux_loss` selects (0 when it selects none).  Forward value; inside `training_step` the same kernel
        also emits the gradient."""
        if not self.use_clip_auxiliary_loss:
            raise RuntimeError("use_clip_auxiliary_loss is off")
        return self.engine.clip_forward(seq_vis_feat, encoded_lang, use_for_aux_loss)

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

    def on_load_checkpoint(self, checkpoint):
        """state_dict keys are the reference's, nothing to translate; the optimizer state (torch.optim.Adam layout) is restored by
        FusedAdam.load_state_dict, which Lightning calls with `checkpoint["optimizer_states"][0]`."""

