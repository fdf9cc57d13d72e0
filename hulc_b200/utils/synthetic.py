"""Seeded synthetic configs, weights and batches shared by tests, `bench.py`, `smoke()` and the golden script.

Everything is a pure function of integers so the build container (which has `/root/reference` and writes the
golden fixtures) and the GPU box (which has not) regenerate identical tensors with the same torch version.

* `model_config(...)`    — the resolved Hydra tree of `conf/model/{hulc,gcbc,mcil}.yaml` as nested dicts
                           (reference: conf/model/*.yaml, conf/loss/default.yaml, conf/datamodule/default.yaml).
* `fill_state_dict_(sd)` — deterministic values for every parameter, keyed on the state_dict key.
* `make_batch(...)`      — the `{"vis": ..., "lang": ...}` batch contract of `Hulc.training_step`
                           (reference: hulc/models/hulc.py:390-419).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Optional

import torch

from ..spec import param_spec  # noqa: F401  (re-exported: tests and tools reach it through this module too)


class AttrDict(dict):
    """Attribute-style dict (stands in for omegaconf.DictConfig when omegaconf is not installed)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _attr(d):
    if isinstance(d, dict):
        return AttrDict({k: _attr(v) for k, v in d.items()})
    return d


def model_config(
    model: str = "hulc",
    rnn_model: str = "rnn_decoder",
    max_window: int = 32,
    dropout_p: float = 0.1,
    target_root: str = "hulc",
    bc_z: bool = False,
    mia: bool = False,
) -> AttrDict:
    """Resolved `conf/model/<model>.yaml`.  `target_root` is the package the `_target_` strings point into.  The ablation switches select
    `model/bc_z_lang_decoder=default` (+ use_bc_z_auxiliary_loss) and `model/mia_lang_discriminator=default` (+ use_mia_auxiliary_loss)."""
    r = target_root
    action_space = 7
    act_max = [1.0] * 7
    act_min = [-1.0] * 7
    cfg: Dict = dict(
        perceptual_encoder=dict(
            _target_=f"{r}.models.perceptual_encoders.concat_encoders.ConcatEncoders",
            _recursive_=False,
            rgb_static=dict(
                _target_=f"{r}.models.perceptual_encoders.vision_network.VisionNetwork",
                input_width=200,
                input_height=200,
                activation_function="ReLU",
                dropout_vis_fc=0.0,
                l2_normalize_output=False,
                visual_features=64,
                num_c=3,
                use_sinusoid=False,
                spatial_softmax_temp=1.0,
            ),
            rgb_gripper=dict(
                _target_=f"{r}.models.perceptual_encoders.vision_network_gripper.VisionNetwork",
                input_width=84,
                input_height=84,
                activation_function="ReLU",
                dropout_vis_fc=0.0,
                l2_normalize_output=False,
                visual_features=64,
                conv_encoder="nature_cnn",
                num_c=3,
            ),
            depth_static=None,
            depth_gripper=None,
            proprio=None,
            tactile=None,
        ),
        plan_proposal=dict(
            _target_=f"{r}.models.plan_encoders.plan_proposal_net.PlanProposalNetwork",
            perceptual_features=None,
            latent_goal_features=32,
            plan_features=None,
            activation_function="ReLU",
            hidden_size=2048,
        ),
        plan_recognition=dict(
            _target_=f"{r}.models.plan_encoders.plan_recognition_net.PlanRecognitionTransformersNetwork",
            num_heads=8,
            num_layers=2,
            encoder_hidden_size=2048,
            fc_hidden_size=4096,
            in_features=None,
            plan_features=None,
            action_space=action_space,
            dropout_p=dropout_p,
            encoder_normalize=False,
            positional_normalize=False,
            position_embedding=True,
            max_position_embeddings=max_window,
        ),
        distribution=dict(_target_=f"{r}.utils.distributions.Distribution", dist="discrete", category_size=32, class_size=32),
        visual_goal=dict(
            _target_=f"{r}.models.encoders.goal_encoders.VisualGoalEncoder",
            in_features=None,
            hidden_size=2048,
            latent_goal_features=32,
            l2_normalize_goal_embeddings=False,
            activation_function="ReLU",
        ),
        language_goal=dict(
            _target_=f"{r}.models.encoders.goal_encoders.LanguageGoalEncoder",
            in_features=384,
            hidden_size=2048,
            latent_goal_features=32,
            l2_normalize_goal_embeddings=False,
            activation_function="ReLU",
            word_dropout_p=0.0,
        ),
        action_decoder=dict(
            _target_=f"{r}.models.decoders.logistic_decoder_rnn.LogisticDecoderRNN",
            n_mixtures=10,
            hidden_size=2048,
            out_features=action_space,
            log_scale_min=-7.0,
            act_max_bound=act_max,
            act_min_bound=act_min,
            dataset_dir="",
            load_action_bounds=False,
            num_classes=10,
            latent_goal_features=32,
            plan_features=None,
            perceptual_features=None,
            gripper_alpha=1.0,
            perceptual_emb_slice=[64, 128],
            policy_rnn_dropout_p=0.0,
            num_layers=2,
            rnn_model=rnn_model,
            gripper_control=True,
            discrete_gripper=True,
        ),
        optimizer=dict(_target_="torch.optim.Adam", lr=2e-4),
        lr_scheduler=dict(_target_="transformers.get_constant_schedule"),
        bc_z_lang_decoder=None,
        mia_lang_discriminator=None,
        proj_vis_lang=dict(
            _target_=f"{r}.models.auxiliary_loss_networks.proj_vis_lang.ProjVisLang",
            im_dim=4096,
            lang_dim=32,
            output_dim=32,
            proj_lang=True,
        ),
        val_instructions=None,
        kl_beta=0.01,
        kl_balancing_mix=0.8,
        state_recons=False,
        state_recon_beta=0.5,
        use_bc_z_auxiliary_loss=False,
        bc_z_auxiliary_loss_beta=1.0,
        use_mia_auxiliary_loss=False,
        mia_auxiliary_loss_beta=1.0,
        replan_freq=30,
        use_clip_auxiliary_loss=True,
        clip_auxiliary_loss_beta=3.0,
    )
    top = f"{r}.models.hulc.Hulc"
    if model == "gcbc":
        top = f"{r}.models.gcbc.GCBC"
    elif model == "mcil":
        # conf/model/mcil.yaml: BiRNN posterior, continuous latent, 7-dim logistic decoder without TCP/CE/CLIP
        cfg["plan_recognition"] = dict(
            _target_=f"{r}.models.plan_encoders.plan_recognition_net.PlanRecognitionBiRNNNetwork",
            in_features=None,
            plan_features=256,
            action_space=action_space,
            birnn_dropout_p=0.0,
            rnn_type="nn.RNN",
        )
        cfg["distribution"] = dict(_target_=f"{r}.utils.distributions.Distribution", dist="continuous", plan_features=256)
        ad = cfg["action_decoder"]
        ad.update(num_classes=256, gripper_control=False, discrete_gripper=False)
        ad.pop("perceptual_emb_slice")
        cfg["proj_vis_lang"] = None
        cfg["use_clip_auxiliary_loss"] = False
    elif model != "hulc":
        raise ValueError(model)
    if bc_z:  # conf/model/bc_z_lang_decoder/default.yaml
        cfg["bc_z_lang_decoder"] = dict(_target_=f"{r}.models.auxiliary_loss_networks.bc_z_lang_decoder.BCZLangDecoder", in_features=4096, lang_dim=384)
        cfg["use_bc_z_auxiliary_loss"] = True
    if mia:  # conf/model/mia_lang_discriminator/default.yaml
        cfg["mia_lang_discriminator"] = dict(_target_=f"{r}.models.auxiliary_loss_networks.mia_lang_discriminator.MIALangDiscriminator", in_features=32, lang_dim=32,
                                             dropout_p=0.0)
        cfg["use_mia_auxiliary_loss"] = True
    cfg["_target_"] = top
    cfg["_recursive_"] = False
    return _attr(cfg)


def _key_seed(key: str, salt: int) -> int:
    return (zlib.crc32(key.encode()) + 7919 * salt) & 0x7FFFFFFF


@torch.no_grad()
def fill_state_dict_(sd: Dict[str, torch.Tensor], salt: int = 0) -> Dict[str, torch.Tensor]:
    """Overwrite every floating-point *parameter-like* entry of `sd` in place with a value that depends only on
    (key, shape, salt).  Buffers of the reference (`x_map`, `y_map`, `temperature`, `one_hot_embedding_eye`, `ones`,
    `gripper_bounds`, `action_*_bound`) keep their constructor values.

    Scales follow PyTorch's defaults (U(±1/sqrt(fan_in))) so activations stay O(1) and the ReLU-RNN is stable
    (spectral radius of W_hh ≈ 0.58); LayerNorm gains are 1±0.1 and all biases are non-zero so their gradients
    are exercised.
    """
    buffers = ("x_map", "y_map", "temperature", "one_hot_embedding_eye", ".ones", "gripper_bounds", "_bound")
    for key, t in sd.items():
        if not torch.is_floating_point(t) or any(b in key for b in buffers) or key == "ones":
            continue
        g = torch.Generator().manual_seed(_key_seed(key, salt))
        u = torch.rand(t.shape, generator=g, dtype=torch.float32) * 2 - 1
        if key == "logit_scale":
            v = torch.full((), math.log(1 / 0.07))
        elif t.dim() >= 2 and "position_embeddings" not in key:
            fan_in = 1
            for s in t.shape[1:]:
                fan_in *= s
            v = u / math.sqrt(fan_in)
        elif "position_embeddings" in key:
            v = 0.5 * u
        elif key.endswith("weight"):  # LayerNorm gains (1-D weights)
            v = 1.0 + 0.1 * u
        else:  # biases
            v = 0.05 * u
        t.copy_(v.to(t.dtype))
    return sd


def make_modality(
    modality: str, batch: int, seq: int, seed: int = 1, device="cpu", static_hw: int = 200, gripper_hw: int = 84
) -> Dict:
    """One entry of the training batch (reference contract: hulc/models/hulc.py:395-413, dataset/README.md:52-118)."""
    g = torch.Generator().manual_seed(seed * 1000003 + (0 if modality == "vis" else 1))
    B, S = batch, seq

    def U(*shape):
        return torch.rand(*shape, generator=g, dtype=torch.float32) * 2 - 1

    def N(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    rgb_static = U(B, S, 3, static_hw, static_hw)
    rgb_gripper = U(B, S, 3, gripper_hw, gripper_hw)
    robot_obs = N(B, S, 8)
    actions = U(B, S, 7)
    actions[..., 6] = (torch.rand(B, S, generator=g) < 0.5).float() * 2 - 1
    raw = N(B, S, 15)
    raw[..., :3] *= 0.3
    raw[..., 3:6] = U(B, S, 3)
    d = {
        "rgb_obs": {"rgb_static": rgb_static, "rgb_gripper": rgb_gripper},
        "depth_obs": {},
        "robot_obs": robot_obs,
        "actions": actions,
        "state_info": {"robot_obs": raw},
        "idx": torch.arange(B),
    }
    if modality == "lang":
        d["lang"] = N(B, 384)
        d["use_for_aux_lang_loss"] = torch.ones(B, dtype=torch.bool)
    return _to(d, device)


def make_batch(batch: int, seq: int, seed: int = 1, device="cpu", **kw) -> Dict[str, Dict]:
    return {m: make_modality(m, batch, seq, seed, device, **kw) for m in ("vis", "lang")}


def _to(d, device):
    if isinstance(d, dict):
        return {k: _to(v, device) for k, v in d.items()}
    if torch.is_tensor(d):
        return d.to(device)
    return d


def plan_noise(batch: int, seq: int, modality: str, seed: int = 1, n_cat: int = 32, n_cont: int = 256) -> Dict[str, torch.Tensor]:
    """Injected randomness for parity runs: one uniform per (sequence, category) driving the inverse-CDF categorical
    sample (discrete latent) and one standard normal per latent dim (continuous latent)."""
    g = torch.Generator().manual_seed(seed * 7777 + (3 if modality == "vis" else 5))
    return {
        "u": torch.rand(batch, n_cat, generator=g, dtype=torch.float32),
        "eps": torch.randn(batch, n_cont, generator=g, dtype=torch.float32),
    }


def validation_noise(batch: int, seq: int, modality: str, which: str, seed: int = 1, n_dims: int = 6, n_mix: int = 10, n_cat: int = 32,
                     n_cont: int = 256) -> Dict[str, torch.Tensor]:
    """Injected randomness of the validation path for plan source `which` ("pp" | "pr"): the categorical-sample uniforms / Normal noise of
    the latent plan and the two uniform draws of LogisticDecoderRNN._sample (Gumbel-max over the mixture, inverse-CDF logistic)."""
    g = torch.Generator().manual_seed(seed * 4241 + (3 if modality == "vis" else 5) + (100 if which == "pp" else 200))
    return {
        "u": torch.rand(batch, n_cat, generator=g, dtype=torch.float32),
        "eps": torch.randn(batch, n_cont, generator=g, dtype=torch.float32),
        "u_mix": torch.rand(batch, seq, n_dims, n_mix, generator=g, dtype=torch.float32),
        "u_inv": torch.rand(batch, seq, n_dims, generator=g, dtype=torch.float32),
    }


def rollout_inputs(steps: int, goal_kind: str = "lang", seed: int = 1) -> Dict[str, torch.Tensor]:
    """Seeded inputs of an inference rollout (Hulc.step, hulc.py:851-870): `steps` observations, a language embedding or a goal image, and
    the injected randomness (one categorical-sample uniform vector per re-plan, the two _sample draws per control step)."""
    vis = make_modality("vis", 1, steps, seed=seed + 50)
    other = make_modality("lang", 1, 1, seed=seed + 51)
    g = torch.Generator().manual_seed(seed * 6151 + (1 if goal_kind == "lang" else 2))
    return {
        "rgb_static": vis["rgb_obs"]["rgb_static"][0], "rgb_gripper": vis["rgb_obs"]["rgb_gripper"][0],          # (steps, 3, H, W)
        "robot_obs": vis["robot_obs"][0], "robot_obs_raw": vis["state_info"]["robot_obs"][0],                      # (steps, 8) / (steps, 15)
        "lang": other["lang"],                                                                                    # (1, 384)
        "goal_static": other["rgb_obs"]["rgb_static"][0], "goal_gripper": other["rgb_obs"]["rgb_gripper"][0],      # (1, 3, H, W)
        "goal_robot_obs": other["robot_obs"][0],
        "plan_u": torch.rand(steps, 32, generator=g), "u_mix": torch.rand(steps, 1, 1, 6, 10, generator=g), "u_inv": torch.rand(steps, 1, 1, 6, generator=g),
    }


def dropout_masks(
    batch: int, seq: int, modality: str, p: float, seed: int = 1, d_model: int = 128, nhead: int = 8, ff: int = 2048, nlayers: int = 2
) -> Dict[str, torch.Tensor]:
    """Injected keep-masks (bool, True = keep) for the 9 dropout sites of the posterior transformer
    (reference: plan_recognition_net.py:83-89,111) in batch-first layouts: "in"/"l{i}.drop1"/"l{i}.drop2" (B,S,D),
    "l{i}.attn" (B,H,S,S), "l{i}.ffn" (B,S,FF)."""
    g = torch.Generator().manual_seed(seed * 9176 + (11 if modality == "vis" else 13))

    def keep(*shape):
        return torch.rand(*shape, generator=g) >= p

    m = {"in": keep(batch, seq, d_model)}
    for i in range(nlayers):
        m[f"l{i}.attn"] = keep(batch, nhead, seq, seq)
        m[f"l{i}.drop1"] = keep(batch, seq, d_model)
        m[f"l{i}.ffn"] = keep(batch, seq, ff)
        m[f"l{i}.drop2"] = keep(batch, seq, d_model)
    return m


def make_state_dict(model: str = "hulc", rnn_model: str = "rnn_decoder", max_window: int = 32, salt: int = 0, dims=None) -> Dict[str, torch.Tensor]:
    """Seeded parameters for `param_spec` (fp32, CPU); `dims` (spec.ModelDims) selects a non-shipped variant."""
    sd = {k: torch.empty(s, dtype=torch.float32) for k, s in param_spec(model, rnn_model, max_window, dims=dims).items()}
    return fill_state_dict_(sd, salt)
