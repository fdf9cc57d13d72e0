"""Model dimensions and the parameter layout, read from the reference's Hydra config tree.

`dims_from_configs` walks the DictConfigs `Hulc.__init__` receives (hulc/models/hulc.py:58-187 wires them the same way:
`setup_input_sizes` fills `perceptual_features`, `plan_features`, `in_features` from the encoder's `latent_size` and the
distribution) and returns every size the engine needs.  A value the kernels cannot honour raises `NotImplementedError`
naming the config key — nothing is silently ignored.  `param_spec` turns the dimensions into the reference's
`state_dict` contract (parameter key -> shape, registration order).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Dict, Optional, Tuple


@dataclass(frozen=True)
class ModelDims:
    """Defaults are the shipped YAML values (conf/model/*.yaml and the sub-trees they include)."""

    model: str = "hulc"                 # hulc | gcbc | mcil
    rnn_model: str = "rnn_decoder"      # rnn_decoder | gru_decoder (decoders/utils/rnn.py:5-36)
    max_window: int = 32                # plan_recognition.max_position_embeddings
    static_hw: int = 200                # perceptual_encoder.rgb_static.input_{width,height}
    gripper_hw: int = 84
    visual_features: int = 64           # per camera; latent_size = 128 (concat_encoders.py:37-40)
    spatial_softmax_temp: float = 1.0
    latent_goal: int = 32
    lang_in: int = 384                  # language_goal.in_features
    goal_hidden_vis: int = 2048
    goal_hidden_lang: int = 2048
    prior_hidden: int = 2048            # plan_proposal.hidden_size
    nhead: int = 8
    nlayers: int = 2
    ffn_hidden: int = 2048              # plan_recognition.encoder_hidden_size
    fc_hidden: int = 4096               # plan_recognition.fc_hidden_size (transformer) / 2 x birnn hidden
    dropout_p: float = 0.1
    category_size: int = 32
    class_size: int = 32
    cont_plan: int = 256                # distribution.plan_features (continuous latent, MCIL)
    dec_hidden: int = 2048
    n_mix: int = 10
    out_features: int = 7
    num_classes: int = 10
    log_scale_min: float = -7.0
    act_min: float = -1.0
    act_max: float = 1.0
    gripper_alpha: float = 1.0
    clip_hidden: int = 128              # ProjVisLang's hidden width (proj_vis_lang.py:10-21, fixed in the reference)
    clip_out: int = 32
    # ablation blocks, all off in the shipped model YAMLs
    bc_z: bool = False                  # use_bc_z_auxiliary_loss + bc_z_lang_decoder (hulc.py:567-604, bc_z_lang_decoder.py:5-20)
    mia: bool = False                   # use_mia_auxiliary_loss + mia_lang_discriminator (hulc.py:606-648, mia_lang_discriminator.py:5-21)
    aux_hidden: int = 512               # hidden width of both auxiliary MLPs (fixed in the reference)

    @classmethod
    def shipped(cls, model: str = "hulc", rnn_model: str = "rnn_decoder", max_window: int = 32, **kw) -> "ModelDims":
        """The shipped configuration of `model` (conf/model/{hulc,gcbc,mcil}.yaml): MCIL discretises into 256 classes and has no dropout."""
        base = dict(model=model, rnn_model=rnn_model, max_window=max_window)
        if model == "mcil":
            base.update(num_classes=256, dropout_p=0.0)
        base.update(kw)
        return cls(**base)

    @property
    def latent_size(self) -> int:
        return 2 * self.visual_features

    @property
    def discrete(self) -> bool:
        return self.model != "mcil"

    @property
    def plan_features(self) -> int:
        return {"hulc": self.category_size * self.class_size, "gcbc": 0, "mcil": self.cont_plan}[self.model]

    @property
    def state_dim(self) -> int:
        """Width of the prior's / posterior's output state: logits (discrete) or [mean | raw std] (continuous)."""
        plan = self.category_size * self.class_size if self.model != "mcil" else self.cont_plan
        return plan if self.model != "mcil" else 2 * plan

    @property
    def n_dims(self) -> int:
        """Action dimensions under the logistic mixture (the gripper has its own 2-way head unless MCIL)."""
        return self.out_features - 1 if self.model != "mcil" else self.out_features

    @property
    def percep_lo(self) -> int:
        """The decoder sees perceptual_emb[..., 64:128] (action_decoder.perceptual_emb_slice) unless MCIL."""
        return self.visual_features if self.model != "mcil" else 0

    def conv_out(self, hw: int) -> int:
        for k, s in ((8, 4), (4, 2), (3, 1)):
            hw = (hw - k) // s + 1
        return hw


def _get(cfg, key, default=None):
    if cfg is None:
        return default
    try:
        return cfg[key] if key in cfg else default
    except TypeError:
        return getattr(cfg, key, default)


def _require(cfg, path: str, key: str, allowed, default):
    """cfg[key] must be one of `allowed` (the only behaviour the kernels implement)."""
    v = _get(cfg, key, default)
    if v not in allowed:
        raise NotImplementedError(f"{path}.{key}={v!r} is not supported by the hulc_b200 kernels (supported: {list(allowed)})")
    return v


def dims_from_configs(model: str, perceptual_encoder, plan_proposal, plan_recognition, language_goal, visual_goal, action_decoder, distribution,
                      proj_vis_lang=None, bc_z_lang_decoder=None, mia_lang_discriminator=None) -> ModelDims:
    """Every size of the network from the config tree; raises on anything the kernels cannot run."""
    d = ModelDims.shipped(model)
    # ---- perceptual encoders (concat_encoders.py:20-57, vision_network.py:17-53, vision_network_gripper.py:24-47) ----
    for name in ("depth_static", "depth_gripper", "proprio", "tactile"):
        if _get(perceptual_encoder, name) not in (None, {}, "none"):
            raise NotImplementedError(f"perceptual_encoder.{name} is disabled in conf/model/perceptual_encoder/gripper_cam.yaml and not built")
    st, gr = _get(perceptual_encoder, "rgb_static"), _get(perceptual_encoder, "rgb_gripper")
    if st is None or gr is None:
        raise NotImplementedError("perceptual_encoder needs both rgb_static and rgb_gripper (conf/model/perceptual_encoder/gripper_cam.yaml)")
    hw = {}
    for tag, c in (("rgb_static", st), ("rgb_gripper", gr)):
        p = f"perceptual_encoder.{tag}"
        _require(c, p, "activation_function", ("ReLU",), "ReLU")
        _require(c, p, "dropout_vis_fc", (0, 0.0), 0.0)
        _require(c, p, "l2_normalize_output", (False,), False)
        _require(c, p, "num_c", (3,), 3)
        _require(c, p, "visual_features", (64,), 64)
        w, h = int(_get(c, "input_width", 0)), int(_get(c, "input_height", 0))
        if w != h or w < 36:
            raise NotImplementedError(f"{p}.input_width/height = {w}x{h}: square frames of at least 36 pixels")
        hw[tag] = w
    _require(st, "perceptual_encoder.rgb_static", "use_sinusoid", (False,), False)
    _require(gr, "perceptual_encoder.rgb_gripper", "conv_encoder", ("nature_cnn",), "nature_cnn")
    d = replace(d, static_hw=hw["rgb_static"], gripper_hw=hw["rgb_gripper"], spatial_softmax_temp=float(_get(st, "spatial_softmax_temp", 1.0)))
    # ---- goal encoders (goal_encoders.py:20-30, 52-63) -----------------------------------------------------------------
    for tag, c in (("visual_goal", visual_goal), ("language_goal", language_goal)):
        _require(c, tag, "activation_function", ("ReLU",), "ReLU")
        _require(c, tag, "l2_normalize_goal_embeddings", (False,), False)
        _require(c, tag, "latent_goal_features", (32,), 32)
    _require(language_goal, "language_goal", "word_dropout_p", (0, 0.0), 0.0)
    d = replace(d, goal_hidden_vis=int(_get(visual_goal, "hidden_size", 2048)), goal_hidden_lang=int(_get(language_goal, "hidden_size", 2048)),
                lang_in=int(_get(language_goal, "in_features", 384)))
    # ---- prior (plan_proposal_net.py:15-40) ----------------------------------------------------------------------------
    _require(plan_proposal, "plan_proposal", "activation_function", ("ReLU",), "ReLU")
    _require(plan_proposal, "plan_proposal", "latent_goal_features", (32,), 32)
    d = replace(d, prior_hidden=int(_get(plan_proposal, "hidden_size", 2048)))
    # ---- posterior + distribution (plan_recognition_net.py:14-92, distributions.py:15-60) ------------------------------
    birnn = str(_get(plan_recognition, "_target_", "")).endswith("PlanRecognitionBiRNNNetwork")
    continuous = _get(distribution, "dist", "discrete") == "continuous"
    if continuous != birnn or birnn != (model == "mcil"):
        raise NotImplementedError("supported latent plans: transformer posterior + discrete latent (hulc/gcbc) or BiRNN posterior + continuous latent (mcil)")
    if birnn:
        _require(plan_recognition, "plan_recognition", "birnn_dropout_p", (0, 0.0), 0.0)
        _require(plan_recognition, "plan_recognition", "rnn_type", ("nn.RNN",), "nn.RNN")
        d = replace(d, cont_plan=int(_get(distribution, "plan_features", 256)), dropout_p=0.0)
    else:
        _require(distribution, "distribution", "category_size", (32,), 32)
        _require(distribution, "distribution", "class_size", (32,), 32)
        for k in ("encoder_normalize", "positional_normalize"):
            _require(plan_recognition, "plan_recognition", k, (False,), False)
        _require(plan_recognition, "plan_recognition", "position_embedding", (True,), True)
        nhead = int(_get(plan_recognition, "num_heads", 8))
        if d.latent_size % nhead or d.latent_size // nhead not in (8, 16, 32, 64):
            raise NotImplementedError(f"plan_recognition.num_heads={nhead}: the attention kernel takes head sizes 8/16/32/64 of a 128-wide model without padding")
        d = replace(d, nhead=nhead, nlayers=int(_get(plan_recognition, "num_layers", 2)), ffn_hidden=int(_get(plan_recognition, "encoder_hidden_size", 2048)),
                    fc_hidden=int(_get(plan_recognition, "fc_hidden_size", 4096)), dropout_p=float(_get(plan_recognition, "dropout_p", 0.0) or 0.0),
                    max_window=int(_get(plan_recognition, "max_position_embeddings", 32)))
    # ---- action decoder (logistic_decoder_rnn.py:27-83) ----------------------------------------------------------------
    ad = action_decoder
    target = str(_get(ad, "_target_", "LogisticDecoderRNN"))
    if not target.endswith("LogisticDecoderRNN"):
        # (the reference's DeterministicDecoder cannot be constructed: deterministic_decoder.py:33 evaluates the name `rnn_decoder`, which that
        #  module never imports — NameError in the unmodified reference, so there is nothing to be a drop-in for)
        raise NotImplementedError(f"action_decoder._target_={_get(ad, '_target_')!r}: LogisticDecoderRNN is the decoder the reference can build")
    rnn_model = _require(ad, "action_decoder", "rnn_model", ("rnn_decoder", "gru_decoder"), "rnn_decoder")
    _require(ad, "action_decoder", "num_layers", (2,), 2)
    _require(ad, "action_decoder", "policy_rnn_dropout_p", (0, 0.0), 0.0)
    _require(ad, "action_decoder", "latent_goal_features", (32,), 32)
    _require(ad, "action_decoder", "out_features", (7,), 7)
    _require(ad, "action_decoder", "load_action_bounds", (False,), False)
    grip = model != "mcil"
    _require(ad, "action_decoder", "gripper_control", (grip,), grip)
    _require(ad, "action_decoder", "discrete_gripper", (grip,), grip)
    sl = _get(ad, "perceptual_emb_slice")
    if (list(sl) if sl is not None else None) != ([64, 128] if grip else None):
        raise NotImplementedError(f"action_decoder.perceptual_emb_slice={sl!r}: [64, 128] for hulc/gcbc, unset for mcil")
    lo, hi = _get(ad, "act_min_bound", [-1.0] * 7), _get(ad, "act_max_bound", [1.0] * 7)
    if len(set(float(v) for v in lo)) != 1 or len(set(float(v) for v in hi)) != 1:
        raise NotImplementedError("action_decoder.act_{min,max}_bound: one bound shared by all action dimensions")
    d = replace(d, rnn_model=rnn_model, dec_hidden=int(_get(ad, "hidden_size", 2048)), n_mix=int(_get(ad, "n_mixtures", 10)), num_classes=int(_get(ad, "num_classes", 10)),
                log_scale_min=float(_get(ad, "log_scale_min", -7.0)), act_min=float(lo[0]), act_max=float(hi[0]), gripper_alpha=float(_get(ad, "gripper_alpha", 1.0)))
    return _aux_heads(d, model, proj_vis_lang, bc_z_lang_decoder, mia_lang_discriminator)


def _aux_heads(d: ModelDims, model: str, proj_vis_lang, bc_z_lang_decoder, mia_lang_discriminator) -> ModelDims:
    if d.dec_hidden % 4 or d.prior_hidden % 4 or d.ffn_hidden % 4 or d.fc_hidden % 4:
        raise NotImplementedError("hidden sizes must be multiples of 4 (16-byte rows)")
    # ---- BC-Z language regression head / MIA discriminator (bc_z_lang_decoder.py:5-20, mia_lang_discriminator.py:5-21) --
    if bc_z_lang_decoder not in (None, {}, "none"):
        if model == "mcil":
            raise NotImplementedError("bc_z_lang_decoder is built for hulc / gcbc")
        if int(_get(bc_z_lang_decoder, "in_features", d.fc_hidden)) != d.fc_hidden or int(_get(bc_z_lang_decoder, "lang_dim", d.lang_in)) != d.lang_in:
            raise NotImplementedError("bc_z_lang_decoder: in_features = plan_recognition.fc_hidden_size, lang_dim = language_goal.in_features (conf/model/bc_z_lang_decoder/default.yaml)")
        d = replace(d, bc_z=True)
    if mia_lang_discriminator not in (None, {}, "none"):
        if model == "mcil":
            raise NotImplementedError("mia_lang_discriminator is built for hulc / gcbc")
        _require(mia_lang_discriminator, "mia_lang_discriminator", "dropout_p", (0, 0.0), 0.0)
        od = int(_get(proj_vis_lang, "output_dim", 32))
        if int(_get(mia_lang_discriminator, "in_features", od)) != od or int(_get(mia_lang_discriminator, "lang_dim", od)) != od:
            raise NotImplementedError("mia_lang_discriminator: in_features = lang_dim = proj_vis_lang.output_dim (conf/model/mia_lang_discriminator/default.yaml)")
        d = replace(d, mia=True)
    # ---- CLIP projection head (proj_vis_lang.py:7-27) ------------------------------------------------------------------
    if model != "mcil":
        if proj_vis_lang is None:
            raise NotImplementedError("use_clip_auxiliary_loss needs proj_vis_lang (conf/model/proj_vis_lang/default.yaml)")
        _require(proj_vis_lang, "proj_vis_lang", "proj_lang", (True,), True)
        _require(proj_vis_lang, "proj_vis_lang", "lang_dim", (32,), 32)
        if int(_get(proj_vis_lang, "im_dim", d.fc_hidden)) != d.fc_hidden:
            raise NotImplementedError(f"proj_vis_lang.im_dim={_get(proj_vis_lang, 'im_dim')} must equal plan_recognition.fc_hidden_size={d.fc_hidden}")
        d = replace(d, clip_out=int(_get(proj_vis_lang, "output_dim", 32)))
    return d


def param_spec(model: str = "hulc", rnn_model: str = "rnn_decoder", max_window: int = 32, dims: Optional[ModelDims] = None) -> Dict[str, tuple]:
    """state_dict contract of the reference: parameter key -> shape, in registration order, for
    `conf/model/{hulc,gcbc,mcil}.yaml` (or the sizes in `dims`).  Buffers are not listed."""
    d = dims if dims is not None else ModelDims.shipped(model, rnn_model, max_window)
    model, rnn_model = d.model, d.rnn_model
    spec: Dict[str, tuple] = {}
    D, G = d.latent_size, d.latent_goal

    def lin(name, n_out, n_in):
        spec[f"{name}.weight"] = (n_out, n_in)
        spec[f"{name}.bias"] = (n_out,)

    def convs(p):
        spec[f"{p}.conv_model.0.weight"], spec[f"{p}.conv_model.0.bias"] = (32, 3, 8, 8), (32,)
        spec[f"{p}.conv_model.2.weight"], spec[f"{p}.conv_model.2.bias"] = (64, 32, 4, 4), (64,)
        spec[f"{p}.conv_model.4.weight"], spec[f"{p}.conv_model.4.bias"] = (64, 64, 3, 3), (64,)

    def rnn(p, n_in, hidden, layers, gates=1, bidir=False):
        for l in range(layers):
            for sfx in ("", "_reverse") if bidir else ("",):
                i = n_in if l == 0 else hidden * (2 if bidir else 1)
                spec[f"{p}.weight_ih_l{l}{sfx}"] = (gates * hidden, i)
                spec[f"{p}.weight_hh_l{l}{sfx}"] = (gates * hidden, hidden)
                spec[f"{p}.bias_ih_l{l}{sfx}"] = (gates * hidden,)
                spec[f"{p}.bias_hh_l{l}{sfx}"] = (gates * hidden,)

    if model != "mcil":
        spec["logit_scale"] = ()
    pe = "perceptual_encoder.rgb_static_encoder"
    convs(pe)
    lin(f"{pe}.fc1.0", 512, 128), lin(f"{pe}.fc2", d.visual_features, 512)
    spec[f"{pe}.ln.weight"], spec[f"{pe}.ln.bias"] = (d.visual_features,), (d.visual_features,)
    pg = "perceptual_encoder.rgb_gripper_encoder"
    convs(pg)
    k = d.conv_out(d.gripper_hw)
    lin(f"{pg}.conv_model.7", 128, 64 * k * k), lin(f"{pg}.fc1.0", 512, 128), lin(f"{pg}.fc2", d.visual_features, 512)
    spec[f"{pg}.ln.weight"], spec[f"{pg}.ln.bias"] = (d.visual_features,), (d.visual_features,)
    state = d.state_dim
    Hp = d.prior_hidden
    lin("plan_proposal.fc_model.0", Hp, D + G)
    for i in (2, 4, 6):
        lin(f"plan_proposal.fc_model.{i}", Hp, Hp)
    lin("plan_proposal.fc_state.0", state, Hp)
    if model == "mcil":
        rnn("plan_recognition.birnn_model", D, 2048, 2, bidir=True)
        lin("plan_recognition.fc_state.0", state, 4096)
    else:
        spec["plan_recognition.position_embeddings.weight"] = (d.max_window, D)
        for l in range(d.nlayers):
            p = f"plan_recognition.transformer_encoder.layers.{l}"
            spec[f"{p}.self_attn.in_proj_weight"], spec[f"{p}.self_attn.in_proj_bias"] = (3 * D, D), (3 * D,)
            lin(f"{p}.self_attn.out_proj", D, D)
            lin(f"{p}.linear1", d.ffn_hidden, D), lin(f"{p}.linear2", D, d.ffn_hidden)
            spec[f"{p}.norm1.weight"], spec[f"{p}.norm1.bias"] = (D,), (D,)
            spec[f"{p}.norm2.weight"], spec[f"{p}.norm2.bias"] = (D,), (D,)
        lin("plan_recognition.fc", d.fc_hidden, D)
        lin("plan_recognition.fc_state.0", state, d.fc_hidden)
    lin("visual_goal.mlp.0", d.goal_hidden_vis, D), lin("visual_goal.mlp.2", d.goal_hidden_vis, d.goal_hidden_vis), lin("visual_goal.mlp.4", G, d.goal_hidden_vis)
    spec["visual_goal.ln.weight"], spec["visual_goal.ln.bias"] = (G,), (G,)
    lin("language_goal.mlp.1", d.goal_hidden_lang, d.lang_in), lin("language_goal.mlp.3", d.goal_hidden_lang, d.goal_hidden_lang)
    lin("language_goal.mlp.5", G, d.goal_hidden_lang)
    spec["language_goal.ln.weight"], spec["language_goal.ln.bias"] = (G,), (G,)
    dec_in = d.plan_features + (D - d.percep_lo) + G
    rnn("action_decoder.rnn", dec_in, d.dec_hidden, 2, gates=3 if rnn_model == "gru_decoder" else 1)
    n_out = d.n_dims * d.n_mix
    lin("action_decoder.mean_fc", n_out, d.dec_hidden), lin("action_decoder.log_scale_fc", n_out, d.dec_hidden), lin("action_decoder.prob_fc", n_out, d.dec_hidden)
    if model != "mcil":
        lin("action_decoder.gripper_fc", 2, d.dec_hidden)
        lin("proj_vis_lang.mlp_im.0", d.clip_hidden, d.fc_hidden), lin("proj_vis_lang.mlp_im.2", d.clip_out, d.clip_hidden)
        lin("proj_vis_lang.mlp_lang.0", d.clip_hidden, G), lin("proj_vis_lang.mlp_lang.2", d.clip_out, d.clip_hidden)
    if d.bc_z:
        lin("bc_z_lang_decoder.mlp.0", d.aux_hidden, d.fc_hidden), lin("bc_z_lang_decoder.mlp.2", d.lang_in, d.aux_hidden)
    if d.mia:
        lin("mia_lang_discriminator.mlp.0", d.aux_hidden, 2 * d.clip_out), lin("mia_lang_discriminator.mlp.3", 1, d.aux_hidden)
    return spec
