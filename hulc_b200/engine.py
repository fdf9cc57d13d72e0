"""The HULC training step on the sm_100a kernels: forward, hand-written backward, Adam — no torch autograd, no torch
math on the hot path (torch supplies device memory, streams and views only).

Mirrors `Hulc.training_step` + `lmp_train` (reference hulc/models/hulc.py:390-537, 254-299), `GCBC.training_step`
(hulc/models/gcbc.py:50-181) and the MCIL configuration (conf/model/mcil.yaml).  Both modalities of the batch run
through the shared encoders / posterior / decoder as ONE batch (they use the same weights); goal encoders, the
losses and the CLIP term stay per modality, exactly as the reference averages them (hulc.py:464-491).

Layouts: frames and posterior tokens are batch-first rows (b*S + s); everything recurrent is time-major rows
(t*B + b) so each step's hidden state is one contiguous matrix.
"""
from __future__ import annotations

import math
import contextlib
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .ops import Drop, NO_DROP, gemm, colsum
from .spec import ModelDims, param_spec

RELU, TANH, GATE_TANH = 1, 2, 4
_TRACE_CASTS = os.environ.get("HULC_B200_TRACE_CASTS", "0") == "1"
_POISON = bool(int(os.environ.get("HULC_B200_POISON", "0")))
# HULC_B200_NVTX=1: one NVTX range per block of the step (encoders, goal, prior, posterior, plan, decoder, losses, and their backward
# counterparts), so a timeline (nsys / ncu --nvtx) reads in the reference's vocabulary
_NVTX = bool(int(os.environ.get("HULC_B200_NVTX", "0")))
_nvtx_open = [False]


def _mark(name: Optional[str]):
    """Close the current NVTX range and open `name` (None: just close)."""
    if not _NVTX:
        return
    if _nvtx_open[0]:
        torch.cuda.nvtx.range_pop()
    _nvtx_open[0] = name is not None
    if name is not None:
        torch.cuda.nvtx.range_push(name)

# parameters that must sit next to each other in the flat buffer so one GEMM covers them (decoder heads:
# logit_probs | means | log_scales | gripper — the row layout hulc_logistic_loss expects).  The group is padded with zero
# rows to a multiple of 4 (182 -> 184) so that the fused [rows, heads] activations have 16-byte-aligned rows and all three
# head products run on the tensor cores; the pad rows belong to no state_dict key and stay zero under Adam (zero gradient).
_HEAD_ORDER = ("prob_fc", "mean_fc", "log_scale_fc", "gripper_fc")


class ParamStore:
    """Flat fp32 parameter / gradient / Adam-moment buffers with named views under the reference's state_dict keys."""

    def __init__(self, spec: Dict[str, tuple], device):
        keys = [k for k in spec if not k.startswith("action_decoder.") or k.split(".")[1] not in _HEAD_ORDER]
        head_w = [f"action_decoder.{h}.weight" for h in _HEAD_ORDER if f"action_decoder.{h}.weight" in spec]
        head_b = [f"action_decoder.{h}.bias" for h in _HEAD_ORDER if f"action_decoder.{h}.bias" in spec]
        groups = [[k] for k in keys] + [head_w, head_b]
        self.offsets: Dict[str, Tuple[int, tuple]] = {}
        off = 0
        self.n_heads = sum(spec[k][0] for k in head_w)
        self.n_heads_padded = (self.n_heads + 3) // 4 * 4
        for grp in groups:
            off = (off + 7) // 8 * 8  # 32-byte alignment per group: its bf16 copy (same element offsets) is then 16-byte aligned, as TMA needs
            for k in grp:
                n = int(math.prod(spec[k])) if len(spec[k]) else 1
                self.offsets[k] = (off, tuple(spec[k]))
                off += n
            if grp is head_w and head_w:
                off += (self.n_heads_padded - self.n_heads) * spec[head_w[0]][1]
        self.numel = (off + 7) // 8 * 8
        self.device = torch.device(device)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.keys = list(spec)
        self.step_count = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=device)  # device copy of step_count (read by the Adam kernel)
        self._head_w, self._head_b = head_w, head_b
        self.flat_bf16: Optional[torch.Tensor] = None  # bf16 operand copy of `flat` (bf16 path), refreshed by the Adam kernel
        self.bf16_stale = True
        self.rebuild_views()

    def enable_bf16(self):
        """Keep a bf16 copy of the parameters next to the fp32 master weights (what torch.autocast re-creates on every use)."""
        if self.flat_bf16 is None or self.flat_bf16.device != self.flat.device:
            self.flat_bf16 = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.flat.device)
            self.bf16_stale = True

    def refresh_bf16(self):
        if self.flat_bf16 is not None and self.bf16_stale:
            ops.cast_bf16(self.flat, self.flat_bf16)
            self.bf16_stale = False

    def rebuild_views(self):
        """(Re)create the named views after the flat buffers were allocated or moved to another device."""
        self.p = {k: self._view(self.flat, k) for k in self.keys}
        self.g = {k: self._view(self.grad, k) for k in self.keys}
        if self._head_w:  # fused views over the decoder-head group
            o, (_, n_in) = self.offsets[self._head_w[0]]
            rows = self.n_heads_padded
            self.heads_w = self.flat[o : o + rows * n_in].view(rows, n_in)
            self.heads_gw = self.grad[o : o + rows * n_in].view(rows, n_in)
            ob, _ = self.offsets[self._head_b[0]]
            self.heads_b, self.heads_gb = self.flat[ob : ob + rows], self.grad[ob : ob + rows]

    def _view(self, flat, k):
        off, shape = self.offsets[k]
        n = int(math.prod(shape)) if len(shape) else 1
        return flat[off : off + n].view(shape)

    @torch.no_grad()
    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.keys if k not in sd]
        if missing and strict:
            raise KeyError(f"missing parameters: {missing[:5]}")
        for k in self.keys:
            if k in sd:
                self.p[k].copy_(sd[k].to(torch.float32))
        self.bf16_stale = True

    def state_dict(self):
        return {k: self.p[k].detach().clone() for k in self.keys}

    def zero_grad(self):
        self.grad.zero_()

    def adam_step(self, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        """torch.optim.Adam update of the whole flat buffer in one launch.  The step count lives on the device so that the
        launch can be captured in (and replayed from) a CUDA graph."""
        self.step_count += 1
        self.step_dev.add_(1)
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, lr=lr, beta1=betas[0], beta2=betas[1], eps=eps,
                      step=0, step_dev=self.step_dev, grad_scale=grad_scale, p_bf16=self.flat_bf16)


class StepGraph:
    """A captured training step (see HulcEngine.capture)."""

    def __init__(self, engine, graph, out, optimizer, launches, namespace=None):
        self.engine, self.graph, self.out, self.optimizer = engine, graph, out, optimizer
        self.launches = launches  # kernels of libhulc_b200.so inside one replay
        self.namespace = namespace  # the engine's buffer set this graph's pointers refer to

    def replay(self) -> Dict[str, torch.Tensor]:
        self.graph.replay()
        if self.optimizer:
            self.engine.ps.step_count += 1
        return self.out


class HulcEngine:
    """Owns the parameters and runs fused forward+backward steps.  `model` in {"hulc", "gcbc", "mcil"}."""

    def __init__(self, model="hulc", rnn_model="rnn_decoder", max_window=32, device="cuda", dropout_p=0.1, kl_beta=0.01,
                 kl_balancing_mix=0.8, clip_beta=3.0, gripper_alpha=1.0, nhead=8, nlayers=2, lr=2e-4, precision="tf32", dims: Optional[ModelDims] = None):
        """precision: "tf32" (default) runs the convolutions and the large backward GEMMs on the tensor cores with tf32
        operands and the large forward GEMMs as 3xTF32 (fp32-level accuracy); "fp32" keeps every product on the exact-fp32
        CUDA-core kernels (also the only mode of the host-emulated build used by the CPU tests).  `dims` (hulc_b200.spec.ModelDims,
        normally read from the Hydra config tree by `dims_from_configs`) overrides the individual size arguments."""
        if dims is None:
            dims = ModelDims.shipped(model, rnn_model, max_window, nhead=nhead, nlayers=nlayers, gripper_alpha=float(gripper_alpha),
                                     **({} if model == "mcil" else {"dropout_p": float(dropout_p)}))
        model, rnn_model = dims.model, dims.rnn_model
        assert model in ("hulc", "gcbc", "mcil") and rnn_model in ("rnn_decoder", "gru_decoder")
        assert precision in ("tf32", "fp32", "bf16")
        self.dims = dims
        self.precision = precision
        self.bf16 = precision == "bf16"        # bf16 tensor-core operands for every Linear product (and bf16 activations between the conv layers)
        self.tc = precision in ("tf32", "bf16")
        # bf16 activations between the conv layers (HULC_B200_BF16_CONV=0 keeps the fp32 / tf32 conv stack under the bf16 Linear products)
        self.bf16_conv = self.bf16 and os.environ.get("HULC_B200_BF16_CONV", "1") != "0"
        self.persistent_rnn = os.environ.get("HULC_B200_PERSISTENT_RNN", "1") != "0"  # whole recurrence in one launch (csrc/rnn_tc.cu)
        self.model, self.rnn_model = model, rnn_model
        self.device = torch.device(device)
        self.spec = param_spec(dims=dims)
        self.ps = ParamStore(self.spec, device)
        if self.bf16:
            self.ps.enable_bf16()
        self._twins: Dict[tuple, list] = {}     # (ptr, shape, strides) of an fp32 tensor -> [bf16 twin, generation it is valid for, byte range]
        self._twin_gen = 0
        self._bias_jobs: list = []
        self._side, self._side_dirty = None, False
        self.overlap_wgrad = os.environ.get("HULC_B200_OVERLAP_WGRAD", "1") != "0"  # Linear weight gradients on a side stream (gemm_wgrad)
        self.augment_pad: Dict[str, int] = {}   # camera -> RandomShiftsAug pad (set_augmentation); empty: no device-side augmentation
        self._aug_ctx = None
        self._bf16_only: set = set()            # keys whose fp32 storage was never written this step (the producer emitted bf16 only)
        self.dropout_p = float(dims.dropout_p) if model != "mcil" else 0.0
        self.kl_beta, self.kl_alpha, self.clip_beta, self.gripper_alpha = float(kl_beta), float(kl_balancing_mix), float(clip_beta), float(dims.gripper_alpha)
        self.bc_z_beta, self.mia_beta = 1.0, 1.0  # conf/loss/default.yaml:4-5 (the module overwrites them from its constructor arguments)
        self.nhead, self.nlayers, self.lr = dims.nhead, dims.nlayers, lr
        self.discrete = dims.discrete
        self.plan_features = dims.plan_features
        self.percep_lo = dims.percep_lo  # decoder sees emb[..., 64:128] (gripper features) unless MCIL
        self.n_dims = dims.n_dims
        self.n_mix = dims.n_mix
        self.num_classes = dims.num_classes
        self.H = dims.dec_hidden
        self.H_prior = dims.prior_hidden
        self.H_bi = 2048  # hidden size of the MCIL BiRNN posterior (plan_recognition_net.py:27-34)
        self.gates = 3 if rnn_model == "gru_decoder" else 1
        self._bufs: Dict[str, torch.Tensor] = {}
        self._train_root = self._bufs  # the dictionary in place outside any namespace
        self._buf_namespaces: Dict[object, Dict[str, torch.Tensor]] = {}
        self._infer_state = None   # persistent rollout tensors (goal, plan, hidden, const); allocated by the first infer_plan
        self._infer_planned = False
        self._infer_graph = None
        self._step_shapes: Dict[str, tuple] = {}
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=device)
        # per-step Philox seed, device resident: advanced on the device so a captured step draws fresh randomness per replay
        self.rng_dev = torch.zeros(1, dtype=torch.int64, device=device)

    # ------------------------------------------------------------------------------------------------------------------
    def buf(self, name, *shape, zero=False):
        t = self._bufs.get(name)
        seen = self._step_shapes.setdefault(name, tuple(shape))
        if seen != tuple(shape):
            raise RuntimeError(f"buffer name {name!r} requested with two shapes in one step: {seen} and {tuple(shape)}")
        if t is None or tuple(t.shape) != tuple(shape):
            t = (torch.zeros if zero else torch.empty)(*shape, dtype=torch.float32, device=self.device)
            if _POISON and not zero:
                t.fill_(float("nan"))  # debugging aid: any read of a never-written element surfaces as NaN
            self._bufs[name] = t
        return t

    def hbuf(self, name, *shape):
        """persistent bf16 buffer (bf16 path: the channels-last activations of the conv stack and their gradients)"""
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != torch.bfloat16:
            t = self._bufs[name] = torch.zeros(*shape, dtype=torch.bfloat16, device=self.device)
        return t

    def ibuf(self, name, *shape):
        """persistent int32 buffer (ReLU sign masks: one word per pixel and 32 channels)"""
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self._bufs[name] = torch.zeros(*shape, dtype=torch.int32, device=self.device)
        return t

    @contextlib.contextmanager
    def _buffers(self, namespace):
        """Run a block on its own set of persistent buffers.  The training step's buffers are the static inputs / outputs of its captured CUDA
        graphs: validation batches or single-frame inference re-using the same names with other shapes would re-allocate them under a graph."""
        saved = self._bufs
        self._bufs = self._buf_namespaces.setdefault(namespace, {})
        try:
            yield
        finally:
            self._bufs = saved

    def load_state_dict(self, sd, strict=True):
        self.ps.load_state_dict(sd, strict)

    def state_dict(self):
        return self.ps.state_dict()

    # ------------------------------------------------------------------------------------------------------------------
    # GEMM dispatch: which products go to the tensor cores
    # ------------------------------------------------------------------------------------------------------------------
    def _tc_mode(self, M, N, K, role):
        """0 = exact fp32 on CUDA cores; 1 = tf32 (backward products: they only feed gradients); 3 = 3xTF32 (forward
        products, whose outputs are held to the fp32 parity tolerance)."""
        if not self.tc or N < 32 or K < 32:
            return 0
        # Both flavours run on the TMA-fed kernel (csrc/gemm_bf16_tc.cu, fp32 operands): ~5-7 us for the smallest product inside a graph, which the
        # CUDA-core kernel (~10 us) only beats below these sizes.  (With the cp.async kernels of round 1 the thresholds were K >= 64 and M >= 256.)
        if role == "bwd":
            return 1
        return 3

    def gemm_fwd(self, A, B, C=None, out="f32", **kw):
        """out (bf16 mode only): "f32" writes C; "bf16" writes only C's bf16 twin (C stays a handle: for activations that are consumed by
        further products and ReLU gates only); "both" writes the two."""
        if self.bf16:
            return self._gemm16(A, B, C, out=out, **kw)
        M, K = (A.shape[1], A.shape[0]) if kw.get("transA") else A.shape
        N = B.shape[0] if kw.get("transB") else B.shape[1]
        return gemm(A, B, C, tc=self._tc_mode(M, N, K, "fwd"), **kw)

    def gemm_bwd(self, A, B, C=None, out="f32", **kw):
        if self.bf16:
            return self._gemm16(A, B, C, out=out, **kw)
        M, K = (A.shape[1], A.shape[0]) if kw.get("transA") else A.shape
        N = B.shape[0] if kw.get("transB") else B.shape[1]
        return gemm(A, B, C, tc=self._tc_mode(M, N, K, "bwd"), **kw)

    def gemm_wgrad(self, dy, x, gw, beta=0.0):
        """gw = beta * gw + dy^T x — the weight gradient of a Linear layer.  Nothing downstream in the backward chain depends on it, so in the
        tensor-core modes it is launched on a side stream (a fork inside the captured graph) and runs BESIDE the data-gradient chain, whose
        products are small and latency-bound; the operands (bf16 twins) are prepared on the main stream first.  `_join_side` re-joins."""
        M, K, N = dy.shape[1], dy.shape[0], x.shape[1]
        side = self._side_stream()
        if side is None:
            return self.gemm_bwd(dy, x, gw, transA=True, beta=beta)
        if self.bf16:
            A16, B16 = self._tw(dy), self._tw(x)
            if not ops.gemm_bf16_ok(A16, B16):
                return self.gemm_bwd(dy, x, gw, transA=True, beta=beta)
        else:
            mode = self._tc_mode(M, N, K, "bwd")
            if not mode or not ops._tc_ok(dy, x, M, N, K, True, False):  # the CUDA-core kernel shares the split-K workspace with the main stream
                return self.gemm_bwd(dy, x, gw, transA=True, beta=beta)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            if self.bf16:
                ops.gemm_bf16(A16, B16, gw, None, transA=True, beta=beta)
            else:
                gemm(dy, x, gw, transA=True, beta=beta, tc=mode)
        self._side_dirty = True
        return gw

    def _side_stream(self):
        if not self.tc or self.device.type != "cuda" or not self.overlap_wgrad:
            return None
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()  # ("cuda" has no index: compare resolved ones)
        if self._side is None or self._side.device.index != idx:
            self._join_side()
            self._side = torch.cuda.Stream(device=torch.device("cuda", idx))
        return self._side

    def _join_side(self):
        """The main stream waits for the weight-gradient products on the side stream (end of a backward pass / before gradients are consumed)."""
        if self._side_dirty:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_dirty = False

    def gemm_rec(self, A, B, C, role, **kw):
        """The per-step products of the recurrences that do not run in the persistent kernel (GRU, windows above 32 steps)."""
        if self.bf16:
            return self._gemm16(A, B, C, **kw)
        return gemm(A, B, C, tc=(3 if role == "fwd" else 1) if self.tc else 0, **kw)

    # ---- bf16 path: operand twins ---------------------------------------------------------------------------------------
    # Every product of the bf16 path reads bf16 operands (hulc_gemm_bf16).  Parameters have a persistent bf16 copy (ParamStore.flat_bf16,
    # refreshed by the Adam kernel).  An fp32 activation gets a bf16 "twin": written by the producing product's epilogue (out="bf16" /
    # "both") or by one cast launch the first time a product consumes it in a step; the twin is remembered for the rest of the step and
    # dropped when a product writes the memory again.
    @staticmethod
    def _key(t):
        return (t.data_ptr(), tuple(t.shape), tuple(t.stride()))

    @staticmethod
    def _span(t):
        return (t.data_ptr(), t.data_ptr() + 4 * (sum((n - 1) * st for n, st in zip(t.shape, t.stride())) + 1))

    def _twin_buffer(self, t, key):
        ent = self._twins.get(key)
        if ent is None:
            if t.dim() == 2:
                ld = (t.shape[1] + 7) // 8 * 8  # 16-byte aligned rows (TMA)
                tw = torch.zeros(t.shape[0], ld, dtype=torch.bfloat16, device=t.device)[:, : t.shape[1]]
            else:
                tw = torch.zeros(t.shape, dtype=torch.bfloat16, device=t.device)
            ent = self._twins[key] = [tw, -1, self._span(t)]
        return ent

    def _tw(self, t, produce=False):
        """bf16 twin of the fp32 tensor `t` (a parameter view or an activation).  produce=True: the caller is about to write it."""
        if t.dtype == torch.bfloat16:
            return t
        ps = self.ps
        off = t.data_ptr() - ps.flat.data_ptr()
        if 0 <= off < 4 * ps.numel:
            return torch.as_strided(ps.flat_bf16, t.shape, t.stride(), off // 4)
        key = self._key(t)
        ent = self._twin_buffer(t, key)
        if produce:
            ent[1] = self._twin_gen
        elif ent[1] != self._twin_gen:
            assert key not in self._bf16_only
            if _TRACE_CASTS:  # development aid: which activations still reach a product as fp32 (one cast launch each)
                name = next((n for n, b in self._bufs.items() if b.data_ptr() <= t.data_ptr() < b.data_ptr() + b.numel() * b.element_size()), "?")
                print(f"[cast] {name} {tuple(t.shape)}")
            ops.cast_bf16(t, ent[0])
            ent[1] = self._twin_gen
        return ent[0]

    def _written(self, t, keep=None):
        """fp32 memory of `t` is (re)written: twins of anything overlapping it are stale."""
        lo, hi = self._span(t)
        for k, ent in self._twins.items():
            if ent[1] == self._twin_gen and k != keep and ent[2][0] < hi and lo < ent[2][1]:
                ent[1] = -1
                self._bf16_only.discard(k)

    def _new_generation(self):
        self._twin_gen += 1
        self._bf16_only.clear()
        if self.bf16:
            self.ps.enable_bf16()  # (no-op once allocated; a replaced / moved ParamStore gets its copy here)
            self.ps.refresh_bf16()

    def _gemm16(self, A, B, C, *, out="f32", gate=None, **kw):
        if C.dtype == torch.bfloat16:  # a bf16 activation buffer (conv stack): the result goes there directly
            A16, B16 = self._tw(A), self._tw(B)
            if gate is not None and gate.dtype != torch.bfloat16 and self._key(gate) in self._bf16_only:
                gate = self._tw(gate)
            ops.gemm_bf16(A16, B16, None, C, gate=gate, **kw)
            return C
        A16, B16 = self._tw(A), self._tw(B)
        if not ops.gemm_bf16_ok(A16, B16):
            raise ops._lib.HulcError(f"bf16 product with operands the TMA cannot address: A {tuple(A.shape)} strides {A.stride()}, B {tuple(B.shape)} strides {B.stride()}")
        key = self._key(C)
        self._written(C, keep=key if out != "f32" else None)
        Cb = self._tw(C, produce=True) if out != "f32" else None
        if out == "f32":
            ent = self._twins.get(key)
            if ent is not None:
                ent[1] = -1
            self._bf16_only.discard(key)
        elif out == "bf16":
            if kw.get("beta", 0.0) != 0.0:
                raise ValueError("beta accumulates into the fp32 result")
            self._bf16_only.add(key)
        else:
            self._bf16_only.discard(key)
        if gate is not None and self._key(gate) in self._bf16_only:
            gate = self._tw(gate)
        ops.gemm_bf16(A16, B16, None if out == "bf16" else C, Cb, gate=gate, **kw)
        return C

    # ------------------------------------------------------------------------------------------------------------------
    # small composites
    # ------------------------------------------------------------------------------------------------------------------
    def bias_grad(self, dy, gb):
        """gb += column sums of dy (the bias gradient of a Linear / conv layer).  Deferred: all of a backward pass's bias gradients are summed
        by ONE launch at its end (`_flush_bias_grads`; every dy lives in its own persistent buffer until then)."""
        if dy.shape[0] > 16384:  # (conv maps of the fp32-activation path: ~10^6 rows) — the row-split, ticketed reduction, right away
            colsum(dy, gb, beta=1.0)
            return
        self._bias_jobs.append((dy, gb, 1.0))

    def _flush_bias_grads(self):
        jobs, self._bias_jobs = self._bias_jobs, []
        ops.colsum_multi(jobs)

    # ------------------------------------------------------------------------------------------------------------------
    # BC-Z / MIA auxiliary heads (hulc/models/hulc.py:567-648; bc_z_lang_decoder.py:5-20, mia_lang_discriminator.py:5-21) — ablation
    # configs, language modality only.  A few tiny products: they run on the exact-fp32 kernels in every precision mode.
    # ------------------------------------------------------------------------------------------------------------------
    def _aux_heads_fwd(self, d, mask, sf, im2, tx2, d_im2, d_tx2, Bm):
        P = self.ps.p
        active = True
        if mask is not None:  # (a host sync: the module keeps CUDA graphs off while these heads are on)
            if not bool(mask.all()):
                if bool(mask.any()):
                    raise NotImplementedError("use_for_aux_lang_loss keeps only part of the batch: the reference's BC-Z loss indexes with the mask twice "
                                              "(hulc.py:592-594) and fails there; MIA would need the kept rows compacted")
                active = False  # hulc.py:583-590 / 626-633: dummy forward, loss * 0 -> no loss, no gradient
        losses = self.buf("aux.losses", 2)
        losses.zero_()
        ctx = {"losses": losses, "active": active}
        if not active:
            return ctx
        Ha = self.dims.aux_hidden
        if self.dims.bc_z:
            lang = d["lang"].contiguous()
            h = gemm(sf, P["bc_z_lang_decoder.mlp.0.weight"], self.buf("bcz.h", Bm, Ha), transB=True, bias=P["bc_z_lang_decoder.mlp.0.bias"], act=RELU)
            pred = gemm(h, P["bc_z_lang_decoder.mlp.2.weight"], self.buf("bcz.pred", Bm, lang.shape[1]), transB=True, bias=P["bc_z_lang_decoder.mlp.2.bias"])
            dpred = self.buf("bcz.dpred", *pred.shape)
            ops.cosine_loss(pred, lang, dpred, losses[0:1], grad_scale=self.bc_z_beta)
            ctx.update(bcz_h=h, bcz_dpred=dpred)
        if self.dims.mia:
            Dc = im2.shape[1]
            x = self.buf("mia.x", 2 * Bm, 2 * Dc)  # rows [0, Bm): (im_i, tx_i); rows [Bm, 2 Bm): (im_i, tx_{i-1}) = torch.roll(tx, 1, 0)
            ops.strided_copy(x[:Bm, :Dc], im2), ops.strided_copy(x[:Bm, Dc:], tx2), ops.strided_copy(x[Bm:, :Dc], im2)
            ops.strided_copy(x[Bm : Bm + 1, Dc:], tx2[Bm - 1 : Bm])
            if Bm > 1:
                ops.strided_copy(x[Bm + 1 :, Dc:], tx2[: Bm - 1])
            hm = gemm(x, P["mia_lang_discriminator.mlp.0.weight"], self.buf("mia.h", 2 * Bm, Ha), transB=True, bias=P["mia_lang_discriminator.mlp.0.bias"], act=RELU)
            logit = gemm(hm, P["mia_lang_discriminator.mlp.3.weight"], self.buf("mia.logit", 2 * Bm, 1), transB=True, bias=P["mia_lang_discriminator.mlp.3.bias"])
            dlogit = self.buf("mia.dlogit", 2 * Bm, 1)
            ops.bce_logits_loss(logit.view(-1), dlogit.view(-1), losses[1:2], Bm, Bm, grad_scale=self.mia_beta)
            ctx.update(mia_x=x, mia_h=hm, mia_dlogit=dlogit)
        return ctx

    def _aux_heads_bwd(self, ctx, sf, dseq, d_im2, d_tx2, Bm, *, mia_only):
        if not ctx["active"]:
            return
        P, G = self.ps.p, self.ps.g

        def lin_bwd(name, x, dy, dx, gate=None, dx_beta=0.0):
            gemm(dy, x, G[name + ".weight"], transA=True, beta=1.0)
            self.bias_grad(dy, G[name + ".bias"])
            return gemm(dy, P[name + ".weight"], dx, beta=dx_beta, gate=gate)

        if mia_only:
            if not self.dims.mia:
                return
            x, hm, dlogit = ctx["mia_x"], ctx["mia_h"], ctx["mia_dlogit"]
            Dc = d_im2.shape[1]
            dhm = lin_bwd("mia_lang_discriminator.mlp.3", hm, dlogit, self.buf("mia.dh", *hm.shape), gate=hm)
            dx = lin_bwd("mia_lang_discriminator.mlp.0", x, dhm, self.buf("mia.dx", *x.shape))
            ops.strided_copy(d_im2, dx[:Bm, :Dc], accumulate=True), ops.strided_copy(d_im2, dx[Bm:, :Dc], accumulate=True)
            ops.strided_copy(d_tx2, dx[:Bm, Dc:], accumulate=True)
            ops.strided_copy(d_tx2[Bm - 1 : Bm], dx[Bm : Bm + 1, Dc:], accumulate=True)
            if Bm > 1:
                ops.strided_copy(d_tx2[: Bm - 1], dx[Bm + 1 :, Dc:], accumulate=True)
        elif self.dims.bc_z:
            h, dpred = ctx["bcz_h"], ctx["bcz_dpred"]
            dh = lin_bwd("bc_z_lang_decoder.mlp.2", h, dpred, self.buf("bcz.dh", *h.shape), gate=h)
            lin_bwd("bc_z_lang_decoder.mlp.0", sf, dh, dseq, dx_beta=1.0)

    def _linear_bwd(self, name, x, dy, dx=None, *, gate=None, dx_beta=0.0, need_dx=True, act=0, addend=None, drop=NO_DROP):
        """Gradients of y = x W^T + b: accumulates dW, db; returns dx = dy W (optionally gated by the producer's ReLU)."""
        P, G = self.ps.p, self.ps.g
        self.gemm_wgrad(dy, x, G[name + ".weight"], beta=1.0)
        self.bias_grad(dy, G[name + ".bias"])
        if not need_dx:
            return None
        # bf16 path: dx usually is the dy of the next layer down (a product operand again): let the epilogue write its bf16 twin too
        return self.gemm_bwd(dy, P[name + ".weight"], dx, beta=dx_beta, gate=gate, act=act, addend=addend, drop=drop, out="both")

    def _mlp_ln_fwd(self, tag, x, names, ln, out, w0=None):
        """x -> [Linear+ReLU]* -> Linear -> LayerNorm written into `out` (a 2-D view).  Saves activations under `tag`.
        `w0` replaces the first layer's weight (a column-permuted copy, see _encoder_fwd_tc)."""
        P = self.ps.p
        acts = [x]
        for i, n in enumerate(names):
            last = i == len(names) - 1
            y = self.buf(f"{tag}.mlp{i}", x.shape[0], P[n + ".weight"].shape[0])
            w = w0 if (i == 0 and w0 is not None) else P[n + ".weight"]
            self.gemm_fwd(acts[-1], w, y, transB=True, bias=P[n + ".bias"], act=0 if last else RELU, out="f32" if last else "bf16")
            acts.append(y)
        stats = self.buf(f"{tag}.mlp_stats", x.shape[0], 2)
        ops.layernorm_fwd(acts[-1], P[ln + ".weight"], P[ln + ".bias"], out, stats)
        return acts, stats

    def _mlp_ln_bwd(self, tag, acts, stats, names, ln, dout, dx=None, dx_beta=0.0, need_dx=True):
        P, G = self.ps.p, self.ps.g
        d = self.buf(f"{tag}.mlp_dz", *acts[-1].shape)
        ops.layernorm_bwd(dout, acts[-1], stats, P[ln + ".weight"], G[ln + ".weight"], G[ln + ".bias"], dz=d)
        for i in reversed(range(len(names))):
            first = i == 0
            if first:
                return self._linear_bwd(names[i], acts[i], d, dx, dx_beta=dx_beta, need_dx=need_dx)
            nd = self.buf(f"{tag}.mlp_d{i}", *acts[i].shape)
            self._linear_bwd(names[i], acts[i], d, nd, gate=acts[i])
            d = nd

    # ------------------------------------------------------------------------------------------------------------------
    # perceptual encoders (concat_encoders.py:59-109, vision_network.py:55-65, vision_network_gripper.py:49-56)
    # ------------------------------------------------------------------------------------------------------------------
    _CONVS = ((0, 4), (2, 2), (4, 1))

    def set_augmentation(self, static_pad: int = 10, gripper_pad: int = 4):
        """Apply the reference's training-time RandomShiftsAug (conf/datamodule/transforms/rand_shift.yaml:2-22: pad 10 / 4) on the device to
        uint8 frames of TRAINING steps, fused with their scale + normalise pass.  0 disables it for a camera.  Validation / inference never shift."""
        self.augment_pad = {"static": int(static_pad), "gripper": int(gripper_pad)}

    def _normalised_frames(self, which, frames: List[torch.Tensor]) -> List[torch.Tensor]:
        """uint8 frames (as read from disk) are scaled / normalised on the device into a persistent fp32 buffer — the
        deterministic part of the reference's image transforms (conf/datamodule/transforms/rand_shift.yaml:3-22) and, in training steps
        with `set_augmentation`, its RandomShiftsAug (one random whole-pixel shift per frame); fp32 frames (the reference's batch contract)
        pass through untouched."""
        out = []
        aug = self._aug_ctx
        pad = self.augment_pad.get(which, 0) if aug is not None else 0
        n0 = 0
        for i, f in enumerate(frames):
            if f.dtype == torch.uint8:
                dst = self.buf(f"{which}.frames{i}", *f.shape)
                if pad > 0:
                    sh = aug["shifts"].get(which) if aug["shifts"] is not None else None
                    ops.frames_u8_shift_to_f32(f.contiguous(), dst, pad, shifts=None if sh is None else sh[n0 : n0 + f.shape[0]].contiguous(), seed=aug["seed"],
                                               site=400 + 2 * i + (which == "gripper"))
                    f = dst
                else:
                    f = ops.frames_u8_to_f32(f.contiguous(), dst)
            n0 += f.shape[0]
            out.append(f)
        return out

    def _encoder_fwd(self, which, frames: List[torch.Tensor], emb):
        frames = self._normalised_frames(which, frames)
        # the tensor-core first layer reads the NCHW rows in 16-byte pieces: other widths take the exact-fp32 kernels
        if self.bf16_conv and frames[0].shape[-1] % 4 == 0:
            ctx = self._encoder_fwd_bf16(which, frames, emb)
            ctx["bf16"] = True
            return ctx
        if self.tc and frames[0].shape[-1] % 4 == 0:
            ctx = self._encoder_fwd_tc(which, frames, emb)
            ctx["tc"] = True
            return ctx
        P = self.ps.p
        pre = f"perceptual_encoder.rgb_{which}_encoder"
        N = sum(f.shape[0] for f in frames)
        hw = frames[0].shape[-1]
        sizes = [hw]
        for (_, s), k in zip(self._CONVS, (8, 4, 3)):
            sizes.append((sizes[-1] - k) // s + 1)
        a1 = self.buf(f"{which}.a1", N, 32, sizes[1], sizes[1])
        n0 = 0
        for f in frames:  # conv1 per modality tensor, written into the shared activation buffer
            ops.conv2d_fwd(f, P[f"{pre}.conv_model.0.weight"], P[f"{pre}.conv_model.0.bias"], 4, a1[n0 : n0 + f.shape[0]])
            n0 += f.shape[0]
        a2 = ops.conv2d_fwd(a1, P[f"{pre}.conv_model.2.weight"], P[f"{pre}.conv_model.2.bias"], 2, self.buf(f"{which}.a2", N, 64, sizes[2], sizes[2]))
        a3 = ops.conv2d_fwd(a2, P[f"{pre}.conv_model.4.weight"], P[f"{pre}.conv_model.4.bias"], 1, self.buf(f"{which}.a3", N, 64, sizes[3], sizes[3]))
        if which == "static":
            feat = ops.spatial_softmax_fwd(a3, self.buf("static.ss", N, 128), temperature=self.dims.spatial_softmax_temp)
            names = [f"{pre}.fc1.0", f"{pre}.fc2"]
            out = emb[:, 0:64]
        else:
            feat = a3.view(N, -1)
            names = [f"{pre}.conv_model.7", f"{pre}.fc1.0", f"{pre}.fc2"]
            out = emb[:, 64:128]
        acts, stats = self._mlp_ln_fwd(which, feat, names, f"{pre}.ln", out)
        return dict(frames=frames, a1=a1, a2=a2, a3=a3, acts=acts, stats=stats, names=names, pre=pre)

    def _encoder_bwd(self, which, ctx, demb):
        if ctx.get("bf16"):
            return self._encoder_bwd_bf16(which, ctx, demb)
        if ctx.get("tc"):
            return self._encoder_bwd_tc(which, ctx, demb)
        P, G = self.ps.p, self.ps.g
        pre, a1, a2, a3 = ctx["pre"], ctx["a1"], ctx["a2"], ctx["a3"]
        dout = demb[:, 0:64] if which == "static" else demb[:, 64:128]
        N = a3.shape[0]
        if which == "static":
            dss = self._mlp_ln_bwd(which, ctx["acts"], ctx["stats"], ctx["names"], f"{pre}.ln", dout, self.buf("static.dss", N, 128))
            da3 = ops.spatial_softmax_bwd(a3, dss, self.buf("static.da3", *a3.shape), temperature=self.dims.spatial_softmax_temp, relu_gate=True)
        else:
            # dgrad of the flatten-FC, gated by conv3's ReLU
            acts, names = ctx["acts"], ctx["names"]
            d = self.buf(f"{which}.mlp_dz", *acts[-1].shape)
            ops.layernorm_bwd(dout, acts[-1], ctx["stats"], P[f"{pre}.ln.weight"], G[f"{pre}.ln.weight"], G[f"{pre}.ln.bias"], dz=d)
            for i in (2, 1):
                nd = self.buf(f"{which}.mlp_d{i}", *acts[i].shape)
                self._linear_bwd(names[i], acts[i], d, nd, gate=acts[i])
                d = nd
            da3 = self.buf("gripper.da3", *a3.shape)
            self._linear_bwd(names[0], acts[0], d, da3.view(N, -1), gate=acts[0])
        ops.conv2d_wgrad(a2, da3, G[f"{pre}.conv_model.4.weight"], 1, beta=1.0)
        ops.nchw_channel_sum(da3, G[f"{pre}.conv_model.4.bias"])
        da2 = ops.conv2d_dgrad(da3, P[f"{pre}.conv_model.4.weight"], a2.shape, 1, gate=a2, dx=self.buf(f"{which}.da2", *a2.shape))
        ops.conv2d_wgrad(a1, da2, G[f"{pre}.conv_model.2.weight"], 2, beta=1.0)
        ops.nchw_channel_sum(da2, G[f"{pre}.conv_model.2.bias"])
        da1 = ops.conv2d_dgrad(da2, P[f"{pre}.conv_model.2.weight"], a1.shape, 2, gate=a1, dx=self.buf(f"{which}.da1", *a1.shape))
        ops.nchw_channel_sum(da1, G[f"{pre}.conv_model.0.bias"])
        n0 = 0
        for f in ctx["frames"]:  # the images need no gradient: conv1 has a weight gradient only
            ops.conv2d_wgrad(f, da1[n0 : n0 + f.shape[0]], G[f"{pre}.conv_model.0.weight"], 4, beta=1.0)
            n0 += f.shape[0]

    # tensor-core variant: channels-last activations, tcgen05 implicit-GEMM convolutions (conv_tc.cu)
    def _encoder_fwd_tc(self, which, frames: List[torch.Tensor], emb):
        P = self.ps.p
        pre = f"perceptual_encoder.rgb_{which}_encoder"
        N = sum(f.shape[0] for f in frames)
        sizes = [frames[0].shape[-1]]
        for (_, s), k in zip(self._CONVS, (8, 4, 3)):
            sizes.append((sizes[-1] - k) // s + 1)
        a1 = self.buf(f"{which}.a1", N, sizes[1], sizes[1], 32)
        # sign masks of the two hidden activations, written by the forward epilogues: the data gradients gate on 4 bytes per
        # (pixel, 32 channels) instead of re-reading the fp32 activation
        a1_bits = self.ibuf(f"{which}.a1_bits", N, sizes[1], sizes[1], 1)
        a2_bits = self.ibuf(f"{which}.a2_bits", N, sizes[2], sizes[2], 2)
        n0 = 0
        for f in frames:  # the first layer reads the reference's NCHW frames as they are
            ops.conv2d_tc_fwd(f, P[f"{pre}.conv_model.0.weight"], P[f"{pre}.conv_model.0.bias"], 4, a1[n0 : n0 + f.shape[0]],
                              relu_bits=a1_bits[n0 : n0 + f.shape[0]])
            n0 += f.shape[0]
        a2 = ops.conv2d_tc_fwd(a1, P[f"{pre}.conv_model.2.weight"], P[f"{pre}.conv_model.2.bias"], 2, self.buf(f"{which}.a2", N, sizes[2], sizes[2], 64),
                               relu_bits=a2_bits)
        a3 = ops.conv2d_tc_fwd(a2, P[f"{pre}.conv_model.4.weight"], P[f"{pre}.conv_model.4.bias"], 1, self.buf(f"{which}.a3", N, sizes[3], sizes[3], 64))
        w0 = None
        if which == "static":
            feat = ops.spatial_softmax_nhwc_fwd(a3, self.buf("static.ss", N, 128), temperature=self.dims.spatial_softmax_temp)
            names = [f"{pre}.fc1.0", f"{pre}.fc2"]
            out = emb[:, 0:64]
        else:
            # nn.Flatten of the reference runs over (C, H, W); the channels-last map flattens as (H, W, C): permute the
            # columns of the following Linear instead of the activation
            feat = a3.view(N, -1)
            PP = sizes[3] * sizes[3]
            w7 = P[f"{pre}.conv_model.7.weight"]
            w0 = self.buf("gripper.w7p", w7.shape[0], PP * 64)
            ops.strided_copy(w0.view(-1, PP, 64), w7.view(-1, 64, PP).transpose(1, 2))
            names = [f"{pre}.conv_model.7", f"{pre}.fc1.0", f"{pre}.fc2"]
            out = emb[:, 64:128]
        acts, stats = self._mlp_ln_fwd(which, feat, names, f"{pre}.ln", out, w0=w0)
        return dict(frames=frames, a1=a1, a2=a2, a3=a3, acts=acts, stats=stats, names=names, pre=pre, w0=w0, a1_bits=a1_bits, a2_bits=a2_bits)

    def _encoder_bwd_tc(self, which, ctx, demb):
        P, G = self.ps.p, self.ps.g
        pre, a1, a2, a3 = ctx["pre"], ctx["a1"], ctx["a2"], ctx["a3"]
        dout = demb[:, 0:64] if which == "static" else demb[:, 64:128]
        N = a3.shape[0]
        if which == "static":
            dss = self._mlp_ln_bwd(which, ctx["acts"], ctx["stats"], ctx["names"], f"{pre}.ln", dout, self.buf("static.dss", N, 128))
            da3 = ops.spatial_softmax_nhwc_bwd(a3, dss, self.buf("static.da3", *a3.shape), temperature=self.dims.spatial_softmax_temp, relu_gate=True)
        else:
            acts, names, w0 = ctx["acts"], ctx["names"], ctx["w0"]
            d = self.buf(f"{which}.mlp_dz", *acts[-1].shape)
            ops.layernorm_bwd(dout, acts[-1], ctx["stats"], P[f"{pre}.ln.weight"], G[f"{pre}.ln.weight"], G[f"{pre}.ln.bias"], dz=d)
            for i in (2, 1):
                nd = self.buf(f"{which}.mlp_d{i}", *acts[i].shape)
                self._linear_bwd(names[i], acts[i], d, nd, gate=acts[i])
                d = nd
            # flatten-FC in the permuted column order: weight gradient into a scratch, then back to the (C,H,W) order
            da3 = self.buf("gripper.da3", *a3.shape)
            PP = a3.shape[1] * a3.shape[2]
            dw0 = self.buf("gripper.dw7p", *w0.shape)
            self.gemm_bwd(d, acts[0], dw0, transA=True)  # (main stream: the permuting copy below reads dw0 right away)
            g7 = G[f"{pre}.conv_model.7.weight"]
            ops.strided_copy(g7.view(-1, 64, PP).transpose(1, 2), dw0.view(-1, PP, 64), accumulate=True)
            self.bias_grad(d, G[f"{pre}.conv_model.7.bias"])
            self.gemm_bwd(d, w0, da3.view(N, -1), gate=acts[0])
        ops.conv2d_tc_wgrad(a2, da3, G[f"{pre}.conv_model.4.weight"], 1, beta=1.0)
        self.bias_grad(da3.view(-1, 64), G[f"{pre}.conv_model.4.bias"])
        da2 = ops.conv2d_tc_dgrad(da3, P[f"{pre}.conv_model.4.weight"], self.buf(f"{which}.da2", *a2.shape), 1, gate=a2, gate_bits=ctx["a2_bits"])
        ops.conv2d_tc_wgrad(a1, da2, G[f"{pre}.conv_model.2.weight"], 2, beta=1.0)
        self.bias_grad(da2.view(-1, 64), G[f"{pre}.conv_model.2.bias"])
        da1 = ops.conv2d_tc_dgrad(da2, P[f"{pre}.conv_model.2.weight"], self.buf(f"{which}.da1", *a1.shape), 2, gate=a1, gate_bits=ctx["a1_bits"])
        n0 = 0
        for f in ctx["frames"]:  # the bias gradient (column sums of da1) comes out of the same tensor-core pass
            ops.conv2d_tc_wgrad(f, da1[n0 : n0 + f.shape[0]], G[f"{pre}.conv_model.0.weight"], 4, beta=1.0, db=G[f"{pre}.conv_model.0.bias"])
            n0 += f.shape[0]

    # bf16 variant (BASELINE config 3): the first layer reads the fp32 frames, everything between the conv layers is channels-last bf16
    def _encoder_fwd_bf16(self, which, frames: List[torch.Tensor], emb):
        P = self.ps.p
        pre = f"perceptual_encoder.rgb_{which}_encoder"
        N = sum(f.shape[0] for f in frames)
        sizes = [frames[0].shape[-1]]
        for (_, s), k in zip(self._CONVS, (8, 4, 3)):
            sizes.append((sizes[-1] - k) // s + 1)
        a1 = self.hbuf(f"{which}.a1h", N, sizes[1], sizes[1], 32)
        a1_bits = self.ibuf(f"{which}.a1_bits", N, sizes[1], sizes[1], 1)
        a2_bits = self.ibuf(f"{which}.a2_bits", N, sizes[2], sizes[2], 2)
        n0 = 0
        for f in frames:
            ops.conv2d_bf16_fwd(f, P[f"{pre}.conv_model.0.weight"], P[f"{pre}.conv_model.0.bias"], 4, a1[n0 : n0 + f.shape[0]], relu_bits=a1_bits[n0 : n0 + f.shape[0]])
            n0 += f.shape[0]
        a2 = ops.conv2d_bf16_fwd(a1, P[f"{pre}.conv_model.2.weight"], P[f"{pre}.conv_model.2.bias"], 2, self.hbuf(f"{which}.a2h", N, sizes[2], sizes[2], 64), relu_bits=a2_bits)
        a3 = ops.conv2d_bf16_fwd(a2, P[f"{pre}.conv_model.4.weight"], P[f"{pre}.conv_model.4.bias"], 1, self.hbuf(f"{which}.a3h", N, sizes[3], sizes[3], 64))
        w0 = None
        if which == "static":
            feat = ops.spatial_softmax_nhwc_bf16_fwd(a3, self.buf("static.ss", N, 128), temperature=self.dims.spatial_softmax_temp)
            names = [f"{pre}.fc1.0", f"{pre}.fc2"]
            out = emb[:, 0:64]
        else:
            feat = a3.view(N, -1)  # bf16: the operand of the flatten-FC as it is
            PP = sizes[3] * sizes[3]
            w7 = P[f"{pre}.conv_model.7.weight"]
            w0 = self.buf("gripper.w7p", w7.shape[0], PP * 64)
            ops.strided_copy(w0.view(-1, PP, 64), w7.view(-1, 64, PP).transpose(1, 2))
            names = [f"{pre}.conv_model.7", f"{pre}.fc1.0", f"{pre}.fc2"]
            out = emb[:, 64:128]
        acts, stats = self._mlp_ln_fwd(which, feat, names, f"{pre}.ln", out, w0=w0)
        return dict(frames=frames, a1=a1, a2=a2, a3=a3, acts=acts, stats=stats, names=names, pre=pre, w0=w0, a1_bits=a1_bits, a2_bits=a2_bits)

    def _encoder_bwd_bf16(self, which, ctx, demb):
        P, G = self.ps.p, self.ps.g
        pre, a1, a2, a3 = ctx["pre"], ctx["a1"], ctx["a2"], ctx["a3"]
        dout = demb[:, 0:64] if which == "static" else demb[:, 64:128]
        N = a3.shape[0]
        da3 = self.hbuf(f"{which}.da3h", *a3.shape)
        if which == "static":
            dss = self._mlp_ln_bwd(which, ctx["acts"], ctx["stats"], ctx["names"], f"{pre}.ln", dout, self.buf("static.dss", N, 128))
            ops.spatial_softmax_nhwc_bf16_bwd(a3, dss, da3, temperature=self.dims.spatial_softmax_temp, relu_gate=True)
        else:
            acts, names, w0 = ctx["acts"], ctx["names"], ctx["w0"]
            d = self.buf(f"{which}.mlp_dz", *acts[-1].shape)
            ops.layernorm_bwd(dout, acts[-1], ctx["stats"], P[f"{pre}.ln.weight"], G[f"{pre}.ln.weight"], G[f"{pre}.ln.bias"], dz=d)
            for i in (2, 1):
                nd = self.buf(f"{which}.mlp_d{i}", *acts[i].shape)
                self._linear_bwd(names[i], acts[i], d, nd, gate=acts[i])
                d = nd
            PP = a3.shape[1] * a3.shape[2]
            dw0 = self.buf("gripper.dw7p", *w0.shape)
            self.gemm_bwd(d, acts[0], dw0, transA=True)  # (main stream: the permuting copy below reads dw0 right away)
            g7 = G[f"{pre}.conv_model.7.weight"]
            ops.strided_copy(g7.view(-1, 64, PP).transpose(1, 2), dw0.view(-1, PP, 64), accumulate=True)
            self.bias_grad(d, G[f"{pre}.conv_model.7.bias"])
            self.gemm_bwd(d, w0, da3.view(N, -1), gate=acts[0])  # bf16 result, gated by the bf16 activation (conv3's ReLU)
        # weight gradients come with the bias gradients (a ones row in the same tensor-core pass): no column-sum launches
        ops.conv2d_bf16_wgrad(a2, da3, G[f"{pre}.conv_model.4.weight"], 1, beta=1.0, db=G[f"{pre}.conv_model.4.bias"])
        da2 = ops.conv2d_bf16_dgrad(da3, P[f"{pre}.conv_model.4.weight"], self.hbuf(f"{which}.da2h", *a2.shape), 1, ctx["a2_bits"])
        ops.conv2d_bf16_wgrad(a1, da2, G[f"{pre}.conv_model.2.weight"], 2, beta=1.0, db=G[f"{pre}.conv_model.2.bias"])
        da1 = ops.conv2d_bf16_dgrad(da2, P[f"{pre}.conv_model.2.weight"], self.hbuf(f"{which}.da1h", *a1.shape), 2, ctx["a1_bits"])
        n0 = 0
        for f in ctx["frames"]:
            ops.conv2d_bf16_wgrad(f, da1[n0 : n0 + f.shape[0]], G[f"{pre}.conv_model.0.weight"], 4, beta=1.0, db=G[f"{pre}.conv_model.0.bias"])
            n0 += f.shape[0]

    # ------------------------------------------------------------------------------------------------------------------
    # recurrent layers (torch.nn.RNN / nn.GRU semantics; decoders/utils/rnn.py:5-36, plan_recognition_net.py:27-34)
    # ------------------------------------------------------------------------------------------------------------------
    def _rnn_fwd(self, tag, pre, w_hh, b_hh, hbuf, col0, S, B, *, kind, reverse=False):
        """Run one direction of one layer.  pre [S*B, G*H] holds x W_ih^T + b_ih (+ b_hh for Elman cells); hbuf has S+2
        time slots of [B, ld] (slot 0 and S+1 stay zero), h_t is written to slot t+1, columns col0:col0+H.
        Tensor-core mode: every step is a split-K 3xTF32 product (fp32-level accuracy on the 32-step chain)."""
        H = w_hh.shape[1]
        h = lambda slot: hbuf[slot, :, col0 : col0 + H]
        pre3 = pre.view(S, B, -1)
        saved = self.buf(f"{tag}.saved", S, B, 4 * H) if kind == "gru" else None
        gh = self.buf(f"{tag}.gh", B, 3 * H) if kind == "gru" else None
        if self.bf16 and self.persistent_rnn and kind != "gru" and ops.rnn_seq_bf16_ok(B, H):
            # bf16 path: one persistent launch (csrc/rnn_push_tc.cu) with the bf16 copy of W_hh resident and the state exchanged as bf16, what
            # the reference's nn.RNN does under 16-bit autocast; any window length.  The exchange buffer doubles as the bf16 twin of the
            # hidden states for the products that consume them (next layer's input product, the weight gradients).
            x16 = self.hbuf(f"{tag}.x16", (S + 1) * B * H)
            st, sp = hbuf.stride(0), pre3.stride(0)
            act = TANH if kind == "tanh" else RELU
            self._written(hbuf)
            if reverse:
                ops.rnn_seq_bf16(self._tw(w_hh), h(S + 1), x16, h(S), pre3[S - 1], S, out_step=-st, add_step=-sp, act=act)
            else:
                ops.rnn_seq_bf16(self._tw(w_hh), h(0), x16, h(1), pre3[0], S, out_step=st, add_step=sp, act=act)
                if col0 == 0 and hbuf.shape[2] == H:
                    x3 = x16.view(S + 1, B, H)
                    for lo in (0, 1):  # h_{t-1} (slots 0..S-1) and h_t (slots 1..S) as [S*B, H] operands
                        t32 = hbuf[lo : lo + S].view(S * B, H)
                        self._twins[self._key(t32)] = [x3[lo : lo + S].view(S * B, H), self._twin_gen, self._span(t32)]
            return saved
        if self.tc and self.persistent_rnn and kind != "gru" and ops.rnn_tc_seq_ok(B, H) and S <= 32:
            # one persistent launch for the whole chain (W_hh resident in shared memory, csrc/rnn_tc.cu).  Forward only up to 32
            # steps: the kernel makes ONE tf32 pass per step, whose rounding of W_hh accumulates along the chain — 0.28 of the
            # logit tolerance at S = 32 but 0.72 at S = 64 (measured with the oracle); longer windows keep the per-step 3xTF32
            # products below.  The backward recurrence only feeds gradients and always uses the persistent kernel.
            st, sp = hbuf.stride(0), pre3.stride(0)
            if reverse:
                ops.rnn_tc_seq(w_hh, h(S + 1), h(S), pre3[S - 1], S, prev_step=-st, out_step=-st, add_step=-sp, act=TANH if kind == "tanh" else RELU)
            else:
                ops.rnn_tc_seq(w_hh, h(0), h(1), pre3[0], S, prev_step=st, out_step=st, add_step=sp, act=TANH if kind == "tanh" else RELU)
            return saved
        for t in (range(S - 1, -1, -1) if reverse else range(S)):
            prev = h(t + 2) if reverse else h(t)
            if kind == "gru":
                self.gemm_rec(prev, w_hh, gh, "fwd", transB=True, bias=b_hh)
                ops.gru_gates_fwd(pre3[t], gh, prev, h(t + 1), saved[t])
            else:
                self.gemm_rec(prev, w_hh, h(t + 1), "fwd", transB=True, addend=pre3[t], act=RELU if kind == "relu" else TANH)
        return saved

    def _rnn_bwd(self, tag, dh_above, w_hh, hbuf, col0, S, B, *, kind, saved=None, reverse=False):
        """BPTT through one direction of one layer.  dh_above [S*B, ld_above] (a column view is fine) is dL/dh_t from the
        consumer.  Returns (dpre [S*B, G*H], dgh [S*B, G*H]): gradients w.r.t. the input-side and hidden-side
        pre-activations (identical tensors for Elman cells)."""
        H, Gn = w_hh.shape[1], self.gates if kind == "gru" else 1
        h = lambda slot: hbuf[slot, :, col0 : col0 + H]
        ab = lambda t: dh_above[t * B : (t + 1) * B]
        if kind == "gru":
            dgi = self.buf(f"{tag}.dgi", S, B, 3 * H)
            dgh = self.buf(f"{tag}.dgh", S, B, 3 * H)
            carry = self.buf(f"{tag}.carry", B, H)
            rec = self.buf(f"{tag}.rec", B, H)
            first = True
            for t in (range(S) if reverse else range(S - 1, -1, -1)):
                prev = h(t + 2) if reverse else h(t)
                ops.gru_gates_bwd(ab(t), None if first else rec, saved[t], prev, dgi[t], dgh[t], carry)
                self.gemm_rec(dgh[t], w_hh, rec, "bwd", addend=carry)
                first = False
            return dgi.view(S * B, 3 * H), dgh.view(S * B, 3 * H)
        # Elman: dpre_t = (dh_above_t + dpre_{t+1} W_hh) * act'(h_t); slot S (or slot 0 for reverse) of dbuf stays zero
        dbuf = self.buf(f"{tag}.dpre", S + 1, B, H, zero=True)
        act = GATE_TANH if kind == "tanh" else 0
        if self.bf16 and self.persistent_rnn and ops.rnn_seq_bf16_ok(B, H):
            sd, sh, sa = dbuf.stride(0), hbuf.stride(0), B * dh_above.stride(0)
            x16 = self.hbuf(f"{tag}.dx16", (S + 1) * B * H)
            self._written(dbuf)
            if reverse:
                ops.rnn_seq_bf16(self._tw(w_hh), dbuf[0], x16, dbuf[1], ab(0), S, out_step=sd, add_step=sa, gate0=h(1), gate_step=sh, act=act, transW=True)
            else:
                ops.rnn_seq_bf16(self._tw(w_hh), dbuf[S], x16, dbuf[S - 1], ab(S - 1), S, out_step=-sd, add_step=-sa, gate0=h(S), gate_step=-sh, act=act,
                                 transW=True)
            d = (dbuf[1:] if reverse else dbuf[:S]).reshape(S * B, H)
            return d, d
        if self.tc and self.persistent_rnn and ops.rnn_tc_seq_ok(B, H):
            sd, sh, sa = dbuf.stride(0), hbuf.stride(0), B * dh_above.stride(0)
            if reverse:
                ops.rnn_tc_seq(w_hh, dbuf[0], dbuf[1], ab(0), S, prev_step=sd, out_step=sd, add_step=sa, gate0=h(1), gate_step=sh, act=act, transW=True)
            else:
                ops.rnn_tc_seq(w_hh, dbuf[S], dbuf[S - 1], ab(S - 1), S, prev_step=-sd, out_step=-sd, add_step=-sa, gate0=h(S), gate_step=-sh, act=act,
                               transW=True)
            d = (dbuf[1:] if reverse else dbuf[:S]).reshape(S * B, H)
            return d, d
        for t in (range(S) if reverse else range(S - 1, -1, -1)):
            nxt, cur = (dbuf[t], dbuf[t + 1]) if reverse else (dbuf[t + 1], dbuf[t])
            self.gemm_rec(nxt, w_hh, cur, "bwd", addend=ab(t), gate=h(t + 1), act=act)
        d = (dbuf[1:] if reverse else dbuf[:S]).reshape(S * B, H)
        return d, d

    # ------------------------------------------------------------------------------------------------------------------
    # the step
    # ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def batch_signature(batch: Dict[str, Dict]) -> tuple:
        """Shapes that size the step's activation buffers: (modality, sequences, window, frame shapes / dtypes)."""
        sig = []
        for m, d in batch.items():
            rgb = d.get("rgb_obs", {})
            sig.append((m, tuple(d["actions"].shape[:2])) + tuple((k, tuple(v.shape[2:]), str(v.dtype)) for k, v in sorted(rgb.items())))
        return tuple(sig)

    @torch.no_grad()
    def step(self, batch: Dict[str, Dict], **kw) -> Dict[str, torch.Tensor]:
        """See _step.  Every distinct batch signature gets its own set of persistent activation buffers: a CUDA graph captured on one shape keeps
        valid pointers when a batch of another shape (the last, partial batch of an epoch) comes through in between."""
        if self._bufs is not self._train_root:  # already inside a namespace (validation / inference / lmp_train)
            return self._step(batch, **kw)
        with self._buffers(("train", self.batch_signature(batch))):
            return self._step(batch, **kw)

    def train_buffers(self, batch) -> Dict[str, torch.Tensor]:
        """The persistent activation buffers of the training step for batches shaped like `batch` (after a step ran)."""
        return self._buf_namespaces[("train", self.batch_signature(batch))]

    def release_buffers(self, namespace) -> None:
        """Drop a buffer namespace (its graphs were evicted)."""
        self._buf_namespaces.pop(namespace, None)

    def _step(self, batch: Dict[str, Dict], *, plan_idx=None, plan_u=None, plan_eps=None, dropout_masks=None, seed: Optional[int] = None,
              backward: bool = True, plan_from: str = "posterior", emb_override: Optional[torch.Tensor] = None,
              goal_override: Optional[torch.Tensor] = None, with_clip: bool = True, defer_encoder_bwd: bool = False,
              aug_shifts: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """One fused forward(+backward) over `batch` (the reference's {"vis": ..., "lang": ...} contract).  Gradients of
        total_loss land in `self.ps.grad` (zeroed first).  Randomness: `plan_idx[m]` / `plan_u[m]` / `plan_eps[m]` and
        `dropout_masks` (dict site -> uint8 keep mask over the whole batch, modalities concatenated in batch order)
        inject it for parity runs; otherwise Philox streams keyed on `seed` (None: the previous seed + 1).  `plan_from="prior"` (forward
        only, validation path) draws the latent plan from the plan-proposal network instead of the recognition network; the KL it reports is
        then meaningless and ignored by the caller.  `emb_override` [nB, S, 128] / `goal_override` [nB, 32] (forward only) replace the outputs of
        the perceptual / goal encoders — the entry point of `Hulc.lmp_train`, which receives them from its caller (hulc.py:254-299)."""
        assert plan_from in ("posterior", "prior") and not (backward and plan_from == "prior")
        assert not (backward and (emb_override is not None or goal_override is not None))
        P, G, ps = self.ps.p, self.ps.g, self.ps
        self._step_shapes.clear()
        self._new_generation()
        self._bias_jobs = []
        if seed is not None:
            self.rng_dev.fill_(int(seed))
        else:
            self.rng_dev.add_(1)
        if self.device.type == "cuda":
            ops.set_rng_offset(self.rng_dev)
        seed = 0 if self.device.type == "cuda" else int(self.rng_dev.item())  # by-value part of the seed (all of it on the emulator)
        mods = list(batch.keys())
        n_mod = len(mods)
        Bs = [batch[m]["actions"].shape[0] for m in mods]
        S = batch[mods[0]]["actions"].shape[1]
        b0s = [sum(Bs[:i]) for i in range(n_mod)]
        nB = sum(Bs)
        N = nB * S
        H = self.H
        p_drop = self.dropout_p
        out: Dict[str, torch.Tensor] = {}

        def drop(site_name, site_id):
            if p_drop <= 0:
                return NO_DROP
            if dropout_masks is not None:
                return Drop(p_drop, keep=dropout_masks[site_name])
            return Drop(p_drop, seed=seed, site=site_id)

        if backward:
            ps.zero_grad()
        # device-side RandomShiftsAug: training steps only; aug_shifts = {"static" | "gripper": [frames, 2] int32 (sx, sy)} injects the draws
        self._aug_ctx = dict(shifts=aug_shifts, seed=seed) if (backward and self.augment_pad) else None
        losses = self.buf("losses", 16, zero=True)  # per modality m: [4m] nll, [4m+1] ce, [4m+2] kl, [4m+3] clip
        losses.zero_()  # slots of a modality order / count seen on an earlier step must not leak into this step's totals
        lview = lambda i: losses[i : i + 1]

        _mark("fwd/perceptual_encoders")
        # ---- perceptual encoders --------------------------------------------------------------------------------------
        emb = self.buf("emb", N, 128)
        if emb_override is not None:
            ops.strided_copy(emb.view(nB, S, 128), emb_override)
            ctx_s = ctx_g = None
        else:
            ctx_s = self._encoder_fwd("static", [batch[m]["rgb_obs"]["rgb_static"].flatten(0, 1) for m in mods], emb)
            ctx_g = self._encoder_fwd("gripper", [batch[m]["rgb_obs"]["rgb_gripper"].flatten(0, 1) for m in mods], emb)
        emb3 = emb.view(nB, S, 128)
        out["perceptual_emb"] = emb3

        _mark("fwd/goal_encoders")
        # ---- goal encoders (goal_encoders.py:31-36, 64-69) ---------------------------------------------------------------
        goal = self.buf("goal", nB, 32)
        goal_ctx = []
        for m, b0, Bm in zip(mods, b0s, Bs):
            if goal_override is not None:
                ops.strided_copy(goal[b0 : b0 + Bm], goal_override[b0 : b0 + Bm])
                continue
            if "lang" in m:
                x, names, ln = batch[m]["lang"], [f"language_goal.mlp.{i}" for i in (1, 3, 5)], "language_goal.ln"
            else:
                x, names, ln = emb3[b0 : b0 + Bm, S - 1, :], [f"visual_goal.mlp.{i}" for i in (0, 2, 4)], "visual_goal.ln"
            acts, stats = self._mlp_ln_fwd(f"goal.{m}", x, names, ln, goal[b0 : b0 + Bm])
            goal_ctx.append((acts, stats, names, ln))
        out["latent_goal"] = goal

        _mark("fwd/plan_proposal")
        # ---- plan proposal = prior (plan_proposal_net.py:42-47): cat(emb[:,0], goal) never materialised ----------------------
        if self.model != "gcbc":
            w0 = P["plan_proposal.fc_model.0.weight"]
            pp = [None, self.buf("pp.a1", nB, self.H_prior)]
            self.gemm_fwd(emb3[:, 0, :], w0[:, :128], pp[1], transB=True, bias=P["plan_proposal.fc_model.0.bias"])
            self.gemm_fwd(goal, w0[:, 128:], pp[1], transB=True, beta=1.0, act=RELU, out="both")
            for j, i in enumerate((2, 4, 6)):
                y = self.buf(f"pp.a{j + 2}", nB, self.H_prior)
                self.gemm_fwd(pp[-1], P[f"plan_proposal.fc_model.{i}.weight"], y, transB=True, bias=P[f"plan_proposal.fc_model.{i}.bias"], act=RELU, out="bf16")
                pp.append(y)
            state_dim = P["plan_proposal.fc_state.0.weight"].shape[0]
            pp_state = self.gemm_fwd(pp[-1], P["plan_proposal.fc_state.0.weight"], self.buf("pp.state", nB, state_dim), transB=True,
                            bias=P["plan_proposal.fc_state.0.bias"])
            out["pp_state"] = pp_state

        _mark("fwd/plan_recognition")
        # ---- plan recognition = posterior -----------------------------------------------------------------------------------
        if self.model == "mcil":
            post = self._birnn_fwd(emb3, S, nB)
            seq_feat = post["seq_feat"]
        else:
            post = self._transformer_fwd(emb3, S, nB, drop)
            seq_feat = post["seq_feat"]
        state_dim = P["plan_recognition.fc_state.0.weight"].shape[0]
        pr_state = self.gemm_fwd(seq_feat, P["plan_recognition.fc_state.0.weight"], self.buf("pr.state", nB, state_dim), transB=True,
                        bias=P["plan_recognition.fc_state.0.bias"])
        out["pr_state"], out["seq_feat"] = pr_state, seq_feat

        _mark("fwd/latent_plan_kl")
        # ---- latent plan sample + KL (hulc.py:289-291, 539-561; distributions.py:23-60) ---------------------------------------
        PF = self.plan_features
        plan = None
        if self.model != "gcbc":
            plan = self.buf("plan", nB, PF)
            if self.discrete:
                kl_rows = self.buf("kl_rows", nB * 32)
                idx_out = self._bufs.get("plan_idx")
                if idx_out is None or idx_out.shape[0] != nB * 32:
                    idx_out = self._bufs["plan_idx"] = torch.zeros(nB * 32, dtype=torch.int32, device=self.device)
                for i, (m, b0, Bm) in enumerate(zip(mods, b0s, Bs)):
                    r0, r1 = b0 * 32, (b0 + Bm) * 32
                    u = plan_u[m].reshape(-1) if plan_u is not None and plan_idx is None else None
                    idx_in = plan_idx[m].reshape(-1).to(torch.int32) if plan_idx is not None else None
                    src, oth = (pr_state, pp_state) if plan_from == "posterior" else (pp_state, pr_state)
                    ops.plan_discrete_fwd(src[b0 : b0 + Bm], oth[b0 : b0 + Bm], plan[b0 : b0 + Bm], kl_rows[r0:r1], u=u,
                                          idx_in=idx_in, idx_out=idx_out[r0:r1], seed=seed, site=100 + i)
                    ops.sum_to(kl_rows[r0:r1], lview(4 * i + 2), 1.0 / Bm)
                out["plan_idx"] = idx_out.view(nB, 32)
            else:
                kl_el = self.buf("kl_el", nB, PF)
                for i, (m, b0, Bm) in enumerate(zip(mods, b0s, Bs)):
                    eps = plan_eps[m] if plan_eps is not None else None
                    src, oth = (pr_state, pp_state) if plan_from == "posterior" else (pp_state, pr_state)
                    ops.plan_cont_fwd(src[b0 : b0 + Bm], oth[b0 : b0 + Bm], plan[b0 : b0 + Bm], kl_el[b0 : b0 + Bm], eps=eps,
                                      seed=seed, site=100 + i)
                    ops.sum_to(kl_el[b0 : b0 + Bm], lview(4 * i + 2), 1.0 / Bm)
            out["sampled_plan"] = plan

        _mark("fwd/action_decoder")
        # ---- action decoder forward (logistic_decoder_rnn.py:260-287) ------------------------------------------------------------
        kind = "gru" if self.rnn_model == "gru_decoder" else "relu"
        Gn = self.gates
        C = 128 - self.percep_lo
        rp = "action_decoder.rnn"
        w_ih0 = P[f"{rp}.weight_ih_l0"]
        w_plan, w_pc, w_goal = w_ih0[:, :PF], w_ih0[:, PF : PF + C], w_ih0[:, PF + C :]
        percep_tm = self.buf("dec.percep", S, nB, C)
        ops.strided_copy(percep_tm, emb3[:, :, self.percep_lo :].transpose(0, 1))
        const = self.buf("dec.const", nB, Gn * H)
        self.gemm_fwd(goal, w_goal, const, transB=True, bias=P[f"{rp}.bias_ih_l0"])
        if PF:
            self.gemm_fwd(plan, w_plan, const, transB=True, beta=1.0)
        hb = [self.buf(f"dec.h{l}", S + 2, nB, H, zero=True) for l in range(2)]
        pre0 = self.buf("dec.pre0", S * nB, Gn * H)
        self.gemm_fwd(percep_tm.view(S * nB, C), w_pc, pre0, transB=True, addend=const, add_mod=nB,
             bias=None if kind == "gru" else P[f"{rp}.bias_hh_l0"])
        sv0 = self._rnn_fwd("dec.l0", pre0, P[f"{rp}.weight_hh_l0"], P[f"{rp}.bias_hh_l0"], hb[0], 0, S, nB, kind=kind)
        h0_all = hb[0][1 : S + 1].view(S * nB, H)
        pre1 = self.buf("dec.pre1", S * nB, Gn * H)
        self.gemm_fwd(h0_all, P[f"{rp}.weight_ih_l1"], pre1, transB=True, bias=P[f"{rp}.bias_ih_l1"],
             addend=None if kind == "gru" else P[f"{rp}.bias_hh_l1"].view(1, -1), add_mod=1)
        sv1 = self._rnn_fwd("dec.l1", pre1, P[f"{rp}.weight_hh_l1"], P[f"{rp}.bias_hh_l1"], hb[1], 0, S, nB, kind=kind)
        h1_all = hb[1][1 : S + 1].view(S * nB, H)
        # heads and their gradient live in [rows, n_pad] buffers (n_pad = n_heads rounded up to 4: aligned rows for the
        # tensor-core products); the pad columns are 0 (zero weight rows / never written) and hidden from the consumers
        n_heads, n_pad = ps.n_heads, ps.n_heads_padded
        heads_p = self.gemm_fwd(h1_all, ps.heads_w, self.buf("dec.heads", S * nB, n_pad), transB=True, bias=ps.heads_b)
        heads = heads_p[:, :n_heads]
        out["heads_tm"] = heads_p.view(S, nB, n_pad)[:, :, :n_heads]

        _mark("fwd/losses")
        # ---- losses (logistic_decoder_rnn.py:121-155,184-231; gripper_control.py:16-36) ----------------------------------------
        has_grip = self.model != "mcil"
        acts_all = self.buf("dec.actions", nB, S, 7)
        for m, b0, Bm in zip(mods, b0s, Bs):
            if has_grip:
                ops.world_to_tcp(batch[m]["actions"].contiguous(), batch[m]["state_info"]["robot_obs"].contiguous(), acts_all[b0 : b0 + Bm], self.nan_flag)
            else:
                ops.strided_copy(acts_all[b0 : b0 + Bm], batch[m]["actions"])
        out["actions_tcp"] = acts_all
        dheads_p = self.buf("dec.dheads", S * nB, n_pad, zero=True)
        dheads = dheads_p[:, :n_heads]
        for i, (m, b0, Bm) in enumerate(zip(mods, b0s, Bs)):
            ops.logistic_loss(heads, acts_all, dheads, losses[4 * i : 4 * i + 2], nB, S, b0, Bm, time_major=True, n_dims=self.n_dims,
                              n_mix=self.n_mix, num_classes=self.num_classes, has_gripper=has_grip, gripper_alpha=self.gripper_alpha,
                              grad_scale=1.0 / n_mod, log_scale_min=self.dims.log_scale_min, act_min=self.dims.act_min, act_max=self.dims.act_max)

        _mark("fwd/clip_aux")
        # ---- CLIP auxiliary loss (hulc.py:650-695, proj_vis_lang.py:23-27), language modality only --------------------------------
        clip_ctx = aux_ctx = None
        if self.model != "mcil" and with_clip:
            for i, (m, b0, Bm) in enumerate(zip(mods, b0s, Bs)):
                if "lang" not in m:
                    continue
                sf, gl = seq_feat[b0 : b0 + Bm], goal[b0 : b0 + Bm]
                H1, Dc = P["proj_vis_lang.mlp_im.0.weight"].shape[0], P["proj_vis_lang.mlp_im.2.weight"].shape[0]
                im1 = self.gemm_fwd(sf, P["proj_vis_lang.mlp_im.0.weight"], self.buf("clip.im1", Bm, H1), transB=True, bias=P["proj_vis_lang.mlp_im.0.bias"], act=RELU)
                im2 = self.gemm_fwd(im1, P["proj_vis_lang.mlp_im.2.weight"], self.buf("clip.im2", Bm, Dc), transB=True, bias=P["proj_vis_lang.mlp_im.2.bias"])
                tx1 = self.gemm_fwd(gl, P["proj_vis_lang.mlp_lang.0.weight"], self.buf("clip.tx1", Bm, H1), transB=True, bias=P["proj_vis_lang.mlp_lang.0.bias"], act=RELU)
                tx2 = self.gemm_fwd(tx1, P["proj_vis_lang.mlp_lang.2.weight"], self.buf("clip.tx2", Bm, Dc), transB=True, bias=P["proj_vis_lang.mlp_lang.2.bias"])
                mask = batch[m].get("use_for_aux_lang_loss")
                mask8 = mask.to(torch.uint8) if mask is not None else None
                d_im2, d_tx2 = self.buf("clip.dim2", Bm, Dc), self.buf("clip.dtx2", Bm, Dc)
                ops.clip_loss(im2, tx2, P["logit_scale"].view(1), mask8, lview(4 * i + 3), d_im2, d_tx2, G["logit_scale"].view(1), grad_scale=self.clip_beta)
                clip_ctx = (i, b0, Bm, sf, gl, im1, tx1, d_im2, d_tx2)
                if self.dims.bc_z or self.dims.mia:
                    aux_ctx = self._aux_heads_fwd(batch[m], mask, sf, im2, tx2, d_im2, d_tx2, Bm)

        _mark("totals")
        # ---- totals (hulc.py:464-491,525) ------------------------------------------------------------------------------------------
        L = losses.view(4, 4)[:n_mod]
        act_m = L[:, 0] + (self.gripper_alpha * L[:, 1] if has_grip else 0.0)
        kl_m = self.kl_beta * L[:, 2] if self.model != "gcbc" else torch.zeros_like(L[:, 2])
        clip = L[:, 3].sum() if (self.model != "mcil" and with_clip) else None
        total = (act_m + kl_m).sum() / n_mod
        if aux_ctx is not None:  # hulc.py:500-519: beta * loss, added before the CLIP term
            if self.dims.bc_z:
                out["lang_pred_loss"] = aux_ctx["losses"][0]
                total = total + self.bc_z_beta * aux_ctx["losses"][0]
            if self.dims.mia:
                out["lang_contrastive_loss"] = aux_ctx["losses"][1]
                total = total + self.mia_beta * aux_ctx["losses"][1]
        if clip is not None:
            total = total + self.clip_beta * clip
            out["lang_clip_loss"] = clip
        out["total_loss"], out["action_loss"], out["kl_loss"] = total, act_m.mean(), kl_m.mean()
        for i, m in enumerate(mods):
            out[f"action_loss_{m}"], out[f"kl_loss_{m}"] = act_m[i], kl_m[i]
        if not backward:
            _mark(None)
            return out

        # ============================================ backward =====================================================================
        demb = self.buf("demb", N, 128)
        demb.zero_()
        demb3 = demb.view(nB, S, 128)
        dgoal = self.buf("dgoal", nB, 32)

        _mark("bwd/action_decoder")
        # heads
        self.gemm_wgrad(dheads_p, h1_all, ps.heads_gw, beta=1.0)
        self.bias_grad(dheads_p, ps.heads_gb)
        dh1 = self.gemm_bwd(dheads_p, ps.heads_w, self.buf("dec.dh1", S * nB, H))
        # layer 1
        dpre1, dgh1 = self._rnn_bwd("dec.l1", dh1, P[f"{rp}.weight_hh_l1"], hb[1], 0, S, nB, kind=kind, saved=sv1)
        self.gemm_wgrad(dgh1, hb[1][0:S].view(S * nB, H), G[f"{rp}.weight_hh_l1"], beta=1.0)
        self.gemm_wgrad(dpre1, h0_all, G[f"{rp}.weight_ih_l1"], beta=1.0)
        self.bias_grad(dpre1, G[f"{rp}.bias_ih_l1"])
        self.bias_grad(dgh1, G[f"{rp}.bias_hh_l1"])
        dh0 = self.gemm_bwd(dpre1, P[f"{rp}.weight_ih_l1"], self.buf("dec.dh0", S * nB, H))
        # layer 0
        dpre0, dgh0 = self._rnn_bwd("dec.l0", dh0, P[f"{rp}.weight_hh_l0"], hb[0], 0, S, nB, kind=kind, saved=sv0)
        self.gemm_wgrad(dgh0, hb[0][0:S].view(S * nB, H), G[f"{rp}.weight_hh_l0"], beta=1.0)
        self.bias_grad(dgh0, G[f"{rp}.bias_hh_l0"])
        g_ih0 = G[f"{rp}.weight_ih_l0"]
        self.gemm_wgrad(dpre0, percep_tm.view(S * nB, C), g_ih0[:, PF : PF + C], beta=1.0)
        dconst = self.buf("dec.dconst", nB, Gn * H)
        colsum(dpre0.view(S, nB * Gn * H), dconst.view(-1))
        self.bias_grad(dconst, G[f"{rp}.bias_ih_l0"])
        self.gemm_wgrad(dconst, goal, g_ih0[:, PF + C :], beta=1.0)
        self.gemm_bwd(dconst, w_goal, dgoal)
        dplan = None
        if PF:
            self.gemm_wgrad(dconst, plan, g_ih0[:, :PF], beta=1.0)
            dplan = self.gemm_bwd(dconst, w_plan, self.buf("dplan", nB, PF))
        dpercep = self.gemm_bwd(dpre0, w_pc, self.buf("dec.dpercep", S * nB, C))
        ops.strided_copy(demb3[:, :, self.percep_lo :].transpose(0, 1), dpercep.view(S, nB, C), accumulate=True)

        _mark("bwd/clip_aux")
        # CLIP head
        dseq = self.buf("dseq", nB, seq_feat.shape[1])
        dseq.zero_()
        if clip_ctx is not None:
            i, b0, Bm, sf, gl, im1, tx1, d_im2, d_tx2 = clip_ctx
            if aux_ctx is not None:  # the MIA gradient joins d_im2 / d_tx2 before they go back through the shared projection head
                self._aux_heads_bwd(aux_ctx, sf, dseq[b0 : b0 + Bm], d_im2, d_tx2, Bm, mia_only=True)
            d_im1 = self._linear_bwd("proj_vis_lang.mlp_im.2", im1, d_im2, self.buf("clip.dim1", *im1.shape), gate=im1)
            self._linear_bwd("proj_vis_lang.mlp_im.0", sf, d_im1, dseq[b0 : b0 + Bm])
            d_tx1 = self._linear_bwd("proj_vis_lang.mlp_lang.2", tx1, d_tx2, self.buf("clip.dtx1", *tx1.shape), gate=tx1)
            self._linear_bwd("proj_vis_lang.mlp_lang.0", gl, d_tx1, dgoal[b0 : b0 + Bm], dx_beta=1.0)
            if aux_ctx is not None:  # BC-Z: accumulates into the sequence feature's gradient after the CLIP head wrote it
                self._aux_heads_bwd(aux_ctx, sf, dseq[b0 : b0 + Bm], d_im2, d_tx2, Bm, mia_only=False)

        _mark("bwd/latent_plan_kl")
        # latent plan
        d_pr = self.buf("d_pr", *pr_state.shape)
        if self.model == "gcbc":
            d_pr.zero_()
        else:
            d_pp = self.buf("d_pp", *pp_state.shape)
            for i, (m, b0, Bm) in enumerate(zip(mods, b0s, Bs)):
                cl, cr = self.kl_beta * self.kl_alpha / Bm / n_mod, self.kl_beta * (1 - self.kl_alpha) / Bm / n_mod
                sl = slice(b0, b0 + Bm)
                if self.discrete:
                    ops.plan_discrete_bwd(pr_state[sl], pp_state[sl], dplan[sl], d_pr[sl], d_pp[sl], cl, cr)
                else:
                    eps = plan_eps[m] if plan_eps is not None else None
                    ops.plan_cont_bwd(pr_state[sl], pp_state[sl], dplan[sl], d_pr[sl], d_pp[sl], cl, cr, eps=eps, seed=seed, site=100 + i)

        _mark("bwd/plan_recognition")
        # posterior
        self._linear_bwd("plan_recognition.fc_state.0", seq_feat, d_pr, dseq, dx_beta=1.0)
        if self.model == "mcil":
            self._birnn_bwd(post, dseq, demb3, S, nB)
        else:
            self._transformer_bwd(post, dseq, demb3, S, nB, drop)

        _mark("bwd/plan_proposal")
        # prior
        if self.model != "gcbc":
            d = d_pp
            names = ["plan_proposal.fc_state.0"] + [f"plan_proposal.fc_model.{i}" for i in (6, 4, 2)]
            for j, n in enumerate(names):
                x = pp[len(pp) - 1 - j]
                nd = self.buf(f"pp.d{j}", nB, self.H_prior)
                self._linear_bwd(n, x, d, nd, gate=x)
                d = nd
            g0 = G["plan_proposal.fc_model.0.weight"]
            self.gemm_wgrad(d, emb3[:, 0, :], g0[:, :128], beta=1.0)
            self.gemm_wgrad(d, goal, g0[:, 128:], beta=1.0)
            self.bias_grad(d, G["plan_proposal.fc_model.0.bias"])
            self.gemm_bwd(d, w0[:, :128], demb3[:, 0, :], beta=1.0)
            self.gemm_bwd(d, w0[:, 128:], dgoal, beta=1.0)

        _mark("bwd/goal_encoders")
        # goal encoders
        for (m, b0, Bm), (acts, stats, names, ln) in zip(zip(mods, b0s, Bs), goal_ctx):
            if "lang" in m:
                self._mlp_ln_bwd(f"goal.{m}", acts, stats, names, ln, dgoal[b0 : b0 + Bm], need_dx=False)
            else:
                self._mlp_ln_bwd(f"goal.{m}", acts, stats, names, ln, dgoal[b0 : b0 + Bm], demb3[b0 : b0 + Bm, S - 1, :], dx_beta=1.0)

        if defer_encoder_bwd:
            # data-parallel overlap: everything but the perceptual encoders' gradients (97 % of the gradient bytes) is final here — the
            # caller starts their all-reduce, then runs `finish_backward()` (the conv stack's backward) underneath it
            self._flush_bias_grads()
            self._join_side()
            self._tail = (ctx_s, ctx_g, demb)
            _mark(None)
            return out
        self._tail = (ctx_s, ctx_g, demb)
        self.finish_backward()
        return out

    @torch.no_grad()
    def finish_backward(self):
        """Backward of the perceptual encoders (the last block of the step; see `defer_encoder_bwd`)."""
        ctx_s, ctx_g, demb = self._tail
        self._tail = None
        _mark("bwd/perceptual_encoders")
        self._encoder_bwd("static", ctx_s, demb)
        self._encoder_bwd("gripper", ctx_g, demb)
        _mark("bwd/bias_gradients")
        self._flush_bias_grads()
        self._join_side()
        _mark(None)

    def encoder_grad_split(self) -> int:
        """Element offset in the flat gradient buffer below which the perceptual encoders' gradients (and `logit_scale`) live: the layout
        is the reference's registration order, encoders first."""
        first = next(k for k in self.ps.keys if not (k.startswith("perceptual_encoder.") or k == "logit_scale"))
        return self.ps.offsets[first][0]

    # ------------------------------------------------------------------------------------------------------------------
    # posterior: transformer (plan_recognition_net.py:94-117)
    # ------------------------------------------------------------------------------------------------------------------
    def _transformer_fwd(self, emb3, S, nB, drop):
        P = self.ps.p
        T, D, Hh = nB * S, 128, self.nhead
        x = ops.add_posemb_fwd(emb3, P["plan_recognition.position_embeddings.weight"], self.buf("tr.x0", nB, S, D), drop("in", 0)).view(T, D)
        layers = []
        for l in range(self.nlayers):
            pre = f"plan_recognition.transformer_encoder.layers.{l}"
            c = dict(x=x, pre=pre)
            c["qkv"] = self.gemm_fwd(x, P[f"{pre}.self_attn.in_proj_weight"], self.buf(f"tr{l}.qkv", T, 3 * D), transB=True, bias=P[f"{pre}.self_attn.in_proj_bias"])
            c["probs"] = self.buf(f"tr{l}.probs", nB, Hh, S, S)
            c["ctx"] = ops.attention_fwd(c["qkv"], self.buf(f"tr{l}.ctx", T, D), c["probs"], nB, S, Hh, drop(f"l{l}.attn", 1 + 4 * l))
            o = self.gemm_fwd(c["ctx"], P[f"{pre}.self_attn.out_proj.weight"], self.buf(f"tr{l}.o", T, D), transB=True, bias=P[f"{pre}.self_attn.out_proj.bias"])
            c["z1"], c["st1"], c["y1"] = self.buf(f"tr{l}.z1", T, D), self.buf(f"tr{l}.st1", T, 2), self.buf(f"tr{l}.y1", T, D)
            ops.layernorm_fwd(o, P[f"{pre}.norm1.weight"], P[f"{pre}.norm1.bias"], c["y1"], c["st1"], res=x, z=c["z1"], drop=drop(f"l{l}.drop1", 2 + 4 * l))
            c["h"] = self.gemm_fwd(c["y1"], P[f"{pre}.linear1.weight"], self.buf(f"tr{l}.h", T, P[f"{pre}.linear1.weight"].shape[0]), transB=True,
                          bias=P[f"{pre}.linear1.bias"], act=RELU, drop=drop(f"l{l}.ffn", 3 + 4 * l), out="bf16")
            f = self.gemm_fwd(c["h"], P[f"{pre}.linear2.weight"], self.buf(f"tr{l}.f", T, D), transB=True, bias=P[f"{pre}.linear2.bias"])
            c["z2"], c["st2"], c["y2"] = self.buf(f"tr{l}.z2", T, D), self.buf(f"tr{l}.st2", T, 2), self.buf(f"tr{l}.y2", T, D)
            ops.layernorm_fwd(f, P[f"{pre}.norm2.weight"], P[f"{pre}.norm2.bias"], c["y2"], c["st2"], res=c["y1"], z=c["z2"], drop=drop(f"l{l}.drop2", 4 + 4 * l))
            x = c["y2"]
            layers.append(c)
        # fc then mean over time == mean over time then fc (both linear): only the (B,128) mean goes through the 4096-wide GEMM
        ybar = ops.reduce_mid(x.view(nB, S, D), self.buf("tr.ybar", nB, D), 1.0 / S)
        seq_feat = self.gemm_fwd(ybar, P["plan_recognition.fc.weight"], self.buf("tr.seq_feat", nB, P["plan_recognition.fc.weight"].shape[0]), transB=True,
                        bias=P["plan_recognition.fc.bias"])
        return dict(layers=layers, ybar=ybar, seq_feat=seq_feat)

    def _transformer_bwd(self, post, dseq, demb3, S, nB, drop):
        P, G = self.ps.p, self.ps.g
        T, D, Hh = nB * S, 128, self.nhead
        dybar = self._linear_bwd("plan_recognition.fc", post["ybar"], dseq, self.buf("tr.dybar", nB, D))
        dy = self.buf("tr.dy", nB, S, D)
        ops.strided_copy(dy, dybar.view(nB, 1, D).expand(nB, S, D), alpha=1.0 / S)
        dy = dy.view(T, D)
        for l in reversed(range(self.nlayers)):
            c = post["layers"][l]
            pre = c["pre"]
            dz2, df = self.buf(f"tr{l}.dz2", T, D), self.buf(f"tr{l}.df", T, D)
            ops.layernorm_bwd(dy, c["z2"], c["st2"], P[f"{pre}.norm2.weight"], G[f"{pre}.norm2.weight"], G[f"{pre}.norm2.bias"], dz=dz2, dx=df,
                              drop=drop(f"l{l}.drop2", 4 + 4 * l))
            dh = self._linear_bwd(f"{pre}.linear2", c["h"], df, self.buf(f"tr{l}.dh", *c["h"].shape), gate=c["h"], drop=drop(f"l{l}.ffn", 3 + 4 * l))
            dy1 = self._linear_bwd(f"{pre}.linear1", c["y1"], dh, self.buf(f"tr{l}.dy1", T, D), addend=dz2)
            dz1, do = self.buf(f"tr{l}.dz1", T, D), self.buf(f"tr{l}.do", T, D)
            ops.layernorm_bwd(dy1, c["z1"], c["st1"], P[f"{pre}.norm1.weight"], G[f"{pre}.norm1.weight"], G[f"{pre}.norm1.bias"], dz=dz1, dx=do,
                              drop=drop(f"l{l}.drop1", 2 + 4 * l))
            dctx = self._linear_bwd(f"{pre}.self_attn.out_proj", c["ctx"], do, self.buf(f"tr{l}.dctx", T, D))
            dqkv = ops.attention_bwd(c["qkv"], c["probs"], dctx, self.buf(f"tr{l}.dqkv", T, 3 * D), nB, S, Hh, drop(f"l{l}.attn", 1 + 4 * l))
            self.gemm_wgrad(dqkv, c["x"], G[f"{pre}.self_attn.in_proj_weight"], beta=1.0)
            self.bias_grad(dqkv, G[f"{pre}.self_attn.in_proj_bias"])
            dy = self.gemm_bwd(dqkv, P[f"{pre}.self_attn.in_proj_weight"], self.buf(f"tr{l}.dx", T, D), addend=dz1)
        dx0 = dy
        d = drop("in", 0)
        if d.p > 0:
            dx0 = ops.dropout_apply(dx0, self.buf("tr.dx0d", T, D), d)
        self.bias_grad(dx0.view(nB, S * D), G["plan_recognition.position_embeddings.weight"][:S].view(-1))
        ops.strided_copy(demb3, dx0.view(nB, S, D), accumulate=True)

    # ------------------------------------------------------------------------------------------------------------------
    # posterior: bidirectional tanh RNN (plan_recognition_net.py:27-42), MCIL
    # ------------------------------------------------------------------------------------------------------------------
    def _birnn_fwd(self, emb3, S, nB):
        P = self.ps.p
        H = self.H_bi
        rp = "plan_recognition.birnn_model"
        x_tm = self.buf("bi.x", S, nB, 128)
        ops.strided_copy(x_tm, emb3.transpose(0, 1))
        inp = x_tm.view(S * nB, 128)
        outs, pres = [], []
        for l in range(2):
            hb = self.buf(f"bi.h{l}", S + 2, nB, 2 * H, zero=True)
            for d, sfx in enumerate(("", "_reverse")):
                pre = self.buf(f"bi.pre{l}{d}", S * nB, H)
                self.gemm_fwd(inp, P[f"{rp}.weight_ih_l{l}{sfx}"], pre, transB=True, bias=P[f"{rp}.bias_ih_l{l}{sfx}"],
                     addend=P[f"{rp}.bias_hh_l{l}{sfx}"].view(1, -1), add_mod=1)
                self._rnn_fwd(f"bi.l{l}{d}", pre, P[f"{rp}.weight_hh_l{l}{sfx}"], None, hb, d * H, S, nB, kind="tanh", reverse=bool(d))
            outs.append(hb)
            pres.append(inp)
            inp = hb[1 : S + 1].view(S * nB, 2 * H)
        seq_feat = outs[1][S]  # x[:, -1]: the last time step, [nB, 4096]
        return dict(outs=outs, inps=pres, seq_feat=seq_feat)

    def _birnn_bwd(self, post, dseq, demb3, S, nB):
        P, G = self.ps.p, self.ps.g
        H = self.H_bi
        rp = "plan_recognition.birnn_model"
        dabove = self.buf("bi.dabove1", S, nB, 2 * H)
        dabove.zero_()
        ops.strided_copy(dabove[S - 1], dseq)
        dabove = dabove.view(S * nB, 2 * H)
        for l in (1, 0):
            hb, inp = post["outs"][l], post["inps"][l]
            dinp = self.buf(f"bi.dinp{l}", *inp.shape)
            for d, sfx in enumerate(("", "_reverse")):
                dpre, _ = self._rnn_bwd(f"bi.l{l}{d}", dabove[:, d * H : (d + 1) * H], P[f"{rp}.weight_hh_l{l}{sfx}"], hb, d * H, S, nB,
                                        kind="tanh", reverse=bool(d))
                hprev = (hb[2 : S + 2] if d else hb[0:S])[:, :, d * H : (d + 1) * H].reshape(S * nB, H)
                self.gemm_wgrad(dpre, hprev, G[f"{rp}.weight_hh_l{l}{sfx}"], beta=1.0)
                self.gemm_wgrad(dpre, inp, G[f"{rp}.weight_ih_l{l}{sfx}"], beta=1.0)
                self.bias_grad(dpre, G[f"{rp}.bias_ih_l{l}{sfx}"])
                self.bias_grad(dpre, G[f"{rp}.bias_hh_l{l}{sfx}"])
                self.gemm_bwd(dpre, P[f"{rp}.weight_ih_l{l}{sfx}"], dinp, beta=float(d))
            dabove = dinp
        ops.strided_copy(demb3.transpose(0, 1), dabove.view(S, nB, 128), accumulate=True)

    # ------------------------------------------------------------------------------------------------------------------
    # ------------------------------------------------------------------------------------------------------------------
    # validation: lmp_val / validation_step, hulc/models/hulc.py:301-388, 739-841
    # ------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def validation_step(self, batch: Dict[str, Dict], **kw) -> Dict[str, torch.Tensor]:
        """See _validation_step; runs on its own buffer set so that a validation batch of another size cannot re-allocate the buffers
        the training step's CUDA graphs were captured on."""
        with self._buffers("validation"):
            return self._validation_step(batch, **kw)

    def _validation_step(self, batch: Dict[str, Dict], *, plan_idx=None, plan_eps=None, sample_u=None, seed: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """The reference's validation metrics for one batch, dropout off (eval mode): for the plan drawn from the proposal network ("pp")
        and from the recognition network ("pr") — action loss, actions sampled from the logistic mixture and mapped back to the world frame,
        their per-sequence L1 error (mae_*: [B, 6]) and gripper success rate — plus the KL and (language modality) the CLIP loss.
        Keys: "<name>_<pp|pr>_<modality>", "kl_loss_<modality>", "val_pred_clip_loss", "sampled_plan_<pp|pr>_<modality>".
        Randomness for parity runs: plan_idx / plan_eps = {"pp" | "pr": {modality: ...}} as in `step`; sample_u = {"pp" | "pr": {modality:
        (u_mix [B,S,n_dims,n_mix], u_inv [B,S,n_dims])}}; otherwise Philox streams keyed on `seed`.
        The shared front of the network is evaluated once per plan source (two forward passes): this is not the hot path."""
        mods = list(batch.keys())
        Bs = [batch[m]["actions"].shape[0] for m in mods]
        S = batch[mods[0]]["actions"].shape[1]
        b0s = [sum(Bs[:i]) for i in range(len(mods))]
        nB = sum(Bs)
        has_grip = self.model != "mcil"
        A = self.n_dims + (1 if has_grip else 0)
        res: Dict[str, torch.Tensor] = {}
        p_save, self.dropout_p = self.dropout_p, 0.0
        try:
            sources = ("pp", "pr") if self.model != "gcbc" else ("pr",)
            for k, which in enumerate(sources):
                out = self.step(batch, plan_idx=(plan_idx or {}).get(which), plan_eps=(plan_eps or {}).get(which), backward=False,
                                seed=None if seed is None else 2 * int(seed) + k, plan_from="prior" if which == "pp" else "posterior")
                rng = 0 if self.device.type == "cuda" else int(self.rng_dev.item())
                heads = self._bufs["dec.heads"][:, : self.ps.n_heads]
                pred_tcp = self.buf(f"val.pred_tcp", nB, S, A)
                pred = self.buf(f"val.pred", nB, S, A)
                mae = self.buf("val.mae", nB, A - 1)
                hits = self.buf("val.hits", nB)
                acts = self.buf("val.actions", nB, S, A)
                for i, (m, b0, Bm) in enumerate(zip(mods, b0s, Bs)):
                    u = (sample_u or {}).get(which, {}).get(m)
                    ops.logistic_sample(heads, pred_tcp, nB, S, b0, Bm, time_major=True, n_dims=self.n_dims, n_mix=self.n_mix, has_gripper=has_grip, log_scale_min=self.dims.log_scale_min,
                                        u_mix=None if u is None else u[0].contiguous(), u_inv=None if u is None else u[1].contiguous(), seed=rng,
                                        site=200 + 2 * i)
                    ops.strided_copy(acts[b0 : b0 + Bm], batch[m]["actions"])
                    if has_grip:
                        ops.tcp_to_world(pred_tcp[b0 : b0 + Bm], batch[m]["state_info"]["robot_obs"].contiguous(), pred[b0 : b0 + Bm], self.nan_flag)
                    else:
                        ops.strided_copy(pred[b0 : b0 + Bm], pred_tcp[b0 : b0 + Bm])
                ops.val_metrics(pred, acts, mae, hits)
                tag = f"_{which}" if self.model != "gcbc" else ""
                for m, b0, Bm in zip(mods, b0s, Bs):
                    res[f"action_loss{tag}_{m}"] = out[f"action_loss_{m}"].clone()
                    res[f"mae{tag}_{m}"] = mae[b0 : b0 + Bm].clone()
                    res[f"gripper_sr{tag}_{m}"] = hits[b0 : b0 + Bm].sum() / float(Bm * S)
                    res[f"sample_act{tag}_{m}"] = pred[b0 : b0 + Bm].clone()
                    if "sampled_plan" in out:
                        res[f"sampled_plan_{which}_{m}"] = out["sampled_plan"][b0 : b0 + Bm].clone()
                    if which == "pr":
                        res[f"kl_loss_{m}"] = out[f"kl_loss_{m}"].clone()
                if which == "pr":
                    res["seq_feat"] = out["seq_feat"].clone()
                    if "lang_clip_loss" in out:
                        res["val_pred_clip_loss"] = out["lang_clip_loss"].clone()
        finally:
            self.dropout_p = p_save
        return res

    # ------------------------------------------------------------------------------------------------------------------
    # the blocks behind Hulc.lmp_train / Hulc.clip_auxiliary_loss as stand-alone forward passes (hulc.py:254-299, 650-695)
    # ------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def lmp_forward(self, perceptual_emb, latent_goal, actions, robot_obs, **inject) -> Dict[str, torch.Tensor]:
        """Prior, posterior, plan sample, decoder and losses for ONE modality from already-encoded inputs: perceptual_emb [B, S, 128],
        latent_goal [B, 32], actions [B, S, 7], robot_obs [B, S, 15] (raw).  Forward only — the training step fuses this block with the
        encoders and the backward pass.  `inject`: plan_idx / plan_u / plan_eps / dropout_masks / seed as in `step`, keyed by "vis"."""
        batch = {"vis": {"actions": actions.contiguous(), "state_info": {"robot_obs": robot_obs.contiguous()}}}
        with self._buffers(("lmp", tuple(actions.shape[:2]))):
            out = self._step(batch, backward=False, emb_override=perceptual_emb.contiguous(), goal_override=latent_goal.contiguous(), with_clip=False, **inject)
            return {k: v.clone() for k, v in out.items()}

    @torch.no_grad()
    def clip_forward(self, seq_feat, latent_goal, mask=None) -> torch.Tensor:
        """CLIP-style contrastive loss of the projected sequence features and language goals (hulc.py:650-695, proj_vis_lang.py:23-27);
        `mask` [B] bool selects the rows (none selected: 0, as the reference's dummy pass)."""
        P, G = self.ps.p, self.ps.g
        Bm = seq_feat.shape[0]
        with self._buffers(("clip", Bm)):
            self._step_shapes.clear()
            self._new_generation()
            sf, gl = seq_feat.contiguous(), latent_goal.contiguous()
            H1, Dc = P["proj_vis_lang.mlp_im.0.weight"].shape[0], P["proj_vis_lang.mlp_im.2.weight"].shape[0]
            im1 = self.gemm_fwd(sf, P["proj_vis_lang.mlp_im.0.weight"], self.buf("clip.im1", Bm, H1), transB=True, bias=P["proj_vis_lang.mlp_im.0.bias"], act=RELU)
            im2 = self.gemm_fwd(im1, P["proj_vis_lang.mlp_im.2.weight"], self.buf("clip.im2", Bm, Dc), transB=True, bias=P["proj_vis_lang.mlp_im.2.bias"])
            tx1 = self.gemm_fwd(gl, P["proj_vis_lang.mlp_lang.0.weight"], self.buf("clip.tx1", Bm, H1), transB=True, bias=P["proj_vis_lang.mlp_lang.0.bias"], act=RELU)
            tx2 = self.gemm_fwd(tx1, P["proj_vis_lang.mlp_lang.2.weight"], self.buf("clip.tx2", Bm, Dc), transB=True, bias=P["proj_vis_lang.mlp_lang.2.bias"])
            loss = self.buf("clip.loss", 1, zero=True)
            loss.zero_()
            scratch = self.buf("clip.dscale", 1)  # the kernel also emits gradients: they go to scratch here
            ops.clip_loss(im2, tx2, P["logit_scale"].view(1), None if mask is None else mask.to(torch.uint8), loss, self.buf("clip.dim2", Bm, Dc), self.buf("clip.dtx2", Bm, Dc),
                          scratch, grad_scale=1.0)
            return loss[0].clone()

    # ------------------------------------------------------------------------------------------------------------------
    # inference: Hulc.step / get_pp_plan_{lang,vision} / predict_with_plan (hulc.py:851-957),
    # LogisticDecoderRNN.act with the carried hidden state (logistic_decoder_rnn.py:104-119)
    # ------------------------------------------------------------------------------------------------------------------
    def _infer_embed(self, rgb_static, rgb_gripper):
        n = rgb_static.shape[0]
        emb = self.buf("emb", n, 128)
        self._encoder_fwd("static", [rgb_static], emb)
        self._encoder_fwd("gripper", [rgb_gripper], emb)
        return emb

    @torch.no_grad()
    def infer_plan(self, rgb_static, rgb_gripper, *, lang=None, plan_idx=None, plan_u=None, plan_eps=None, seed: Optional[int] = None):
        """Start (or re-plan) a rollout: encode the goal and sample a latent plan from the proposal network.  rgb_*: [T, 3, H, W] frames — the
        current observation (T = 1, language goal `lang` [1, 384]) or observation + goal image (T = 2, visual goal = embedding of the last
        frame).  Clears the decoder's hidden state like the reference (hulc.py:926,952).  GCBC: goal only, no plan."""
        P = self.ps.p
        H = self.H
        with self._buffers("inference"):
            self._step_shapes.clear()
            self._new_generation()
            if seed is not None:
                self.rng_dev.fill_(int(seed))
            else:
                self.rng_dev.add_(1)
            if self.device.type == "cuda":
                ops.set_rng_offset(self.rng_dev)
            rng = 0 if self.device.type == "cuda" else int(self.rng_dev.item())
            emb = self._infer_embed(rgb_static, rgb_gripper)
            n = emb.shape[0]
            goal = self.buf("goal", 1, 32)
            if lang is not None:
                self._mlp_ln_fwd("goal.lang", lang, [f"language_goal.mlp.{i}" for i in (1, 3, 5)], "language_goal.ln", goal)
            else:
                self._mlp_ln_fwd("goal.vis", emb[n - 1 : n], [f"visual_goal.mlp.{i}" for i in (0, 2, 4)], "visual_goal.ln", goal)
            plan = None
            if self.model != "gcbc":
                w0 = P["plan_proposal.fc_model.0.weight"]
                x = self.buf("pp.a1", 1, self.H_prior)
                self.gemm_fwd(emb[0:1], w0[:, :128], x, transB=True, bias=P["plan_proposal.fc_model.0.bias"])
                self.gemm_fwd(goal, w0[:, 128:], x, transB=True, beta=1.0, act=RELU)
                for j, i in enumerate((2, 4, 6)):
                    y = self.buf(f"pp.a{j + 2}", 1, self.H_prior)
                    self.gemm_fwd(x, P[f"plan_proposal.fc_model.{i}.weight"], y, transB=True, bias=P[f"plan_proposal.fc_model.{i}.bias"], act=RELU)
                    x = y
                state_dim = P["plan_proposal.fc_state.0.weight"].shape[0]
                pp_state = self.gemm_fwd(x, P["plan_proposal.fc_state.0.weight"], self.buf("pp.state", 1, state_dim), transB=True,
                                         bias=P["plan_proposal.fc_state.0.bias"])
                plan = self.buf("plan", 1, self.plan_features)
                if self.discrete:
                    idx_out = self._bufs.get("plan_idx")
                    if idx_out is None:
                        idx_out = self._bufs["plan_idx"] = torch.zeros(32, dtype=torch.int32, device=self.device)
                    ops.plan_discrete_fwd(pp_state, pp_state, plan, self.buf("kl_rows", 32), u=None if plan_u is None or plan_idx is not None else plan_u.reshape(-1),
                                          idx_in=None if plan_idx is None else plan_idx.reshape(-1).to(torch.int32), idx_out=idx_out, seed=rng, site=300)
                else:
                    ops.plan_cont_fwd(pp_state, pp_state, plan, self.buf("kl_el", 1, self.plan_features), eps=plan_eps, seed=rng, site=300)
            # rollout state in PERSISTENT tensors (a captured control step keeps pointing at them across re-plans): goal, plan, the decoder's
            # carried hidden state (cleared here) and the part of the first layer's input projection that is constant between re-plans
            st = self._infer_state
            if st is None:
                st = self._infer_state = dict(goal=torch.empty(1, 32, device=self.device), plan=None if plan is None else torch.empty_like(plan),
                                              hidden=torch.zeros(2, 1, H, device=self.device), const=torch.empty(1, self.gates * H, device=self.device))
            ops.strided_copy(st["goal"], goal)
            if plan is not None:
                ops.strided_copy(st["plan"], plan)
            st["hidden"].zero_()
            PF, C, rp = self.plan_features, 128 - self.percep_lo, "action_decoder.rnn"
            w_ih0 = P[f"{rp}.weight_ih_l0"]
            self.gemm_fwd(st["goal"], w_ih0[:, PF + C :], st["const"], transB=True, bias=P[f"{rp}.bias_ih_l0"])
            if PF:
                self.gemm_fwd(st["plan"], w_ih0[:, :PF], st["const"], transB=True, beta=1.0)
            self._infer_planned = True
        return self._infer_state["plan"], self._infer_state["goal"]

    def infer_reset(self):
        """Forget the current rollout (the persistent state tensors stay: captured control steps point at them)."""
        self._infer_planned = False

    def enable_infer_graph(self, flag: bool = True):
        """Replay the control step (`infer_act`) from a CUDA graph: the observation is copied into static input buffers, one
        cudaGraphLaunch runs the ~60 kernels of the step, the carried hidden state and the RNG seed advance on the device."""
        self._infer_graph = {} if flag else None

    @torch.no_grad()
    def infer_act(self, rgb_static, rgb_gripper, robot_obs_raw, *, sample_u=None, seed: Optional[int] = None) -> torch.Tensor:
        """One control step: encode the observation ([1, 3, H, W] frames), advance the decoder RNN by one step from the carried hidden state,
        sample an action from the mixture and map it to the world frame (robot_obs_raw [1, 15]).  Returns [1, 1, 7]."""
        if not self._infer_planned:
            raise RuntimeError("infer_plan() starts a rollout")
        graphs = self._infer_graph
        if graphs is None or sample_u is not None or seed is not None or self.device.type != "cuda":
            return self._infer_act_body(rgb_static, rgb_gripper, robot_obs_raw, sample_u=sample_u, seed=seed).clone()
        key = (tuple(rgb_static.shape), tuple(rgb_gripper.shape), str(rgb_static.dtype))
        ent = graphs.get(key)
        if ent is None:
            ins = (torch.empty(rgb_static.shape, dtype=rgb_static.dtype, device=self.device), torch.empty(rgb_gripper.shape, dtype=rgb_gripper.dtype, device=self.device),
                   torch.empty(1, robot_obs_raw.numel(), device=self.device))
            for dst, src in zip(ins, (rgb_static, rgb_gripper, robot_obs_raw.reshape(1, -1))):
                dst.copy_(src, non_blocking=True)
            st = self._infer_state
            keep = (st["hidden"].clone(), self.rng_dev.clone())
            self._infer_act_body(*ins)  # eager warm-up (allocates the buffers); undo its effect on the rollout state
            st["hidden"].copy_(keep[0]); self.rng_dev.copy_(keep[1])
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._infer_act_body(*ins)
            ent = graphs[key] = (g, ins, out)
        g, ins, out = ent
        for dst, src in zip(ins, (rgb_static, rgb_gripper, robot_obs_raw.reshape(1, -1))):
            dst.copy_(src, non_blocking=True)
        g.replay()
        return out.clone()

    def _infer_act_body(self, rgb_static, rgb_gripper, robot_obs_raw, *, sample_u=None, seed: Optional[int] = None) -> torch.Tensor:
        P, ps, st = self.ps.p, self.ps, self._infer_state
        H, Gn, PF = self.H, self.gates, self.plan_features
        kind = "gru" if self.rnn_model == "gru_decoder" else "relu"
        has_grip = self.model != "mcil"
        A = self.n_dims + (1 if has_grip else 0)
        with self._buffers("inference"):
            self._step_shapes.clear()
            self._new_generation()
            if seed is not None:
                self.rng_dev.fill_(int(seed))
            else:
                self.rng_dev.add_(1)
            if self.device.type == "cuda":
                ops.set_rng_offset(self.rng_dev)
            rng = 0 if self.device.type == "cuda" else int(self.rng_dev.item())
            emb = self._infer_embed(rgb_static, rgb_gripper)
            C = 128 - self.percep_lo
            rp = "action_decoder.rnn"
            w_pc = P[f"{rp}.weight_ih_l0"][:, PF : PF + C]
            percep = self.buf("dec.percep", 1, C)
            ops.strided_copy(percep, emb[0:1, self.percep_lo :])
            hb = [self.buf(f"dec.h{l}", 3, 1, H, zero=True) for l in range(2)]
            for l in range(2):
                ops.strided_copy(hb[l][0], st["hidden"][l])  # slot 0 = the hidden state carried from the previous control step
            pre0 = self.buf("dec.pre0", 1, Gn * H)
            self.gemm_fwd(percep, w_pc, pre0, transB=True, addend=st["const"], add_mod=1, bias=None if kind == "gru" else P[f"{rp}.bias_hh_l0"])
            self._rnn_fwd("dec.l0", pre0, P[f"{rp}.weight_hh_l0"], P[f"{rp}.bias_hh_l0"], hb[0], 0, 1, 1, kind=kind)
            pre1 = self.buf("dec.pre1", 1, Gn * H)
            self.gemm_fwd(hb[0][1], P[f"{rp}.weight_ih_l1"], pre1, transB=True, bias=P[f"{rp}.bias_ih_l1"],
                          addend=None if kind == "gru" else P[f"{rp}.bias_hh_l1"].view(1, -1), add_mod=1)
            self._rnn_fwd("dec.l1", pre1, P[f"{rp}.weight_hh_l1"], P[f"{rp}.bias_hh_l1"], hb[1], 0, 1, 1, kind=kind)
            heads_p = self.gemm_fwd(hb[1][1], ps.heads_w, self.buf("dec.heads", 1, ps.n_heads_padded), transB=True, bias=ps.heads_b)
            for l in range(2):
                ops.strided_copy(st["hidden"][l], hb[l][1])
            pred_tcp = self.buf("pred_tcp", 1, 1, A)
            ops.logistic_sample(heads_p[:, : ps.n_heads], pred_tcp, 1, 1, 0, 1, time_major=True, n_dims=self.n_dims, n_mix=self.n_mix, has_gripper=has_grip, log_scale_min=self.dims.log_scale_min,
                                u_mix=None if sample_u is None else sample_u[0].contiguous(), u_inv=None if sample_u is None else sample_u[1].contiguous(),
                                seed=rng, site=310)
            if not has_grip:
                return pred_tcp
            out = self.buf("pred_world", 1, 1, A)
            ops.tcp_to_world(pred_tcp, robot_obs_raw.reshape(1, 1, -1).contiguous(), out, self.nan_flag)
            return out

    def optimizer_step(self, grad_scale=1.0):
        self.ps.adam_step(lr=self.lr, grad_scale=grad_scale)

    def capture(self, batch, *, optimizer: bool = True, grad_scale: float = 1.0) -> "StepGraph":
        """Capture forward + backward (+ Adam) over `batch` into a CUDA graph: ~500 kernel launches become one
        cudaGraphLaunch, which removes the host launch cost that otherwise bounds the 128 dependent recurrent steps.  The
        tensors of `batch` are the graph's static inputs (copy new data into them before each replay); the per-step RNG
        seed and the Adam step count advance on the device."""
        seed0 = self.rng_dev.clone()
        self.step(batch)  # eager warm-up: allocates every activation buffer the step needs
        self.rng_dev.copy_(seed0)  # capture + first replay consume ONE seed increment, like an eager step
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(g):
            out = self.step(batch)
            if optimizer:
                self.ps.adam_step(lr=self.lr, grad_scale=grad_scale)
        if optimizer:
            self.ps.step_count -= 1  # the capture itself did not execute the update
        return StepGraph(self, g, out, optimizer, ops.launch_count() - n0, namespace=("train", self.batch_signature(batch)))

    def capture_split(self, batch):
        """Two graphs for data-parallel training: (1) forward + the backward of everything but the perceptual encoders, (2) the encoders'
        backward.  Between the two replays the caller launches the all-reduce of the already-final part of the gradient
        (`encoder_grad_split()` onwards), which then overlaps the conv stack's backward (hulc_b200.ddp.FlatGradientSync)."""
        seed0 = self.rng_dev.clone()
        ns = ("train", self.batch_signature(batch))
        self.step(batch, defer_encoder_bwd=True)  # eager warm-up in the same two parts (buffers, cached job tables of the bias-gradient launches)
        with self._buffers(ns):
            self.finish_backward()
        self.rng_dev.copy_(seed0)
        torch.cuda.synchronize(self.device)
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(g1):
            out = self.step(batch, defer_encoder_bwd=True)
        n1 = ops.launch_count()
        with torch.cuda.graph(g2):
            with self._buffers(ns):
                self.finish_backward()
        return StepGraph(self, g1, out, False, n1 - n0, namespace=ns), StepGraph(self, g2, out, False, ops.launch_count() - n1, namespace=ns)

    def check_nan_flag(self):
        """The reference asserts on NaNs inside world_to_tcp_frame every step (gripper_control.py:35), which stalls the
        stream; the kernel raises a device flag instead and this reads it on demand."""
        if int(self.nan_flag.item()) != 0:
            raise AssertionError("NaN in world_to_tcp_frame output")
