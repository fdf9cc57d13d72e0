#!/usr/bin/env python
"""bench.py — training sequences/sec of the HULC per-step hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores (oracle port)

A step = forward + backward + (gradient all-reduce) + Adam over one synthetic batch of 32 vision + 32 language
sequences of 32 frames per GPU (BASELINE config 2 shape, fp32).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "training sequences/sec (B=32, seq_len=32)"
UNIT = "seq/s"
B_PER_MODALITY, SEQ_LEN = 32, 32
GFLOP_PER_SEQ = 13.03  # fwd+bwd, SURVEY.md §8(d) / BASELINE.md §3
MB_PER_SEQ = 41.6      # mandatory HBM bytes per sequence, ibid.

# BASELINE.json configs: name -> (model, rnn_model, seq_len, workload text).  `hulc` is the one the metric is quoted on (config 2; with
# --dtype bf16: config 3); the others are the ablations of configs 4 and 5 at their full shapes.
CONFIGS = {
    "hulc": ("hulc", "rnn_decoder", 32, "HULC full model"),
    "mcil": ("mcil", "rnn_decoder", 32, "MCIL ablation (BiRNN posterior, continuous latent; BASELINE config 4)"),
    "gcbc64": ("gcbc", "rnn_decoder", 64, "GCBC ablation, seq_len=64, ReLU-RNN decoder (BASELINE config 5)"),
    "gcbc64-gru": ("gcbc", "gru_decoder", 64, "GCBC ablation, seq_len=64, GRU decoder (BASELINE config 5)"),
}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
def oracle_step_fn(B, S, threads, model="hulc", rnn_model="rnn_decoder", device="cpu", autocast=None):
    """The oracle (restatement of the reference in plain torch, oracle/hulc_oracle.py) as a training step: forward + autograd
    backward + torch.optim.Adam — BASELINE.md §4's baseline recipe.  device="cpu": the CPU baseline (fp32, `threads` host threads);
    device="cuda": the same eager torch program on the GPU through cuDNN / cuBLAS (`eager_b200`), fp32 with
    torch's default TF32 convolutions or under bf16 autocast."""
    import torch

    from hulc_b200.utils import synthetic
    from oracle import hulc_oracle as O

    if device == "cpu":
        torch.set_num_threads(threads)
    mw = max(32, S)
    sd = {k: v.to(device).requires_grad_(True) for k, v in synthetic.make_state_dict(model, rnn_model, max_window=mw).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=2e-4)
    batch = synthetic._to(synthetic.make_batch(B, S, seed=1), device)
    noise = {m: synthetic.plan_noise(B, S, m) for m in batch}
    p = 0.1 if model != "mcil" else 0.0
    masks = {m: {k: v.to(device) for k, v in synthetic.dropout_masks(B, S, m, p).items()} for m in batch} if p > 0 else None
    kw = dict(model=model, rnn_model=rnn_model, dropout_p=p, plan_u={m: noise[m]["u"].to(device) for m in batch}, plan_eps={m: noise[m]["eps"].to(device) for m in batch},
              dropout_masks=masks)

    def step():
        opt.zero_grad(set_to_none=True)
        if autocast is not None:
            with torch.autocast("cuda", dtype=autocast):
                out = O.training_step(sd, batch, **kw)
        else:
            out = O.training_step(sd, batch, **kw)
        out["total_loss"].backward()
        opt.step()
        return out["total_loss"].detach()

    return step


def workload_text(args):
    model, rnn_model, S, what = CONFIGS[args.config]
    prec = "fp32 (BASELINE config 2)" if args.dtype == "fp32" else "bf16 storage / tensor-core operands, fp32 master weights (BASELINE config 3)"
    if args.config != "hulc":
        prec = args.dtype
    return f"{what}, batch=32 vis + 32 lang sequences per GPU, seq_len={S}, 200x200 + 84x84 RGB fp32 frames, 384-d lang emb, {prec}, fwd+bwd+Adam"


def run_reference(args):
    """`--impl reference`: the reference's algorithm on the host cores.  The reference is pure Python/PyTorch and is not
    present on the GPU box, so this arm times the oracle port with every host thread (kind "port") on the SAME workload as
    our arm: the full 32 + 32 sequences per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    model, rnn_model, S, _ = CONFIGS[args.config]
    Bs = B_PER_MODALITY
    step = oracle_step_fn(Bs, S, cores, model, rnn_model)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = 2 * Bs / dt
    sample = f"the full step: {Bs}+{Bs} sequences x {S} frames, fwd+bwd+Adam, fp32, torch CPU ({cores} threads), {args.steps} timed steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------------------------
def dominant_kernel_roofline(torch, eng, batch, peaks, step_ms):
    """Time the heavy kernels of the step one by one (CUDA events on the launching stream) at the step's own shapes —
    the conv layers of the static-camera encoder, the recurrent step GEMM and the large dense GEMMs — and report the
    roofline of the one that takes the largest share of the step."""
    from hulc_b200 import ops

    P = eng.ps.p
    pre = "perceptual_encoder.rgb_static_encoder"
    x = batch["vis"]["rgb_obs"]["rgb_static"].flatten(0, 1)  # one modality: 1024 frames (conv1 runs once per modality)
    n_mod = len(batch)
    B_ = eng.train_buffers(batch)
    sfx = "h" if eng.bf16_conv else ""  # bf16 path: the conv stack's activations are bf16 buffers
    a1, a2, a3, da1, da2, da3 = (B_[f"static.{k}{sfx}"] for k in ("a1", "a2", "a3", "da1", "da2", "da3"))
    n1, N = x.shape[0], a1.shape[0]
    w0, w2, w4 = (P[f"{pre}.conv_model.{i}.weight"] for i in (0, 2, 4))
    b0, b2, b4 = (P[f"{pre}.conv_model.{i}.bias"] for i in (0, 2, 4))
    s0, s2, s4 = torch.empty_like(w0), torch.empty_like(w2), torch.empty_like(w4)
    fl = lambda n, co, ho, k: 2.0 * n * co * ho * ho * k
    if eng.bf16_conv:  # bf16 activations between the conv layers; layer 1 reads the fp32 frames
        cands = {
            "conv1_fwd": (lambda: ops.conv2d_bf16_fwd(x, w0, b0, 4, a1[:n1], relu_bits=B_["static.a1_bits"][:n1]), fl(n1, 32, 49, 192), n_mod),
            "conv2_fwd": (lambda: ops.conv2d_bf16_fwd(a1, w2, b2, 2, a2, relu_bits=B_["static.a2_bits"]), fl(N, 64, 23, 512), 1),
            "conv3_fwd": (lambda: ops.conv2d_bf16_fwd(a2, w4, b4, 1, a3), fl(N, 64, 21, 576), 1),
            "conv3_dgrad": (lambda: ops.conv2d_bf16_dgrad(da3, w4, da2, 1, B_["static.a2_bits"]), fl(N, 64, 21, 576), 1),
            "conv2_dgrad": (lambda: ops.conv2d_bf16_dgrad(da2, w2, da1, 2, B_["static.a1_bits"]), fl(N, 64, 23, 512), 1),
            "conv3_wgrad": (lambda: ops.conv2d_bf16_wgrad(a2, da3, s4, 1), fl(N, 64, 21, 576), 1),
            "conv2_wgrad": (lambda: ops.conv2d_bf16_wgrad(a1, da2, s2, 2), fl(N, 64, 23, 512), 1),
            "conv1_wgrad": (lambda: ops.conv2d_bf16_wgrad(x, da1[:n1], s0, 4), fl(n1, 32, 49, 192), n_mod),
        }
    elif eng.tc:
        cands = {
            "conv1_fwd": (lambda: ops.conv2d_tc_fwd(x, w0, b0, 4, a1[:n1], relu_bits=B_["static.a1_bits"][:n1]), fl(n1, 32, 49, 192), n_mod),
            "conv2_fwd": (lambda: ops.conv2d_tc_fwd(a1, w2, b2, 2, a2, relu_bits=B_["static.a2_bits"]), fl(N, 64, 23, 512), 1),
            "conv3_fwd": (lambda: ops.conv2d_tc_fwd(a2, w4, b4, 1, a3), fl(N, 64, 21, 576), 1),
            "conv3_dgrad": (lambda: ops.conv2d_tc_dgrad(da3, w4, da2, 1, gate=a2, gate_bits=B_["static.a2_bits"]), fl(N, 64, 21, 576), 1),
            "conv2_dgrad": (lambda: ops.conv2d_tc_dgrad(da2, w2, da1, 2, gate=a1, gate_bits=B_["static.a1_bits"]), fl(N, 64, 23, 512), 1),
            "conv3_wgrad": (lambda: ops.conv2d_tc_wgrad(a2, da3, s4, 1), fl(N, 64, 21, 576), 1),
            "conv2_wgrad": (lambda: ops.conv2d_tc_wgrad(a1, da2, s2, 2), fl(N, 64, 23, 512), 1),
            "conv1_wgrad": (lambda: ops.conv2d_tc_wgrad(x, da1[:n1], s0, 4), fl(n1, 32, 49, 192), n_mod),
        }
    else:
        cands = {
            "conv1_fwd": (lambda: ops.conv2d_fwd(x, w0, b0, 4, a1[:n1]), fl(n1, 32, 49, 192), n_mod),
            "conv2_fwd": (lambda: ops.conv2d_fwd(a1, w2, b2, 2, a2), fl(N, 64, 23, 512), 1),
            "conv3_fwd": (lambda: ops.conv2d_fwd(a2, w4, b4, 1, a3), fl(N, 64, 21, 576), 1),
            "conv3_dgrad": (lambda: ops.conv2d_dgrad(da3, w4, a2.shape, 1, gate=a2, dx=da2), fl(N, 64, 21, 576), 1),
            "conv2_dgrad": (lambda: ops.conv2d_dgrad(da2, w2, a1.shape, 2, gate=a1, dx=da1), fl(N, 64, 23, 512), 1),
            "conv3_wgrad": (lambda: ops.conv2d_wgrad(a2, da3, s4, 1), fl(N, 64, 21, 576), 1),
            "conv2_wgrad": (lambda: ops.conv2d_wgrad(a1, da2, s2, 2), fl(N, 64, 23, 512), 1),
            "conv1_wgrad": (lambda: ops.conv2d_wgrad(x, da1[:n1], s0, 4), fl(n1, 32, 49, 192), n_mod),
        }
    # decoder: the recurrent step (64 sequences x 2048 hidden, 2 layers x 32 steps, forward + backward = 128 per step) and
    # the large dense products (dW_hh x2, dW_ih1, dh0 as tf32; the layer-1 input projection as 3xTF32)
    H, S, nB = eng.H, batch["vis"]["actions"].shape[1], sum(d["actions"].shape[0] for d in batch.values())
    whh = P["action_decoder.rnn.weight_hh_l1"]
    hb, pre1 = B_["dec.h1"], B_["dec.pre1"]
    big_a, big_c = B_["dec.dh0"], torch.empty(H, H, device=x.device)
    elman = eng.rnn_model == "rnn_decoder"
    if not elman:
        pass  # GRU decoder: per-step GEMM + gate kernels (not timed one by one here)
    elif eng.bf16 and eng.persistent_rnn and ops.rnn_seq_bf16_ok(nB, H):
        # bf16 mode: the whole chain of one layer is one persistent launch of the push kernel (csrc/rnn_push_tc.cu), bf16 W_hh and state
        st, sp = hb.stride(0), pre1.view(S, nB, -1).stride(0)
        pre3 = pre1.view(S, nB, -1)
        w16 = eng._tw(whh)
        x16, dx16 = B_["dec.l1.x16"], B_["dec.l1.dx16"]
        cands[f"rnn_seq_fwd_{S}steps"] = (lambda: ops.rnn_seq_bf16(w16, hb[0], x16, hb[1], pre3[0], S, out_step=st, add_step=sp, act=1), 2.0 * nB * H * H * S, 2)
        dbuf = B_["dec.l1.dpre"]
        sd = dbuf.stride(0)
        dh = big_a.view(S, nB, H)
        cands[f"rnn_seq_bwd_{S}steps"] = (lambda: ops.rnn_seq_bf16(w16, dbuf[S], dx16, dbuf[S - 1], dh[S - 1], S, out_step=-sd, add_step=-dh.stride(0),
                                                                   gate0=hb[S], gate_step=-st, act=0, transW=True), 2.0 * nB * H * H * S, 2)
    elif eng.tc and eng.persistent_rnn:
        # the whole 32-step chain of one layer is one persistent launch (csrc/rnn_tc.cu); 2 layers forward + 2 backward per step
        st, sp = hb.stride(0), pre1.view(S, nB, -1).stride(0)
        pre3 = pre1.view(S, nB, -1)
        if S <= 32:
            cands["rnn_seq_fwd_32steps"] = (lambda: ops.rnn_tc_seq(whh, hb[0], hb[1], pre3[0], S, prev_step=st, out_step=st, add_step=sp, act=1),
                                            2.0 * nB * H * H * S, 2)
        dbuf = B_["dec.l1.dpre"]
        sd = dbuf.stride(0)
        dh = big_a.view(S, nB, H)
        cands["rnn_seq_bwd_32steps"] = (lambda: ops.rnn_tc_seq(whh, dbuf[S], dbuf[S - 1], dh[S - 1], S, prev_step=-sd, out_step=-sd, add_step=-dh.stride(0),
                                                               gate0=hb[S], gate_step=-st, act=0, transW=True), 2.0 * nB * H * H * S, 2)
    elif eng.tc:
        cands["rnn_step_fwd_3xtf32"] = (lambda: ops.gemm(hb[1], whh, hb[2], transB=True, addend=pre1[:nB], act=1, tc=3), 2.0 * nB * H * H, 2 * S)
        dbuf = B_["dec.l1.dpre"]
        cands["rnn_step_bwd_tf32"] = (lambda: ops.gemm(dbuf[2], whh, dbuf[1], addend=big_a[:nB], gate=hb[2], tc=1), 2.0 * nB * H * H, 2 * S)
    else:
        cands["rnn_step_gemm"] = (lambda: ops.gemm(hb[1], whh, hb[2], transB=True, addend=pre1[:nB], act=1), 2.0 * nB * H * H, 4 * S)
    mode = 1 if eng.tc else 0
    if elman and eng.bf16:  # bf16 path: bf16 operands (TMA-fed kind::f16 kernel), fp32 result
        a16, h16, w16 = big_a.to(torch.bfloat16), hb[1 : S + 1].reshape(S * nB, H).to(torch.bfloat16), whh.to(torch.bfloat16)
        cands["dense_wgrad_2048^3"] = (lambda: ops.gemm_bf16(a16, h16, big_c, transA=True), 2.0 * H * H * S * nB, 4)
        cands["dense_fwd_2048^3"] = (lambda: ops.gemm_bf16(a16, w16, pre1, transB=True), 2.0 * H * H * S * nB, 1)
    elif elman:
        cands["dense_wgrad_2048^3"] = (lambda: ops.gemm(big_a, hb[1 : S + 1].view(S * nB, H), big_c, transA=True, tc=mode), 2.0 * H * H * S * nB, 4)
        cands["dense_fwd_2048^3"] = (lambda: ops.gemm(big_a, whh, pre1, transB=True, tc=3 if eng.tc else 0), 2.0 * H * H * S * nB, 1)
    res = {}
    for name, (fn, flops, per_step) in cands.items():
        try:
            fn(); fn()
            torch.cuda.synchronize()
        except Exception as e:  # name the kernel before the context dies
            raise RuntimeError(f"roofline candidate {name} failed: {e}") from e
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[name] = dict(ms=ms, flops=flops, per_step=per_step, share=ms * per_step / step_ms)
    # algorithmic HBM bytes per launch: operands read once + result written once, fp32
    fr = lambda n, c, hw: 4.0 * n * c * hw * hw
    abytes = {
        "conv1_fwd": fr(n1, 3, 200) + fr(n1, 32, 49), "conv1_wgrad": fr(n1, 3, 200) + fr(n1, 32, 49),
        # data gradients: dY in, dX out, + the ReLU sign mask of the gating activation (1 bit per element)
        "conv2_fwd": fr(N, 32, 49) + fr(N, 64, 23), "conv2_wgrad": fr(N, 32, 49) + fr(N, 64, 23), "conv2_dgrad": fr(N, 64, 23) + (1 + 1 / 32) * fr(N, 32, 49),
        "conv3_fwd": fr(N, 64, 23) + fr(N, 64, 21), "conv3_wgrad": fr(N, 64, 23) + fr(N, 64, 21), "conv3_dgrad": fr(N, 64, 21) + (1 + 1 / 32) * fr(N, 64, 23),
        "rnn_seq_fwd_32steps": 4.0 * (H * H + 3 * S * nB * H), "rnn_seq_bwd_32steps": 4.0 * (H * H + 4 * S * nB * H),
        "rnn_step_fwd_3xtf32": 4.0 * (H * H + 3 * nB * H), "rnn_step_bwd_tf32": 4.0 * (H * H + 4 * nB * H), "rnn_step_gemm": 4.0 * (H * H + 3 * nB * H),
        "dense_wgrad_2048^3": 4.0 * (2 * S * nB * H + H * H), "dense_fwd_2048^3": 4.0 * (2 * S * nB * H + H * H),
    }
    if eng.bf16_conv:  # bf16 activations (2 bytes); layer 1 still reads the fp32 frames
        f2 = lambda n, c, hw: 2.0 * n * c * hw * hw
        abytes.update({
            "conv1_fwd": fr(n1, 3, 200) + f2(n1, 32, 49), "conv1_wgrad": fr(n1, 3, 200) + f2(n1, 32, 49),
            "conv2_fwd": f2(N, 32, 49) + f2(N, 64, 23), "conv2_wgrad": f2(N, 32, 49) + f2(N, 64, 23), "conv2_dgrad": f2(N, 64, 23) + (1 + 1 / 16) * f2(N, 32, 49),
            "conv3_fwd": f2(N, 64, 23) + f2(N, 64, 21), "conv3_wgrad": f2(N, 64, 23) + f2(N, 64, 21), "conv3_dgrad": f2(N, 64, 21) + (1 + 1 / 16) * f2(N, 64, 23),
        })
    if eng.bf16:  # bf16 operands, fp32 result
        # recurrence: bf16 W_hh once, per step the fp32 addend in, the fp32 result + its bf16 exchange copy out (+ the fp32 gate in BPTT)
        abytes[f"rnn_seq_fwd_{S}steps"] = 2.0 * H * H + 10.0 * S * nB * H
        abytes[f"rnn_seq_bwd_{S}steps"] = 2.0 * H * H + 14.0 * S * nB * H
        abytes["dense_wgrad_2048^3"] = 2.0 * 2 * S * nB * H + 4.0 * H * H
        abytes["dense_fwd_2048^3"] = 2.0 * (S * nB * H + H * H) + 4.0 * S * nB * H
    top = max(res, key=lambda k: res[k]["ms"] * res[k]["per_step"])
    r = res[top]
    tflops = r["flops"] / (r["ms"] * 1e-3) / 1e12
    gbs = abytes[top] / (r["ms"] * 1e-3) / 1e9
    # the kernel runs tf32 operands: its tensor ceiling is half the measured bf16 rate; it is HBM-bound when its arithmetic
    # intensity is below that ridge
    is_bf16_kernel = eng.bf16 and (top.startswith(("dense", "rnn_seq")) or (eng.bf16_conv and top.startswith(("conv2", "conv3"))))
    tf32_peak = peaks["tf_burst"] / (1.0 if is_bf16_kernel else 2.0)
    hbm_bound = (r["flops"] / abytes[top]) < (tf32_peak * 1e12) / (peaks["hbm"] * 1e9)
    kern = {k: {"ms": round(v["ms"], 4), "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2), "gbs": round(abytes[k] / (v["ms"] * 1e-3) / 1e9, 1),
                "per_step": v["per_step"], "share_of_step": round(v["share"], 4)} for k, v in res.items()}
    traffic = None
    tp = ROOT / "profiles" / "r02_traffic.json"  # dram bytes per launch from the committed ncu --set full captures of the same kernels
    if tp.exists():
        traffic = json.loads(tp.read_text()).get("bf16" if eng.bf16 else "tf32", {}).get(top, {}).get("dram_bytes_per_launch")
    peak_note = ("the kernel runs bf16 operands (kind::f16)" if is_bf16_kernel else "the kernel runs tf32 operands, whose tensor-core peak is half of it")
    common = {"kernel": top, "traffic": traffic, "traffic_source": "profiles/r02_traffic.json (ncu --set full, dram read+write bytes per launch)" if traffic else None, "ms_per_launch": r["ms"], "launches_per_step": r["per_step"], "share_of_step": r["share"],
              "algorithmic_bytes_per_launch": abytes[top], "algorithmic_flops_per_launch": r["flops"], "kernels": kern,
              "step": {"tensor_frac": None, "hbm_frac": None}}
    if top.startswith("rnn_seq"):  # a chain of S dependent 64 x 2048 x 2048 products: latency per dependent step is the figure that describes it
        common["note"] = (f"{S} dependent steps in one persistent launch: {r['ms'] * 1e3 / S:.2f} us per dependent step; a latency chain, neither the tensor nor the "
                          "HBM roof bounds it (DESIGN.md section 4)")
        common["us_per_dependent_step"] = r["ms"] * 1e3 / S
    if hbm_bound:
        return {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                "peak_source": f"{peaks['src']} HBM copy bandwidth (MEASURED_PEAKS.json hbm_gbs)", **common}
    return {"bound": "tensor", "achieved": tflops, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": tflops / peaks["tf_burst"],
            "peak_source": f"{peaks['src']} bf16 dense burst (kernel timed alone); {peak_note}", **common}


def eager_b200(torch, args, dev):
    """BASELINE.md §4 step 4 / SURVEY §8d: the same algorithm as an EAGER torch program on this GPU (cuDNN convolutions, cuBLAS GEMMs,
    autograd, torch.optim.Adam) — the "existing Blackwell kernels" bar.  Full batch, CUDA-event timed, resident inputs."""
    import warnings

    from oracle import hulc_oracle as O

    model, rnn_model, S, _ = CONFIGS[args.config]

    # the reference's recurrent layers are nn.RNN / nn.GRU modules, i.e. cuDNN's fused kernels on a GPU: run them that way (the oracle
    # unrolls them step by step in Python, which would understate this bar)
    def flat(sd, prefix, layers, sfxs=("",)):
        return [sd[f"{prefix}.{n}_l{l}{sfx}"] for l in range(layers) for sfx in sfxs for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]

    def elman(sd, prefix, x, num_layers, nonlinearity, reverse_suffix="", h0=None):
        w = flat(sd, prefix, num_layers)
        h0 = x.new_zeros(num_layers, x.shape[0], w[1].shape[1]) if h0 is None else h0
        return (torch._VF.rnn_relu if nonlinearity == "relu" else torch._VF.rnn_tanh)(x, h0, w, True, num_layers, 0.0, True, False, True)[0]

    def birnn(sd, prefix, x, num_layers=2):
        w = flat(sd, prefix, num_layers, ("", "_reverse"))
        return torch._VF.rnn_tanh(x, x.new_zeros(2 * num_layers, x.shape[0], w[1].shape[1]), w, True, num_layers, 0.0, True, True, True)[0]

    def gru(sd, prefix, x, num_layers):
        w = flat(sd, prefix, num_layers)
        return torch._VF.gru(x, x.new_zeros(num_layers, x.shape[0], w[1].shape[1]), w, True, num_layers, 0.0, True, False, True)[0]

    saved = (O.elman_rnn, O.birnn_tanh, O.gru)
    O.elman_rnn, O.birnn_tanh, O.gru = elman, birnn, gru
    warnings.filterwarnings("ignore", message=".*weights are not part of single contiguous chunk.*")
    res = {}
    for name, ac in (("fp32_tf32conv", None), ("bf16_autocast", torch.bfloat16)):
        try:
            step = oracle_step_fn(B_PER_MODALITY, S, 1, model, rnn_model, device=str(dev), autocast=ac)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            res[name] = {"ms_per_step": ms, "value": 2 * B_PER_MODALITY / (ms * 1e-3), "unit": UNIT, "loss": float(loss)}
        except Exception as e:  # a comparator must never take the bench line down
            res[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    O.elman_rnn, O.birnn_tanh, O.gru = saved
    res["what"] = ("oracle/hulc_oracle.py (plain torch restatement of the reference) run eagerly on cuda: cuDNN convolutions and cuDNN nn.RNN/nn.GRU kernels, cuBLAS, autograd, torch.optim.Adam, "
                   "torch defaults (TF32 convolutions, fp32 matmuls) and under torch.autocast(bfloat16); 32+32 sequences, 3 warm-up + 10 timed steps")
    return res


def inference_latency(torch, args, dev):
    """batch-1 latency of the rollout path `Hulc.step` (hulc.py:851-870) — observation on the host -> action on the host,
    per control step, with the control step replayed from a CUDA graph; re-planning steps (every `replan_freq` = 30) reported apart."""
    from hulc_b200.models.hulc import Hulc
    from hulc_b200.utils import synthetic

    cfg = synthetic.model_config("hulc", target_root="hulc_b200")
    cfg.pop("_target_"); cfg.pop("_recursive_")
    model = Hulc(**cfg, device=dev)
    model.load_state_dict(synthetic.make_state_dict("hulc"), strict=False)
    steps = 240
    r = synthetic.rollout_inputs(steps, "lang", seed=3)
    import numpy as np

    model.lang_embeddings = {"goal": r["lang"].numpy()[None]}
    obs = [{"rgb_obs": {"rgb_static": r["rgb_static"][t][None, None].pin_memory(), "rgb_gripper": r["rgb_gripper"][t][None, None].pin_memory()},
            "robot_obs_raw": r["robot_obs_raw"][t][None, None].pin_memory()} for t in range(steps)]
    out = {}
    for mode in ("eager", "graph"):
        model.enable_cuda_graph_inference(mode == "graph")
        model.reset()
        for t in range(35):  # warm-up incl. one re-plan and the graph capture
            model.step(obs[t], "goal")
        model.reset()
        torch.cuda.synchronize()
        act_us, plan_us = [], []
        for t in range(steps):
            t0 = time.perf_counter()
            a = model.step(obs[t], "goal").cpu()
            dt = (time.perf_counter() - t0) * 1e6
            (plan_us if t % model.replan_freq == 0 else act_us).append(dt)
        q = lambda v, p: float(np.percentile(v, p))
        out[mode] = {"p50_us": q(act_us, 50), "p99_us": q(act_us, 99), "mean_us": float(np.mean(act_us)), "replan_step_p50_us": q(plan_us, 50), "steps": len(act_us)}
    out["what"] = "Hulc.step: pinned-host observation (200x200 + 84x84 fp32 frames, 15-d robot state) -> sampled world-frame action on the host, batch 1, wall clock per control step"
    return out


def bind_to_gpu_numa(index: int):
    """Pin this rank's host threads (and so the first-touch placement of its pinned staging buffers) to the CPUs next to its GPU: with all
    eight ranks on NUMA node 0 the host-to-device feed of the end-to-end run saturates one memory controller (r01: e2e efficiency 0.42 at N=8)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    from hulc_b200 import ops
    from hulc_b200.models.gcbc import GCBC
    from hulc_b200.models.hulc import Hulc
    from hulc_b200.utils import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hulc_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    mname, rnn_model, S, _ = CONFIGS[args.config]
    precision = "tf32" if args.dtype == "fp32" else args.dtype

    cfg = synthetic.model_config(mname, rnn_model=rnn_model, max_window=max(32, S), target_root="hulc_b200")
    cfg.pop("_target_"); cfg.pop("_recursive_")
    model = (GCBC if mname == "gcbc" else Hulc)(**cfg, device=dev, precision=precision)
    model.load_state_dict(synthetic.make_state_dict(mname, rnn_model, max_window=max(32, S)), strict=False)  # identical weights on every rank
    eng = model.engine
    opt = model.configure_optimizers()["optimizer"]

    host = synthetic.make_batch(B_PER_MODALITY, S, seed=1 + rank)  # per-rank data (weak scaling)
    batch = synthetic._to(host, dev)

    from hulc_b200.ddp import FlatGradientSync

    sync = FlatGradientSync(eng)
    sync.broadcast_parameters(0)

    def allreduce_and_adam():
        sync.sync()  # the one per-step collective: 188 MB fp32 gradient over NVLink; no-op on one GPU
        sync.step()  # fused Adam with grad_scale = 1 / world

    # the whole step (forward, backward and — on one GPU — Adam) is replayed from one CUDA graph; with several ranks the
    # gradient all-reduce and Adam follow the graph
    if world == 1:
        sg = eng.capture(batch, optimizer=True)
        graph_launches = sg.launches
    else:
        # data parallel: the all-reduce of everything behind the encoders' gradients runs under the conv stack's backward
        sg, sg_tail = eng.capture_split(batch)
        graph_launches = sg.launches + sg_tail.launches

    def step(i, b):
        out = sg.replay()
        if world > 1:
            sync.sync_head()
            sg_tail.replay()
            sync.sync_tail()
            sync.step()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------------------------------
    for i in range(max(3, args.warmup)):
        step(i, batch)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for i in range(args.steps):
            out = step(100 + i, batch)
        e1.record()
        barrier()
    launches = (graph_launches + (1 if world > 1 else 0)) * args.steps
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    seqs = 2 * B_PER_MODALITY * world
    value = seqs / (ms * 1e-3)
    eng.check_nan_flag()
    loss = float(out["total_loss"].item())

    # ---- end to end: pinned host batch -> H2D -> step -> D2H loss, through the public model API ---------------------------
    def pin(d):
        if isinstance(d, dict):
            return {k: pin(v) for k, v in d.items()}
        return d.pin_memory() if torch.is_tensor(d) else d

    def tensors(d):
        for v in d.values():
            if isinstance(v, dict):
                yield from tensors(v)
            elif torch.is_tensor(v):
                yield v

    def run_e2e(host_batch, lightning_contract=False):
        """Pinned host batch -> double-buffered H2D on a copy stream -> model.training_step (graph replay) -> all-reduce + Adam -> loss D2H.
        lightning_contract: what Lightning's loop does with the module on one GPU, autograd enabled —
        `loss = training_step(batch); loss.backward(); optimizer.step(); optimizer.zero_grad()`."""
        hostp = pin(host_batch)
        nbytes = sum(t_.numel() * t_.element_size() for t_ in tensors(hostp))
        copy_stream = torch.cuda.Stream()
        stage = [synthetic._to(host_batch, dev), synthetic._to(host_batch, dev)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def upload(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                for dst, src in zip(tensors(stage[slot]), tensors(hostp)):
                    dst.copy_(src, non_blocking=True)
                ready[slot].record(copy_stream)

        def e2e_loop(n):
            losses = []
            upload(0)
            for i in range(n):
                slot = i & 1
                if i + 1 < n:
                    upload(slot ^ 1)  # next step's inputs stream in while this step computes
                torch.cuda.current_stream().wait_event(ready[slot])
                loss_t = model.training_step(stage[slot], i)
                consumed[slot].record()
                if lightning_contract:
                    loss_t.backward()
                    opt.step()
                    opt.zero_grad()
                else:
                    allreduce_and_adam()
                losses.append(loss_t.item())  # device -> host read of the step's result
            return losses

        for ev in consumed:
            ev.record()
        model.enable_cuda_graphs()
        with torch.set_grad_enabled(lightning_contract):
            e2e_loop(4)  # captures one graph per staging slot, then warm replays
            barrier()
            t0 = time.perf_counter()
            e2e_loop(args.steps)
            barrier()
        ms_ = (time.perf_counter() - t0) / args.steps * 1e3
        t_ = torch.tensor([ms_], device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item()), nbytes

    # End to end through the public API from pinned host memory.  Primary: the frames as the dataset stores them (uint8) — scale + normalise
    # run on the device, 4x fewer host bytes than the reference's fp32 batch contract, which is reported next to it.
    host_u8 = {m: {k: (dict(v) if isinstance(v, dict) else v) for k, v in d.items()} for m, d in host.items()}
    for m in host_u8:
        host_u8[m]["rgb_obs"] = {k: ((v * 0.5 + 0.5) * 255).round().clamp(0, 255).to(torch.uint8) for k, v in host[m]["rgb_obs"].items()}
    u8_ms, u8_bytes = run_e2e(host_u8)
    e2e_ms, h2d_bytes = run_e2e(host)
    e2e_value = seqs / (e2e_ms * 1e-3)
    lc = None
    if world == 1:
        lc_ms, lc_bytes = run_e2e(host_u8, lightning_contract=True)
        lc = {"value": seqs / (lc_ms * 1e-3), "unit": UNIT, "ms_per_step": lc_ms, "h2d_bytes_per_step": lc_bytes, "d2h_bytes_per_step": 4,
              "api": "autograd enabled: loss = Hulc.training_step(batch); loss.backward(); FusedAdam.step(); zero_grad() — uint8 frames; the gradients autograd "
                     "hands to the parameters are views of the engine's flat buffer (no copies)"}

    if rank == 0:
        roof = dominant_kernel_roofline(torch, eng, batch, peaks, ms)
        per_gpu = value / world
        roof["step"] = {
            "tensor_frac": per_gpu * GFLOP_PER_SEQ * 1e9 / (peaks["tf_sustained"] * 1e12),
            "hbm_frac": per_gpu * MB_PER_SEQ * 1e6 / (peaks["hbm"] * 1e9),
            "note": "whole step vs SURVEY.md §8(d) bounds: 13.03 GFLOP and 41.6 MB per sequence (HULC, S=32); sustained bf16 peak",
        }
        cpu = eager = lat = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            Bs = B_PER_MODALITY
            ostep = oracle_step_fn(Bs, S, cores, mname, rnn_model)
            ostep()
            t0 = time.perf_counter()
            n = 2
            for _ in range(n):
                ostep()
            dt = (time.perf_counter() - t0) / n
            cpu = {"value": 2 * Bs / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"the full step ({Bs}+{Bs} sequences x {S} frames, fwd+bwd+Adam fp32), 1 warm-up + {n} timed steps of the oracle (torch CPU, {cores} threads)"}
        if world == 1 and not args.no_eager:
            eager = eager_b200(torch, args, dev)
        if world == 1 and not args.no_latency and args.config == "hulc":
            try:
                lat = inference_latency(torch, args, dev)
            except Exception as e:
                lat = {"error": f"{type(e).__name__}: {e}"[:300]}
        dtype = {"tf32": "f32 storage/accumulate; tf32 tensor-core convs and backward GEMMs, 3xTF32 forward GEMMs, fp32 CUDA-core elsewhere",
                 "fp32": "f32", "bf16": "bf16 activations / weight copies on the tensor cores (kind::f16), fp32 accumulate, fp32 master weights + Adam, fp32 losses / LayerNorm / softmax"}[precision]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload_text(args), "name": args.config,
                       "parallelism": f"dp{world}", "launch": "forward+backward(+Adam) replayed from one CUDA graph" if world == 1 else
                       "two CUDA graphs per step; the NCCL all-reduce of 97 % of the gradient bytes overlaps the second (the conv stack's backward), Adam follows", "l2": "per-step inputs are 1.16 GB per GPU (> 126 MB L2); no explicit flush",
                       "dropout_p": eng.dropout_p},
            "clocks": clk.summary(), "gpu_launches": launches, "launches_per_step": launches / args.steps, "loss": loss,
            "e2e": {"value": seqs / (u8_ms * 1e-3), "unit": UNIT, "ms_per_step": u8_ms, "h2d_bytes_per_step": u8_bytes, "d2h_bytes_per_step": 4,
                    "api": "hulc_b200.models.hulc.Hulc.training_step (CUDA-graph replay per staging slot) + fused Adam; double-buffered pinned-host uploads on a copy "
                           "stream; frames handed over as the uint8 the dataset stores, (x/255-0.5)/0.5 runs on the device (hulc_frames_u8_to_f32)",
                    "host_threads_bound_to_gpu_numa_cpus": numa_cpus},
            "e2e_fp32_frames": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                                "note": "same API fed with the reference's batch contract (frames already normalised to fp32 on the host): 4x the PCIe bytes"},
            "e2e_lightning_contract": lc,
            "roofline": roof, "cpu_baseline": cpu, "eager_b200": eager, "latency": lat,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-torch-on-this-GPU comparator")
    ap.add_argument("--no-latency", action="store_true", help="skip the batch-1 inference latency block")
    ap.add_argument("--config", default="hulc", choices=sorted(CONFIGS))
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
