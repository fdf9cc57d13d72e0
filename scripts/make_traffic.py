"""profiles/r02_traffic.json from the per-kernel summaries of the `ncu --set full` captures (scripts/ncu_summary.py): for each kernel
bench.py reports a roofline for, per precision mode, the DRAM bytes (read + write) of its largest launch.
    python scripts/make_traffic.py        (reads profiles/r02_ncu_full_{bf16,tf32,gemm_bf16}.json)"""
import json
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"


def pick(recs, prefix, which=0):
    """launches of the kernel whose duration is within 25 % of its longest one (the 200x200 camera / full-size product), in launch order"""
    m = [r for r in recs if r["kernel"].startswith(prefix)]
    if not m:
        return None
    top = max(r["duration_us"] for r in m)
    big = [r for r in m if r["duration_us"] > 0.75 * top]
    r = big[min(which, len(big) - 1)]
    return {"kernel": r["kernel"], "duration_us": r["duration_us"], "dram_bytes_per_launch": r["dram_read_bytes"] + r["dram_write_bytes"],
            "dram_read_bytes": r["dram_read_bytes"], "dram_write_bytes": r["dram_write_bytes"], "tensor_pct": r.get("tens_pct"), "l2_pct": r["lts_pct"],
            "sm_pct": r["sm_pct"]}


b, t, g = (json.load(open(P / f"r02_ncu_full_{n}.json")) for n in ("bf16", "tf32", "gemm_bf16"))
out = {"source": "ncu --set full --clock-control none over python scripts/profile_step.py --steps 2 [--precision bf16] on a B200 (round 2); "
                 "summaries: profiles/r02_ncu_full_{bf16,tf32,gemm_bf16}.json"}
out["bf16"] = {
    "conv1_fwd": pick(b, "conv1_view_fwd"), "conv1_wgrad": pick(b, "conv1_band_wgrad"),
    "conv2_fwd": pick(b, "conv_band_kernel<64, 8, 2, 2, 1>"), "conv3_fwd": pick(b, "conv_band_kernel<64, 9, 3, 3, 1>", 0),
    "conv3_dgrad": pick(b, "conv_band_kernel<64, 9, 3, 3, 1>", 1), "conv2_dgrad": pick(b, "conv_band_kernel<128, 4, 2, 2, 1>"),
    "conv3_wgrad": pick(b, "conv_wgrad_bf16_kernel<1>"), "conv2_wgrad": pick(b, "conv_wgrad_bf16_kernel<2>"),
    "rnn_seq_fwd_32steps": pick(b, "rnn_push_kernel<__nv_bfloat16, 0"), "rnn_seq_bwd_32steps": pick(b, "rnn_push_kernel<__nv_bfloat16, 1"),
    "spatial_softmax_fwd": pick(b, "spatial_softmax_vec_kernel<1, 0>") or pick(b, "spatial_softmax_nhwc_reg_kernel<16, 28, 0, 1>"),
    "spatial_softmax_bwd": pick(b, "spatial_softmax_vec_kernel<1, 1>") or pick(b, "spatial_softmax_nhwc_reg_kernel<16, 28, 1, 1>"),
    "adam": pick(b, "adam_kernel"),
}
big = [r for r in g if r["duration_us"] > 30 and (r.get("tens_pct") or 0) > 20]  # the 2048^3-class products of the decoder
if big:
    out["bf16"]["dense_fwd_2048^3"] = pick(big[:1], "gemm_bf16_kernel")
    out["bf16"]["dense_wgrad_2048^3"] = pick(big[1:2] or big[:1], "gemm_bf16_kernel")
out["tf32"] = {
    "conv1_fwd": pick(t, "conv1_view_fwd"), "conv1_wgrad": pick(t, "conv1_band_wgrad"),
    "conv2_fwd": pick(t, "conv_band_kernel<64, 16, 2, 2, 0>"), "conv3_fwd": pick(t, "conv_band_kernel<64, 18, 3, 3, 0>", 0),
    "conv3_dgrad": pick(t, "conv_band_kernel<64, 18, 3, 3, 0>", 1), "conv2_dgrad": pick(t, "conv_band_kernel<128, 8, 2, 2, 0>"),
    "conv3_wgrad": pick(t, "conv_tc_kernel<64, WgradXLoader<64, 3, 1, 0>"), "conv2_wgrad": pick(t, "conv_tc_kernel<64, WgradXLoader<32, 4, 2, 0>"),
    "rnn_seq_fwd_32steps": pick(t, "rnn_seq_kernel<0>"), "rnn_seq_bwd_32steps": pick(t, "rnn_seq_kernel<1>"),
    "spatial_softmax_fwd": pick(t, "spatial_softmax_vec_kernel<0, 0>") or pick(t, "spatial_softmax_nhwc_reg_kernel<16, 28, 0, 0>"),
    "spatial_softmax_bwd": pick(t, "spatial_softmax_vec_kernel<0, 1>") or pick(t, "spatial_softmax_nhwc_reg_kernel<16, 28, 1, 0>"),
    "adam": pick(t, "adam_kernel"),
}
json.dump(out, open(P / "r02_traffic.json", "w"), indent=1)
for mode in ("bf16", "tf32"):
    print(mode, {k: (round(v["dram_bytes_per_launch"] / 1e6, 1) if v else None) for k, v in out[mode].items()})
