"""profiles/r01_traffic.json from the per-kernel summary of an `ncu --set full` capture (scripts/ncu_summary.py): for each kernel
bench.py reports a roofline for, the DRAM bytes (read + write) of its largest launch.
    python scripts/make_traffic.py profiles/r01_ncu_full_top_kernels.json profiles/r01_traffic.json "<capture command>" """
import json
import sys

recs = json.load(open(sys.argv[1]))
for i, r in enumerate(recs):
    r["order"] = i

def pick(pred, which=0):
    """launches matching pred whose duration is within 25 % of the longest one (the 200x200 camera / full-size product), in launch order"""
    m = [r for r in recs if pred(r["kernel"])]
    if not m:
        return None
    top = max(r["duration_us"] for r in m)
    big = [r for r in m if r["duration_us"] > 0.75 * top]
    return big[min(which, len(big) - 1)]

want = {
    "conv1_fwd": pick(lambda k: k.startswith("conv1_view_fwd_kernel")),
    "conv1_wgrad": pick(lambda k: k.startswith("conv1_band_wgrad_kernel")),
    "conv2_fwd": pick(lambda k: k.startswith("conv_band_kernel<64, 16, 2, 2>")),
    "conv3_fwd": pick(lambda k: k.startswith("conv_band_kernel<64, 18, 3, 3>"), 0),
    "conv3_dgrad": pick(lambda k: k.startswith("conv_band_kernel<64, 18, 3, 3>"), 1),
    "conv2_dgrad": pick(lambda k: k.startswith("conv_band_kernel<128, 8, 2, 2>")),
    "conv3_wgrad": pick(lambda k: k.startswith("conv_tc_kernel<64, WgradXLoader<64, 3, 1, 0>")),
    "conv2_wgrad": pick(lambda k: k.startswith("conv_tc_kernel<64, WgradXLoader<32, 4, 2, 0>")),
    "rnn_seq_fwd_32steps": pick(lambda k: k.startswith("rnn_seq_kernel<0>")),
    "rnn_seq_bwd_32steps": pick(lambda k: k.startswith("rnn_seq_kernel<1>")),
    "dense_wgrad_2048^3": pick(lambda k: k.startswith("gemm_tc_kernel<128, 0, 1, tc::MNMajorLoader<128>, tc::MNMajorLoader<128>")),
    "dense_fwd_2048^3": pick(lambda k: k.startswith("gemm_tc_kernel<128, 1, 1, tc::KMajorLoader<128>, tc::KMajorLoader<128>")),
}
out = {"source": sys.argv[3] if len(sys.argv) > 3 else "ncu --set full", "kernels": {}}
for name, r in want.items():
    if r is None:
        continue
    out["kernels"][name] = {"dram_bytes_per_launch": r["dram_read_bytes"] + r["dram_write_bytes"], "kernel": r["kernel"], "duration_us_under_ncu": r["duration_us"],
                            "l2_pct": r.get("lts_pct"), "tensor_pct_of_tf32_peak": r.get("tens_pct"), "sm_pct": r.get("sm_pct")}
    print(f"{name:22s} {r['kernel'][:60]:60s} {r['duration_us']:8.1f} us  {out['kernels'][name]['dram_bytes_per_launch'] / 1e6:8.1f} MB  tensor {r.get('tens_pct')}")
json.dump(out, open(sys.argv[2], "w"), indent=1)
