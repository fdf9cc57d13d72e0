"""Per-kernel summary (duration, DRAM bytes, L2 / tensor / SM utilisation) of an `ncu --set full` capture exported with
`ncu -i X.ncu-rep --page raw --csv > raw.csv`, written as JSON (profiles/) and printed.
    python scripts/ncu_summary.py raw.csv out.json"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
c = hdr.index
M = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "tens": "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed"}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
out = []
tens_cols = [i for i, h in enumerate(hdr) if "utchmma" in h.lower() and h.endswith("pct_of_peak_sustained_elapsed")]  # tcgen05.mma paths (tf32, bf16, ...)
pipe_col = next((i for i, h in enumerate(hdr) if h == "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"), None)
for d in data:
    name = re.sub(r"\(.*", "", re.sub(r"^void ", "", d[c("Kernel Name")]).replace("<unnamed>::", ""))
    rec = {"kernel": name, "grid": d[c("Grid Size")], "block": d[c("Block Size")]}
    rec["duration_us"] = float(d[c(M["dur"])]) * scale[units[c(M["dur"])]]
    rec["dram_read_bytes"] = float(d[c(M["rd"])]) * scale[units[c(M["rd"])]]
    rec["dram_write_bytes"] = float(d[c(M["wr"])]) * scale[units[c(M["wr"])]]
    for k in ("lts", "tens", "sm"):
        try:
            rec[k + "_pct"] = float(d[c(M[k])])
        except Exception:
            rec[k + "_pct"] = None
    best = None
    for i in tens_cols:  # the instruction path this kernel actually uses: the largest of the per-type tensor-op rates
        try:
            v = float(d[i])
            if best is None or v > best[0]:
                best = (v, hdr[i])
        except Exception:
            pass
    if best is not None:
        rec["tens_pct"], rec["tens_metric"] = best
    if pipe_col is not None:
        try:
            rec["tensor_pipe_cycles_active_pct"] = float(d[pipe_col])
        except Exception:
            pass
    out.append(rec)
    print(f"{name[:66]:66s} {rec['grid']:12s} {rec['duration_us']:8.1f}us rd {rec['dram_read_bytes'] / 1e6:8.1f}MB wr {rec['dram_write_bytes'] / 1e6:7.1f}MB "
          f"L2 {rec['lts_pct']:.0f}% tensor {(rec['tens_pct'] or 0):.1f}% sm {rec['sm_pct']:.0f}%")
json.dump(out, open(sys.argv[2], "w"), indent=1)
