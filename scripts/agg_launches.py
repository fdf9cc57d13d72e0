"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel for the LAST training step in it.
    python scripts/agg_launches.py gpurun_out/launches.csv [top_n]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.DictReader([l for l in open(path) if l.startswith('"')]))
adam = [i for i, x in enumerate(rows) if "adam" in x["Kernel Name"].lower()]
s0 = adam[-2] + 1 if len(adam) > 1 else 0
s1 = adam[-1] + 1 if adam else len(rows)
step = rows[s0:s1]
agg = collections.defaultdict(lambda: [0, 0.0])
for x in step:
    k = re.sub(r"^void ", "", re.sub(r"\(.*", "", x["Kernel Name"])).replace("<unnamed>::", "")
    agg[k][0] += 1
    agg[k][1] += float(x["Metric Value"])
tot = sum(v[1] for v in agg.values())
print(f"step: {len(step)} launches, {tot / 1e6:.3f} ms summed kernel time (serialised, cold caches)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1] / 1e3:9.1f} us {v[0]:4d}x {100 * v[1] / tot:5.1f}%  {k[:140]}")
