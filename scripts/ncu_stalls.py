"""Top stall locations (SASS) of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-id ... > f.csv`.
    python scripts/ncu_stalls.py f.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == "Address"), len(rows))
body = rows[hdr_i + 1 : end]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in body:
    try:
        n = int(r[col["# Samples"]])
    except Exception:
        continue
    why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    data.append((n, r[col["Source"]], why, r[col["Instructions Executed"]]))
tot = sum(d[0] for d in data)
print("total samples", tot, " instructions", len(data))
for n, sass, why, ex in sorted(data, key=lambda x: -x[0])[:top]:
    print(f"{n:7d} {100 * n / tot:5.1f}%  {sass[:90]:90s} {why} exec={ex}")
