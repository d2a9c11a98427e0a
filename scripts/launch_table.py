"""Per-kernel table from an ncu --csv launch list (metrics as columns): usage launch_table.py file.csv [first_id]"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = OrderedDict()
for r in rows[h + 1:]:
    if len(r) < len(H):
        continue
    rec = dict(zip(H, r))
    k = int(rec["ID"])
    d.setdefault(k, {"name": rec["Kernel Name"].split("(")[0][-34:], "grid": rec.get("Grid Size", "")})
    d[k][rec["Metric Name"]] = float(rec["Metric Value"].replace(",", "")) if rec["Metric Value"] not in ("", "n/a") else 0.0
short = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "rdMB", "dram__bytes_write.sum": "wrMB",
         "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma%", "smsp__issue_active.avg.pct_of_peak_sustained_active": "iss%",
         "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%", "lts__t_sector_hit_rate.pct": "L2hit%",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%"}
tot = {}
for k, v in d.items():
    if k < first:
        continue
    cells = []
    for m, s in short.items():
        if m in v:
            x = v[m]
            if s == "us":
                x /= 1e3
            if s.endswith("MB"):
                x /= 1e6
            cells.append(f"{s}={x:9.1f}")
            if s in ("us", "rdMB", "wrMB"):
                tot.setdefault(v["name"], {}).setdefault(s, 0.0)
                tot[v["name"]][s] += x
    print(f"{k:5d} {v['name']:36s} {v['grid']:>14s} " + " ".join(cells))
print("\nper kernel name:")
T = sum(t.get("us", 0) for t in tot.values())
for n, t in sorted(tot.items(), key=lambda kv: -kv[1].get("us", 0)):
    print(f"  {n:36s} us={t.get('us', 0):9.1f} ({100 * t.get('us', 0) / T:4.1f} %)  rdMB={t.get('rdMB', 0):9.1f} wrMB={t.get('wrMB', 0):9.1f}")
print(f"  total us={T:.1f}")
