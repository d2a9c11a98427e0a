"""Writes profiles/roofline_traffic.json from an ncu launch list of ONE run (metrics gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum): DRAM bytes of the accumulation stage of the LAST commit in the list,
stamped with the hash of the CUDA sources so that bench.py never prints a figure measured on other kernels.
usage: python scripts/update_traffic.py gpurun_out/launches.csv cfg1_1024x1024 16 "command that produced the list" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

path, workload, bits, cmd = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(path)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
launches = {}
for r in rows[h + 1:]:
    if len(r) < len(H):
        continue
    rec = dict(zip(H, r))
    k = int(rec["ID"])
    launches.setdefault(k, {"name": rec["Kernel Name"].split("(")[0]})
    mv = rec["Metric Value"].replace(",", "")
    launches[k][rec["Metric Name"]] = float(mv) if mv not in ("", "n/a") else 0.0
ids = sorted(launches)
norm = [i for i in ids if "k_normalize" in launches[i]["name"]]
end = norm[-1]
start = norm[-2] if len(norm) > 1 else ids[0]
STAGE = ("k_bat_prefix", "k_bat_finish", "k_ba_prefix", "k_ba_finish", "k_ba_invert", "k_mult_sum_rows", "k_accumulate")
tot, us, ids_used = 0.0, 0.0, []
for i in ids:
    if start < i < end and any(s in launches[i]["name"] for s in STAGE):
        tot += launches[i].get("dram__bytes_read.sum", 0) + launches[i].get("dram__bytes_write.sum", 0)
        us += launches[i].get("gpu__time_duration.sum", 0) / 1e3
        ids_used.append(i)
out = os.path.join(ROOT, "profiles", "roofline_traffic.json")
try:
    data = json.load(open(out))
except Exception:
    data = {}
data[workload] = {
    "accumulate_stage_dram_bytes_per_commit": int(tot),
    "accumulate_stage_us_serialised": us,
    "table_window_bits": bits,
    "source_sha16": bench.source_sha16(),
    "source": f"{os.path.relpath(path, ROOT)}: dram__bytes_read.sum + dram__bytes_write.sum of the {len(ids_used)} stage launches "
              f"(IDs {ids_used[0]}-{ids_used[-1]}) of the last commit in `{cmd}`",
}
json.dump(data, open(out, "w"), indent=1)
print(json.dumps(data[workload], indent=1))
