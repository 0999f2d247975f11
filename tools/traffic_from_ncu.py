"""Dev tool: DRAM traffic of the predict solve's GEMM launches from an ncu metrics pass.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv -k regex:dgemm python tools/prof_factorize.py N 1 KIND
  python tools/traffic_from_ncu.py X.csv <launches of one predict solve> <key> [profiles/traffic.json]

The last <launches> dgemm launches of the capture are one predict solve (prof_factorize.py ends with a predict); prints the per-launch
average and total and merges them into profiles/traffic.json under <key> (what bench.py reports as roofline.traffic)."""
import collections
import csv
import io
import json
import sys

path, n_solve, key = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = sys.argv[4] if len(sys.argv) > 4 else "profiles/traffic.json"
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.OrderedDict()
for row in csv.DictReader(io.StringIO("".join(lines))):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"].lower()
    v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    per[row["ID"]] = per.get(row["ID"], 0.0) + v
vals = list(per.values())[-n_solve:]
tot = sum(vals)
rec = {"dram_bytes_per_launch": tot / len(vals), "launches": len(vals), "dram_bytes_total": tot,
       "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:dgemm, last {len(vals)} launches = one predict solve ({path})"}
print(json.dumps(rec))
try:
    d = json.load(open(out))
except Exception:
    d = {}
d[key] = rec
json.dump(d, open(out, "w"), indent=2)
