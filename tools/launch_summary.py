"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (dev tool)."""
import collections, csv, io, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(io.StringIO("".join(lines))):
    name = row["Kernel Name"][:70]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    a = agg.setdefault(name, [0, 0.0, 1e30, 0.0]); a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':72s} {'n':>6s} {'total ms':>10s} {'avg us':>9s} {'min us':>9s} {'max us':>9s} {'share':>6s}")
for k, (c, t, mn, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {c:6d} {t/1e3:10.3f} {t/c:9.2f} {mn:9.2f} {mx:9.2f} {t/tot*100:5.1f}%")
