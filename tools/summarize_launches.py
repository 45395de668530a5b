"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import csv, sys, re, collections
path = sys.argv[1]; last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^.*::", "", name)
        rows.append((name, v))
if last:
    rows = rows[-last:]
tot = sum(v for _, v in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1; agg[n][1] += v
print(f"launches {len(rows)}  total {tot/1e3:.3f} ms")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:45s} n={c:5d} total={v/1e3:9.3f} ms  avg={v/c:9.2f} us  share={100*v/tot:5.1f}%")
