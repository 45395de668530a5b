"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` (stdin)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) > 5 and r[0].startswith("0x")]
si = hdr.index("# Samples")
def f(x):
    try: return float(x)
    except Exception: return 0.0
tot = sum(f(r[si]) for r in data)
print("total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda i: -f(data[i][si]))
for i in order[:int(sys.argv[1]) if len(sys.argv) > 1 else 16]:
    print(f"{i:5d} {f(data[i][si]):8.0f} {100*f(data[i][si])/max(tot,1):5.1f}%  {data[i][1].strip()[:120]}")
