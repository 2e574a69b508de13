"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time per kernel name.
usage: python scripts/launch_summary.py launches.csv [skip_first_n_launches]"""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
tot = sum(v for _, v in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    agg[k][0] += 1; agg[k][1] += v
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {v/1e3:9.3f} ms {100*v/tot:5.1f}%  x{n:<4d} {k[:110]}")
