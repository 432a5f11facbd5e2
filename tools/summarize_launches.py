"""Summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches and summed device
time per kernel name.  python tools/summarize_launches.py LAUNCHES.csv "title" > SUMMARY.txt"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    name = re.sub(r"<.*", "", name); name = re.sub(r"\(.*", "", name).split("::")[-1]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    rows.append((name, ms))
tot = sum(ms for _, ms in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ms in rows:
    agg[n][0] += 1; agg[n][1] += ms
print("%d launches (%s), gpu__time_duration.sum total %.1f ms (serialised, cold cache: compare SHARES)" % (
    len(rows), sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], tot))
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s %5d launches %10.3f ms %5.1f %%" % (n, c, ms, 100 * ms / max(tot, 1e-12)))
