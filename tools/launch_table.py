"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel.

    python tools/launch_table.py gpurun_out/launches.csv [launches_per_step] [out.txt]
With launches_per_step the LAST complete step of the list is aggregated (steady state)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
per = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
n_all = len(rows)
if per:
    rows = rows[-per:]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")[:64]
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
out = ["# %s: %d launches in the list; table over %d launches, %.1f us (ncu per-launch times: serialised, cold cache)" % (
    path.split("/")[-1], n_all, len(rows), tot)]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-66s %5d %10.1f us %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
text = "\n".join(out) + "\n"
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text)
print(text)
