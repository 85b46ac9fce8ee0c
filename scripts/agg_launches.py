"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel and by launch shape."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
if "--last" in sys.argv:  # keep the last N launches (= the last forward when warm-up forwards were captured too)
    rows = rows[-int(sys.argv[sys.argv.index("--last") + 1]):]
agg = collections.OrderedDict()
for row in rows:
    nm = re.sub(r"\(.*", "", row["Kernel Name"])
    nm = re.sub(r"^void\s+", "", nm).replace("tbn::", "")[:60]
    t = float(row["Metric Value"])
    a = agg.setdefault(nm, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, total {tot/1e3:.1f} us (serialised, cold-cache: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[0]:5d} launches {v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%  {k}")
if "--top" in sys.argv:
    print("\ntop launches:")
    for row in sorted(rows, key=lambda r: -float(r["Metric Value"]))[:25]:
        nm = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("tbn::", "")[:40]
        print(f"  id {row['ID']:>5} {float(row['Metric Value'])/1e3:9.1f} us grid {row['Grid Size']:>16} block {row['Block Size']:>12} {nm}")
