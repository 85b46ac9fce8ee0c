"""Post-process `ncu --csv` metric logs (scripts/ncu_evidence.sh) into the per-launch tables committed under profiles/.

    python scripts/ncu_to_csv.py step  gpurun_out/ncu_step_g0.csv  profiles/r2_ncu_step_per_modality.csv
    python scripts/ncu_to_csv.py mem   gpurun_out/ncu_mem_b256.csv profiles/r2_ncu_memory_kernels_b256.txt
"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict


def read(path):
    rows = OrderedDict()
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        d = rows.setdefault(k, {"kernel": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        elif name.startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1)
        d[name] = v
    return rows


def short(k):
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"tbn::", "", k)
    return re.sub(r"\(.*", "", k)


mode, src, dst = sys.argv[1:4]
rows = read(src)
if mode == "step":
    with open(dst, "w") as f:
        f.write("id,kernel,grid,block,dram_read_bytes,dram_write_bytes,dram_bytes,time_us\n")
        for k, d in rows.items():
            rd, wr = d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
            f.write(f'{k},"{short(d["kernel"])}","{d["grid"]}","{d["block"]}",{rd:.0f},{wr:.0f},{rd + wr:.0f},{d.get("gpu__time_duration.sum", 0):.2f}\n')
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for d in rows.values():
        a = agg[short(d["kernel"])]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0)
        a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    tot = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {tot:.1f} us serialised (cold cache: compare shares)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {a[0]:4d} x {k:44s} {a[1]:8.1f} us {100 * a[1] / tot:5.1f} %   DRAM {a[2] / 1e6:8.1f} MB  ({a[2] / a[0] / 1e6:6.2f} MB/launch, {a[2] / a[1] / 1e3 if a[1] else 0:6.0f} GB/s)")
else:
    peak = float(sys.argv[4]) if len(sys.argv) > 4 else 6555.8
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    out = [f"# {src}: memory-bound kernels, ncu dram__bytes_read+write / gpu__time_duration per launch, against the measured HBM peak {peak} GB/s",
           f"# {'kernel':52s} {'grid':>14s} {'time us':>9s} {'DRAM MB':>9s} {'GB/s':>8s} {'of peak':>8s}"]
    for k, d in rows.items():
        b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        t = d.get("gpu__time_duration.sum", 0)
        a = agg[short(d["kernel"])]
        a[0] += 1; a[1] += t; a[2] += b
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = a[2] / a[1] / 1e3 if a[1] else 0
        out.append(f"  {a[0]:4d} x {k:46s} {'':>14s} {a[1] / a[0]:9.1f} {a[2] / a[0] / 1e6:9.2f} {gbs:8.0f} {gbs / peak:8.1%}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))
