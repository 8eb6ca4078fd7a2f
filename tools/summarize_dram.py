"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log:
per kernel launches, time, DRAM bytes and achieved DRAM GB/s (cold-cache, serialised launches)."""
import collections
import csv
import re
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = collections.defaultdict(dict)
    names = {}
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[row["ID"]][row["Metric Name"]] = v * scale
        names[row["ID"]] = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for i, m in per.items():
        a = agg[names[i]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    rows = ["%-40s %6s %10s %7s %10s %9s" % ("kernel", "count", "total ms", "share", "DRAM MB", "GB/s")]
    for k, (c, t, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
        rows.append("%-40s %6d %10.3f %6.1f%% %10.1f %9.0f" % (k[:40], c, t * 1e3, 100 * t / tot, b / 1e6, b / t / 1e9 if t else 0))
    text = "\n".join(rows)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
