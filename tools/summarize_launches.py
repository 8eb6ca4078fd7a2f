"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections
import csv
import re
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "second": 1e9}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    lines = ["%-44s %7s %11s %7s %9s" % ("kernel", "count", "total ms", "share", "avg us")]
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append("%-44s %7d %11.3f %6.1f%% %9.1f" % (k[:44], c, t / 1e6, 100 * t / tot, t / c / 1e3))
    lines.append("%-44s %7d %11.3f" % ("TOTAL", sum(c for c, _ in agg.values()), tot / 1e6))
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
