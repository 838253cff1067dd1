"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv`):
`python bench/launch_summary.py X.csv [out.csv]` prints kernel, launches, total ms, average us, share."""
import csv
import sys
from collections import OrderedDict


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    agg = OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"][:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = ["kernel,launches,total_ms,avg_us,share_of_listed"]
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('"%s",%d,%.3f,%.1f,%.4f' % (name, cnt, ms, ms * 1e3 / cnt, ms / tot))
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
