"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r".*::", "", name).strip()
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        if unit in ("ns", "nsecond"):
            v /= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e3
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot / 1e3:.2f} ms summed kernel time "
          "(ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes)")
    print(f"{'us':>10} {'share':>6} {'n':>5} {'avg us':>8}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} {100 * t / tot:5.1f}% {n:5d} {t / n:8.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
