"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`)
per kernel: share of the step, launches, DRAM bytes.  usage: summarize_launches.py launches.csv [gemm_traffic.json]"""
import collections
import csv
import json
import re
import sys


def tobytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def main(path, traffic_json=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r".*::", "", name).strip()
        m, u, a = row["Metric Name"], row["Metric Unit"], agg[name]
        if m == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", ""))
            if u in ("ns", "nsecond"):
                v /= 1e3
            elif u in ("ms", "msecond"):
                v *= 1e3
            a["us"] += v
            a["n"] += 1
        elif m == "dram__bytes_read.sum":
            a["rd"] += tobytes(row["Metric Value"], u)
        elif m == "dram__bytes_write.sum":
            a["wr"] += tobytes(row["Metric Value"], u)
    tot = sum(a["us"] for a in agg.values())
    print(f"# {path}: {sum(a['n'] for a in agg.values())} launches of ONE attack iteration, {tot / 1e3:.2f} ms summed kernel time "
          "(ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes)")
    print(f"{'us':>10} {'share':>6} {'n':>5} {'avg us':>8} {'dram rd MB':>11} {'dram wr MB':>11} {'GB/s':>7}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{a['us']:10.1f} {100 * a['us'] / tot:5.1f}% {a['n']:5d} {a['us'] / a['n']:8.1f} {a['rd'] / 1e6:11.1f} {a['wr'] / 1e6:11.1f} "
              f"{(a['rd'] + a['wr']) / a['us'] / 1e3:7.0f}  {k}")
    if traffic_json:
        g = [a for k, a in agg.items() if k.startswith("gemm_bf16_tn_kernel")]
        out = {"gemm_launches": sum(a["n"] for a in g), "dram_read_bytes_per_step": sum(a["rd"] for a in g),
               "dram_write_bytes_per_step": sum(a["wr"] for a in g)}
        out["dram_bytes_per_step"] = out["dram_read_bytes_per_step"] + out["dram_write_bytes_per_step"]
        out["source"] = f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch of one attack iteration ({path})"
        json.dump(out, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:3])
