"""Per-kernel launch counts and GPU-time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_shares.py profiles/r2_i_launches.csv > profiles/r2_i_launch_shares.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [l for l in open(path) if l.startswith('"')]
    total = defaultdict(int)
    count = defaultdict(int)
    for r in csv.DictReader(rows):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:100]
        total[name] += int(float(r["Metric Value"]))
        count[name] += 1
    whole = sum(total.values())
    print("# %s: %d launches, %.1f ms of GPU time (per-launch times are cold-cache and serialised)" % (path, sum(count.values()), whole / 1e6))
    for name in sorted(total, key=total.get, reverse=True):
        print("%-100s %6d %14d ns %5.1f %%" % (name, count[name], total[name], 100.0 * total[name] / whole))


if __name__ == "__main__":
    main(sys.argv[1])
