"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/xyz_summary.txt"""
import collections
import csv
import sys


def main(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(rows)
    hdr = next(rd)
    c = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    total = 0.0
    for r in rd:
        if r[c["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[c["Metric Value"]].replace(",", ""))
        unit = r[c["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        t, n = agg.get(r[c["Kernel Name"]], (0.0, 0))
        agg[r[c["Kernel Name"]]] = (t + us, n + 1)
        total += us
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{t / 1e3:9.3f} ms  n={n:4d}  avg={t / n:8.1f} us  {100 * t / total:5.1f}%  {name[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
