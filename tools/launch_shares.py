#!/usr/bin/env python
"""Per-kernel share of the device time in an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`):
    python tools/launch_shares.py profiles/r02_launches_bench.csv > profiles/r02_launches_bench_shares.txt"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    head = rows[0]
    k_name, k_val, k_unit = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= k_val or "gpu__time_duration" not in ",".join(r):
            continue
        v = float(r[k_val].replace(",", ""))
        v = v / 1e3 if r[k_unit] in ("ns", "nsecond") else (v * 1e3 if r[k_unit] in ("ms", "msecond") else v)   # -> us
        name = r[k_name][:70]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# per-kernel share of the device time in {path} (cold-cache, serialised per-launch times: compare SHARES)")
    print("# kernel, launches, total us, share")
    for name, v in tot.most_common(14):
        print(f"{name}, {cnt[name]}, {v:.1f}, {v / total:.3f}")


if __name__ == "__main__":
    main()
