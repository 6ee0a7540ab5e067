#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel, the slowest
launches in order of appearance.   python scripts/launch_list_summary.py gpurun_out/launches.csv [first_n]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else None
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in data:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("rnb::<unnamed>::", "").replace("rnb::", "")
        v, u = float(r[mv].replace(",", "")), r[mu]
        seq.append((name, v / 1000 if u.startswith("n") else (v if u.startswith("u") else v * 1000)))
    if limit:
        seq = seq[:limit]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for n, us in seq:
        tot[n][0] += 1
        tot[n][1] += us
    total = sum(us for _, us in seq)
    print(f"{len(seq)} launches, {total / 1000:.3f} ms of kernel time (ncu: serialised, cold cache - compare shares)")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1000:9.3f} ms {100 * t / total:5.1f} %  {n:5d} launches  avg {t / n:9.2f} us  {k}")
    print("slowest 20 launches:")
    for i, (n, us) in sorted(enumerate(seq), key=lambda x: -x[1][1])[:20]:
        print(f"  #{i:4d} {us / 1000:8.3f} ms  {n}")


if __name__ == "__main__":
    main()
