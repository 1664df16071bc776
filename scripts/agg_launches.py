"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

usage: python scripts/agg_launches.py launches.csv [top_n]
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'\(anonymous namespace\)::', '', name)
    m = re.match(r'([\w:]+)(<[^(]*>)?', name)
    return (m.group(1) + (m.group(2) or '')) if m else name


def main():
    rows = list(csv.reader(open(sys.argv[1], errors='replace')))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hi]
    kn, mv = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(',', ''))
        except ValueError:
            continue
        a = agg[short(r[kn])]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    print(f"{'kernel':70s} {'n':>6s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[:70]:70s} {v[0]:6d} {v[1] / 1e3:10.1f} {v[1] / 1e3 / v[0]:8.1f} {100 * v[1] / tot:5.1f}%")
    print(f"total_us {tot / 1e3:.1f} over {sum(v[0] for v in agg.values())} launches")


if __name__ == '__main__':
    main()
