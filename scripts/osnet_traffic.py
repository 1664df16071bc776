"""Sums an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` capture of
scripts/time_osnet.py --eager into per-kernel and per-forward DRAM traffic (profiles/r02_osnet_traffic.json feeds
bench.py's roofline.traffic).
usage: python scripts/osnet_traffic.py capture.csv forwards_in_capture out_prefix"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors='replace')))
nfwd = int(sys.argv[2])
out = sys.argv[3]
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]
ki, mi, vi, ui = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
idi = h.index('ID')
per = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
seen = set()
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0,
         'nsecond': 1e-3, 'msecond': 1e3}
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r'^void\s+|\(anonymous namespace\)::|<unnamed>::', '', r[ki])
    name = re.match(r'[\w:]+(<[^(]*>)?', name).group(0)
    if name.startswith('at::'):
        continue                     # torch fill / copy kernels of the harness, not part of the forward
    v = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
    per[name][r[mi]] += v
    if (r[idi], name) not in seen:
        seen.add((r[idi], name))
        cnt[name] += 1
tot_t = tot_r = tot_w = 0.0
lines = [f"{'kernel':58s} {'n/fwd':>6s} {'time_us':>9s} {'rd_MB':>9s} {'wr_MB':>9s} {'GB/s':>8s}"]
for name, m in sorted(per.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
    t, rd, wr = m['gpu__time_duration.sum'] / nfwd, m['dram__bytes_read.sum'] / nfwd, m['dram__bytes_write.sum'] / nfwd
    tot_t += t; tot_r += rd; tot_w += wr
    lines.append(f"{name[:58]:58s} {cnt[name] / nfwd:6.1f} {t:9.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {(rd + wr) / t / 1e3:8.1f}")
lines.append(f"TOTAL per forward: {tot_t / 1e3:.3f} ms under ncu (kernels serialised), DRAM traffic {(tot_r + tot_w) / 1e9:.3f} GB "
             f"(read {tot_r / 1e9:.3f}, write {tot_w / 1e9:.3f})")
open(out + '.txt', 'w').write("\n".join(lines) + "\n")
json.dump({"osnet_forward_dram_bytes": tot_r + tot_w, "read": tot_r, "write": tot_w, "ncu_time_ms": tot_t / 1e3,
           "forwards": nfwd, "source": sys.argv[1]}, open(out + '.json', 'w'), indent=1)
print("\n".join(lines[-6:]))
