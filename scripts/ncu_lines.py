"""Join an ncu SASS source page (per-instruction counters) with nvdisasm line info -> per-source-line totals.

usage: python scripts/ncu_lines.py report.ncu-rep file.cubin kernel_substring [top_n]
(cubins: `cuobjdump -xelf all fastmot_b200/libfastmot_b200.so`)
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, cubin, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
    h = rows[hi]
    ie, ss, si = h.index('Instructions Executed'), h.index('# Samples'), h.index('Source')
    inst = [(r[si].strip(), int(r[ie]), int(r[ss])) for r in rows[hi + 1:] if len(r) > ie and r[ie].isdigit()]
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    lines, cur, infn, k = [], None, False, 0
    for ln in dis.splitlines():
        if ln.startswith('.text.'):
            infn = kern in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            lines.append(cur)
    print(f'{len(inst)} counted instructions, {len(lines)} disassembled', file=sys.stderr)
    n = min(len(inst), len(lines))
    agg = collections.defaultdict(lambda: [0, 0])
    for (src, cnt, smp), loc in zip(inst[:n], lines[:n]):
        agg[loc][0] += cnt
        agg[loc][1] += smp
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    srcs = {}
    for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ''
        if loc:
            try:
                if loc[0] not in srcs:
                    srcs[loc[0]] = open('fastmot_b200/csrc/' + loc[0]).read().splitlines()
                text = srcs[loc[0]][loc[1] - 1].strip()
            except Exception:
                pass
        print(f'{v[0]:11d} {100 * v[0] / tot:5.1f}% inst  {100 * v[1] / tots:5.1f}% samples  {loc}: {text[:100]}')


if __name__ == '__main__':
    main()
