"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (mean / total / share)."""
import collections
import csv
import re
import sys


def main(path, skip_first=0):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1 + skip_first:]:
        if len(r) <= vi:
            continue
        name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('<unnamed>::', '')
        v = float(r[vi].replace(',', ''))
        v = v / 1000.0 if r[ui] == 'ns' else (v * 1000.0 if r[ui] == 'ms' else v)
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f'total {tot/1000:.3f} ms over {sum(v[0] for v in agg.values())} launches')
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
        print(f'{k[:72]:72s} n={n:5d} total={t/1000:9.3f} ms  mean={t/n:8.2f} us share={t/tot:.3f}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
