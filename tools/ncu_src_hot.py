"""Stall-sample hot spots of one kernel launch in an ncu report (needs --import-source on / -lineinfo)."""
import csv
import subprocess
import sys


def main(rep, kernel_regex, skip=0, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel_regex, '--launch-skip',
                          str(skip), '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1][:110])
    hdr = rows[1]
    si, src = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source')
    data = [r for r in rows[2:] if len(r) > si and r[si].strip().isdigit()]
    tot = sum(int(r[si]) for r in data)
    print('total samples', tot, 'instructions', len(data))
    for s, i, t in sorted(((int(r[si]), i, r[src].strip()) for i, r in enumerate(data)), reverse=True)[:top]:
        print(f'{s:5d} ({100*s/tot:4.1f}%) @{i:5d} {t[:90]}')
    print('--- windows of 200 instructions')
    for w in range(0, len(data), 200):
        s = sum(int(r[si]) for r in data[w:w + 200])
        print(f'{w:5d} {s:5d} ({100*s/tot:4.1f}%)  {data[w][src].strip()[:50]}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
