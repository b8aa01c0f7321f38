"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  usage: python profiles/summarize_launches.py file.csv [n]"""
import collections
import csv
import re
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r'\(.*', '', r[ki]).replace('void <unnamed>::', '').replace('<unnamed>::', '').replace('void ', '')
        v = float(r[vi].replace(',', ''))
        if r[ui] in ('ns', 'nsecond'):
            v /= 1e3
        elif r[ui] in ('ms', 'msecond'):
            v *= 1e3
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f'total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{v[1]:12.1f} us {100 * v[1] / tot:6.2f}%  n={v[0]:4d}  avg={v[1] / v[0]:10.1f} us  {k[:100]}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
