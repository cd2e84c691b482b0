#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` launch list: per kernel name, launches, mean duration and
mean of every other metric; optionally per-launch lines for the first N launches."""
import collections
import csv
import sys


def main(path, show=0):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        key = (int(r[idx['ID']]), r[idx['Kernel Name']].split('(')[0][:48], r[idx['Grid Size']])
        per.setdefault(key, {})[r[idx['Metric Name']]] = float(r[idx['Metric Value']].replace(',', ''))
    agg = collections.OrderedDict()
    for (i, k, g), m in per.items():
        if i < show:
            print(i, k, g, {a: round(v / (1e3 if 'time' in a else 1e6), 2) for a, v in m.items()})
        agg.setdefault((k, g), []).append(m)
    total = sum(m.get('gpu__time_duration.sum', 0) for ms in agg.values() for m in ms)
    print(f'{"kernel":50s} {"grid":>14s} {"n":>4s} {"avg us":>8s} {"share":>6s}  other metrics (avg, MB)')
    for (k, g), ms in agg.items():
        names = sorted(ms[0])
        avg = {n: sum(m.get(n, 0) for m in ms) / len(ms) for n in names}
        t = avg.get('gpu__time_duration.sum', 0)
        other = ' '.join(f'{n.split("__")[-1].replace(".sum","")}={avg[n]/1e6:.1f}' for n in names if 'time' not in n)
        print(f'{k:50s} {g:>14s} {len(ms):4d} {t/1e3:8.1f} {t*len(ms)/total:6.1%}  {other}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
