#!/usr/bin/env python
"""Hottest SASS instructions of a kernel from `ncu -i rep --page source --csv` (stdin or file)."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith('0x')] or rows[hi + 1:]
def num(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0
tot = sum(num(r, '# Samples') for r in body)
tot_inst = sum(num(r, 'Instructions Executed') for r in body)
print(f'total samples {tot:.0f}, warp instructions executed {tot_inst:.0f}')
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(num(r, h) for r in body) for h in stall_cols}
print('stall totals:', ', '.join(f'{h[6:]}={v / max(tot, 1):.1%}' for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v / max(tot, 1) > 0.01))
order = sorted(range(len(body)), key=lambda i: -num(body[i], '# Samples'))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((num(r, h), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f'{i:5d} {num(r, "# Samples") / max(tot, 1):6.2%} inst={num(r, "Instructions Executed"):9.0f}  {r[ix["Source"]][:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}')
