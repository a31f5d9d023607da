"""Top source lines by warp-stall samples from `ncu -i rep --page source --csv --print-source sass,cuda`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
funcs, cur, hdr = [], None, None
for r in rows:
    if len(r) >= 2 and r[0] == 'Function Name':
        cur = {'name': r[1], 'lines': []}
        funcs.append(cur)
        continue
    if len(r) >= 2 and r[0] == 'Line No':
        hdr = r
        continue
    if cur is None or len(r) < 10 or r[0] == '' or r[0] == 'File Path':
        continue
    cur['lines'].append(r)
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
merged = {}
for f in funcs:
    key = f['name'][:70]
    merged.setdefault(key, []).extend(f['lines'])
for name, lines in merged.items():
    L = [r for r in lines if r[4].isdigit()]
    tot = sum(int(r[4]) for r in L) or 1
    print('=====', name, 'samples', tot)
    for r in sorted(L, key=lambda r: -int(r[4]))[:topn]:
        print(r[0].rjust(4), '%5.1f%%' % (100 * int(r[4]) / tot), 'inst', r[7].rjust(9), 'long', r[ix['stall_long_sb']].rjust(6),
              'short', r[ix['stall_short_sb']].rjust(5), 'wait', r[ix['stall_wait']].rjust(5), 'bar', r[ix['stall_barrier']].rjust(5),
              'mio', r[ix['stall_mio']].rjust(5), 'lg', r[ix['stall_lg']].rjust(4), '|', r[1].strip()[:80])
