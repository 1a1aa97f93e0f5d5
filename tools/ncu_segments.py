"""Summarise an ncu source-page CSV by barrier-delimited SASS segments.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv
    python tools/ncu_segments.py src.csv
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ia, ie, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
body = [r for r in rows[2:] if len(r) > isamp and r[ie].isdigit()]
# keep the first kernel instance only (the file repeats per captured launch)
first = body[0][h.index('Address')]
cut = [i for i, r in enumerate(body) if r[h.index('Address')] == first]
if len(cut) > 1:
    body = body[:cut[1]]
tot = sum(int(r[ie]) for r in body)
tots = sum(int(r[isamp]) for r in body)
print('total warp instructions', tot, 'samples', tots)
seg, acc, accs, n, start, ops = 0, 0, 0, 0, 0, {}
for k, r in enumerate(body):
    s = r[ia].strip()
    e = int(r[ie])
    acc += e
    accs += int(r[isamp])
    n += 1
    t = s.split()
    op = (t[0] if not t[0].startswith('@') else t[1]).split('.')[0]
    ops[op] = ops.get(op, 0) + e
    if 'BAR.SYNC' in s or 'EXIT' in s or k == len(body) - 1:
        top = sorted(ops.items(), key=lambda x: -x[1])[:7]
        if acc * 200 > tot or accs * 200 > tots:
            print(f'seg{seg} sass[{start}:{k}] n={n} inst={acc} ({100 * acc / tot:.1f}%) samples={accs} '
                  f'({100 * accs / tots:.1f}%)', top)
        seg += 1
        acc, accs, n, start, ops = 0, 0, 0, k + 1, {}
