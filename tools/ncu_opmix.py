"""Opcode mix (executed warp instructions) of a SASS range of the first kernel in an .ncu-rep."""
import csv, subprocess, sys, collections
rep, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[start]; ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[start + 1:] if len(r) == len(h)]
mix = collections.Counter(); tot = 0
for n, r in enumerate(body):
    if lo <= n <= hi:
        src = r[ix['Source']].strip().split()
        op = src[1] if src[0].startswith('@') else src[0]
        op = op.split('.')[0]
        c = int(r[ix['Instructions Executed']] or 0)
        mix[op] += c; tot += c
print('total', tot)
for op, c in mix.most_common(25):
    print(f'  {op:10s} {c:10d}  {100*c/tot:5.1f}%')
if len(sys.argv) > 4:
    for n, r in enumerate(body):
        if lo <= n <= hi:
            print(n, r[ix['Source']].strip()[:70], r[ix['Instructions Executed']])
