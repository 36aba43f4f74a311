"""Split an .ncu-rep source page at BAR.SYNC instructions: warp instructions, shared wavefronts and stall
samples per region (phases of the tiled assembly kernel)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[start]; ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[start + 1:] if len(r) == len(h)]
gi = lambda r, k: int(r[ix[k]] or 0)
reg = [0, 0, 0, 0]; first = 0
tot = [0, 0, 0, 0]
for n, r in enumerate(body):
    vals = [gi(r, 'Instructions Executed'), gi(r, 'L1 Wavefronts Shared'), gi(r, '# Samples'), gi(r, 'L1 Tag Requests Global')]
    reg = [a + b for a, b in zip(reg, vals)]; tot = [a + b for a, b in zip(tot, vals)]
    src = r[ix['Source']]
    if 'BAR.SYNC' in src or n == len(body) - 1:
        print(f'sass {first:5d}-{n:5d}: inst {reg[0]:10d}  shared wf {reg[1]:10d}  samples {reg[2]:6d}  gtag {reg[3]:9d}')
        reg = [0, 0, 0, 0]; first = n + 1
print('total', tot)
