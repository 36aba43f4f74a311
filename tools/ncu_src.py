"""Per-SASS-instruction LSU cost of the first kernel in an .ncu-rep (source page): shared
wavefronts vs ideal, global tag requests / sectors, stall samples.  Usage: ncu_src.py rep [min_frac]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.015
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[start]
ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[start + 1:] if len(r) == len(h)]
gi = lambda r, k: int(r[ix[k]] or 0)
tot = sum(gi(r, '# Samples') for r in body)
sw = sum(gi(r, 'L1 Wavefronts Shared') for r in body)
si = sum(gi(r, 'L1 Wavefronts Shared Ideal') for r in body)
tg = sum(gi(r, 'L1 Tag Requests Global') for r in body)
print(f'samples {tot}  shared wavefronts {sw} (ideal {si})  global tag requests {tg}  instructions {sum(gi(r, "Instructions Executed") for r in body)}')
for n, r in enumerate(body):
    wf, t, s = gi(r, 'L1 Wavefronts Shared'), gi(r, 'L1 Tag Requests Global'), gi(r, '# Samples')
    if wf > sw * frac or t > tg * frac or s > tot * frac:
        print(f'{n:5d} {r[ix["Source"]].strip()[:58]:58s} smp {s:6d} inst {gi(r, "Instructions Executed"):8d} '
              f'shwf {wf:9d} ideal {gi(r, "L1 Wavefronts Shared Ideal"):9d} gtag {t:8d} sect {gi(r, "L2 Theoretical Sectors Global"):8d}')
