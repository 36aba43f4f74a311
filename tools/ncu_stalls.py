"""Top stall sites of the first kernel in an .ncu-rep, with the dominant stall reasons per SASS line."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[start]; ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[start + 1:] if len(r) == len(h)]
stall_cols = [k for k in h if k.startswith('stall_')]
gi = lambda r, k: int(r[ix[k]] or 0)
tot = sum(gi(r, '# Samples') for r in body)
agg = {k: sum(gi(r, k) for r in body) for k in stall_cols}
print('samples', tot, {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
order = sorted(range(len(body)), key=lambda n: -gi(body[n], '# Samples'))[:topn]
for n in sorted(order):
    r = body[n]
    why = sorted(((gi(r, k), k[6:]) for k in stall_cols), reverse=True)[:3]
    print(f'{n:5d} {r[ix["Source"]].strip()[:56]:56s} smp {gi(r, "# Samples"):6d}  ' + ' '.join(f'{k}={v}' for v, k in why if v))
