"""Turn the raw captures of tools/round_profiles.sh (gpurun_out/<round>_*) into the tracked summaries
under profiles/: key ncu metrics per kernel, the per-phase / per-instruction LSU split of the assembly
kernel, the launch list with each kernel's share, the bench JSON lines and traffic.json."""
import csv
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else 'r02'
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
run = lambda *a: subprocess.run([sys.executable, *a], capture_output=True, text=True, cwd=ROOT).stdout

for tag, note in (('assemble_c2', 'k_assemble_runs<P1_TRI,ELASTIC,2,2,128>, config 2 (inside bench.py --no-extra)'),
                  ('assemble_tet', 'k_assemble_runs<P1_TET,ELASTIC,3,3,192>, config 5 on one GPU: 151^3 box (tools/profile_tet151.py)'),
                  ('assemble_p2', 'k_assemble_tiles<P2_TRI,ELASTIC,2,2,128>, config 3: 708^2 vertices, 999 698 six-node triangles (tools/profile_hotpath.py p2)'),
                  ('pcg_c2', 'k_pcg<2>, config 2, 40 iterations (tools/profile_hotpath.py all)')):
    rep = os.path.join(G, f'{R}_{tag}.ncu-rep')
    if not os.path.exists(rep):
        continue
    text = f'ncu --set full --clock-control none: {note}\n' + run('tools/ncu_summary.py', rep)
    if tag.startswith('assemble'):
        text += '\n-- per phase (split at BAR.SYNC); run kernel: prologue | phase 1 | staging + phase 2 | store; tiled kernel: prologue | prologue | phase 1 | coordinates + phase 2 | store | ...\n'
        text += run('tools/ncu_phases.py', rep)
        text += '\n-- LSU cost per SASS instruction (>= 1.5 % of the kernel)\n' + run('tools/ncu_src.py', rep, '0.015')
        text += '\n-- top stall sites\n' + run('tools/ncu_stalls.py', rep, '12')
    open(os.path.join(P, f'{R}_{tag}.txt'), 'w').write(text)

rep = os.path.join(G, f'{R}_assemble_c2.ncu-rep')
if os.path.exists(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    idx = {h: i for i, h in enumerate(rows[0])}
    val = lambda k: float(rows[2][idx[k]]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[rows[1][idx[k]]]
    traffic = int(val('dram__bytes_read.sum') + val('dram__bytes_write.sum'))
    old = json.load(open(os.path.join(P, 'traffic.json'))) if os.path.exists(os.path.join(P, 'traffic.json')) else {}
    old.update({'k_assemble_runs_c2_bytes_per_launch': traffic,
                'source_runs': f'ncu --set full, profiles/{R}_assemble_c2.txt (dram__bytes_read.sum + dram__bytes_write.sum of one '
                               'k_assemble_runs launch = one assembly of config 2)'})
    json.dump(old, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)

src = os.path.join(G, f'{R}_launches_bench_noextra.csv')
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, f'{R}_launches_bench_noextra.csv'))
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    agg, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        name = r[4].split('(')[0][-60:]
        agg[name] += float(r[14]) / 1e3
        cnt[name] += 1
    tot = sum(agg.values())
    lines = ['ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 5 --warmup 3 --no-extra',
             '(cold-cache, serialised launches: compare shares, not absolutes).  The timed region of the bench is the',
             '5 + 3 warm-up steps = launches of k_assemble_runs<1,2,2,2,128> + k_assemble_tiles<1,2,2,2,128> (the remainder tiles);',
             'everything else is the one-off plan build and the',
             'end-to-end leg (which launches the same kernel).', '']
    for name, us in sorted(agg.items(), key=lambda kv: -kv[1]):
        lines.append(f'{us:12.1f} us {100 * us / tot:5.1f}%  x{cnt[name]:<4d} {name}')
    open(os.path.join(P, f'{R}_launches_bench_noextra_summary.txt'), 'w').write('\n'.join(lines) + '\n')

for name in ('bench_n1.json', 'bench_reference.json', 'bench_noextra.json', 'pytest_gpu.log'):
    if os.path.exists(os.path.join(G, f'{R}_{name}')):
        shutil.copy(os.path.join(G, f'{R}_{name}'), os.path.join(P, f'{R}_{name}'))
print(sorted(os.listdir(P)))
