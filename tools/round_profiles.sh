#!/bin/bash
# Everything the round's profiles/ folder is built from, run on the GPU box (outputs in gpurun_out/):
# GPU tests, the bench lines of both arms, the ncu launch list of the bench command and full ncu
# captures of the dominant kernels.  Each profiled command first runs clean without ncu.
cd "$(dirname "$0")/.."
R=${1:-r02}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench_n1.err
python bench.py --steps 5 --warmup 3 --no-extra > $O/${R}_bench_noextra.json 2>> $O/${R}_bench_n1.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_bench_noextra.csv \
    python bench.py --steps 5 --warmup 3 --no-extra > $O/${R}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_assemble_runs -s 4 -c 1 -f -o $O/${R}_assemble_c2 \
    python bench.py --steps 5 --warmup 3 --no-extra > $O/${R}_ncu_c2.log 2>&1
python tools/profile_tet151.py 151 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_assemble_runs -s 2 -c 1 -f -o $O/${R}_assemble_tet \
    python tools/profile_tet151.py 151 > $O/${R}_ncu_tet.log 2>&1
python tools/profile_hotpath.py p2 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiles -s 2 -c 1 -f -o $O/${R}_assemble_p2 \
    python tools/profile_hotpath.py p2 > $O/${R}_ncu_p2.log 2>&1
python tools/profile_hotpath.py all > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pcg -c 1 -f -o $O/${R}_pcg_c2 \
    python tools/profile_hotpath.py all > $O/${R}_ncu_pcg.log 2>&1
ls -la $O | tail -12
