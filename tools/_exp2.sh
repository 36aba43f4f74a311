cd /root/repo
python -m pytest tests/test_gpu_assembly.py tests/test_gpu_assembly_paths.py -m gpu -x -q 2>&1 | tail -2
t() { echo -n "$*: "; env "${@:2}" python tools/time_assembly.py $1 2>&1 | grep "elastic" | sed -e 's/.*"ms": \([0-9.]*\).*"hbm_frac": \([0-9.]*\).*/ms=\1 frac=\2/' | tr '\n' ' '; echo; }
t tet A=1
t tet FES_SLOT_MODE=0
t c3 FES_SLOT_MODE=0
t c3 FES_SLOT_MODE=1
t c2 A=1
