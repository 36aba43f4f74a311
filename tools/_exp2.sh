cd /root/repo
python -m pytest tests/test_gpu_assembly.py tests/test_gpu_assembly_paths.py tests/test_gpu_simp.py -m gpu -x -q 2>&1 | tail -2
t() { echo -n "$*: "; env "${@:2}" python tools/time_assembly.py $1 2>&1 | grep "elastic" | sed -e 's/.*"ms": \([0-9.]*\).*"hbm_frac": \([0-9.]*\).*/ms=\1 frac=\2/' | tr '\n' ' '; echo; }
t c2 A=1
t c2 FES_NO_PDL=1
t c2 A=2
t c2 FES_NO_PDL=1
t tet A=1
t tet FES_NO_PDL=1
t c3 A=1
t c3 FES_NO_PDL=1
