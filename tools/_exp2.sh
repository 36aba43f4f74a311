cd /root/repo
t() { echo -n "$*: "; env "${@:2}" python tools/time_assembly.py $1 2>&1 | grep "elastic" | sed -e 's/.*"ms": \([0-9.]*\).*"hbm_frac": \([0-9.]*\).*/ms=\1 frac=\2/' | tr '\n' ' '; echo; }
t tet A=1
t tet FES_TILE_NODE_MULT=1
t tet FES_TILE_NODE_MULT=8
t tet FES_TILE_INC=420
t c2 FES_TILE_NODE_MULT=16
t c2 FES_TILE_NODE_MULT=8
t c2 FES_TILE_NODE_MULT=16 FES_TILE_INC=300
t c3 FES_TILE_NODE_MULT=8
python -m pytest tests/test_gpu_assembly.py tests/test_gpu_assembly_paths.py -m gpu -x -q 2>&1 | tail -2
