cd /root/repo
t() { echo -n "$*: "; env "${@:2}" python tools/time_assembly.py $1 2>&1 | grep "elastic" | sed -e 's/.*"ms": \([0-9.]*\).*"hbm_frac": \([0-9.]*\).*/ms=\1 frac=\2/' | tr '\n' ' '; echo; }
t tet A=1
t tet FES_TILE_THREADS=256
t tet FES_TILE_THREADS=256 FES_TILE_INC=384
t tet FES_TILE_THREADS=256 FES_TILE_INC=320
t c2 A=1
