cd /root/repo
python -m pytest tests/test_gpu_assembly.py tests/test_gpu_assembly_paths.py -m gpu -x -q 2>&1 | tail -2
python tools/time_assembly.py c1,c2,c3,c4,tet 2>&1 | sed -e 's/.*"case": "\([^"]*\)".*"ms": \([0-9.]*\).*"hbm_frac": \([0-9.]*\).*/\1: ms=\2 frac=\3/'
