"""Build libfes_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libfes_b200.so')
SOURCES = ['core.cu', 'plan.cu', 'plan_runs.cu', 'assemble.cu', 'pcg.cu', 'simp.cu', 'fields.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.hpp'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'fes_b200.h'))
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace('.cu', '.o'))
        if force or _stale(o, [s] + headers):
            cmd = ([nvcc] + NVCC_FLAGS + os.environ.get('FES_NVCC_EXTRA', '').split()  # tuning builds only
                   + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o])
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        logs = list(ex.map(run, jobs))
    objs = [os.path.join(objdir, s.replace('.cu', '.o')) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        # the shared CUDA runtime (the process already holds torch's): the library carries no copy of cudart
        run([nvcc, '-shared', '-cudart', 'shared', '-Xlinker', '-rpath=/usr/local/cuda/lib64', '-o', LIB] + objs)
    if verbose:
        print('\n'.join(logs))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
