"""Import shim for the upstream reference (TEST INFRASTRUCTURE; build container only).

`import fem` fails here because `fem/__init__.py` pulls in matplotlib and
`fem/backends.py` imports pyamg, neither installed.  Registering an empty
`fem` package pointing at the reference tree skips `__init__`, and a stub `pyamg` lets
`fem.backends` import.  The reference location is FES_REFERENCE, else /root/reference (build
container), else `baseline/_ref` -- the git-ignored `pip install --target` of the unmodified
reference that `__graft_entry__.build()` makes and that travels to the GPU box with the snapshot.
Only tests/, smoke() and bench.py's CPU legs may use it.
"""
import os
import sys
import types


_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    cands = [os.environ.get('FES_REFERENCE'), '/root/reference', os.path.join(_REPO, 'baseline', '_ref')]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, 'fem')):
            return c
    return cands[1]


def available():
    return os.path.isdir(os.path.join(reference_root(), 'fem'))


def install():
    if 'fem' in sys.modules:
        return
    if not available():
        raise RuntimeError(f'reference tree not found at {reference_root()}')
    pkg = types.ModuleType('fem')
    pkg.__path__ = [os.path.join(reference_root(), 'fem')]
    sys.modules['fem'] = pkg
    if 'pyamg' not in sys.modules:
        stub = types.ModuleType('pyamg')

        def _missing(*a, **k):
            raise RuntimeError('pyamg is not installed')
        stub.smoothed_aggregation_solver = _missing
        sys.modules['pyamg'] = stub
