"""Shared pytest setup: the `gpu` marker, import paths, golden-fixture loader."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# The stencil-run assembly path is reserved for meshes above 10^5 nodes by default (below that one kernel launch
# wins); the GPU tests send their small structured meshes through it too, so every structured-mesh test covers the
# run kernel + remainder tiles, while tests/test_gpu_runs.py compares it bit for bit with the tiled kernel
# (FES_NO_RUNS_LAUNCH) and the unstructured / P2 / mass cases keep covering the tiled kernel alone.  Read per plan.
os.environ.setdefault('FES_RUNS_MIN_NODES', '0')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on a B200)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


@pytest.fixture(scope='session')
def golden():
    return load_golden
