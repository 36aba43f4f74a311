"""Distributed Jacobi-PCG over NVLink peer memory (needs >= 2 GPUs; skipped otherwise).
Launches tools/dist_check.py under torchrun: partitioned assembly + in-kernel halo exchange and
reductions must reproduce the one-GPU solution (1e-9) with the same iteration count."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_pcg_matches_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    env = dict(os.environ, FES_DIST_N='41')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29541',
                        os.path.join(ROOT, 'tools', 'dist_check.py')],
                       capture_output=True, text=True, timeout=600, env=env)
    assert 'DIST_CHECK PASS' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
