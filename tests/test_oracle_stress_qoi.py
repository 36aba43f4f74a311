"""Oracle restatement of MeanStress / SoftMaxStress (ref fem/sensitivity.py:108-276) against outputs of the
reference itself (tests/golden/stress_qoi.npz).  CPU only."""
import os
import sys

import numpy as np
import pytest
from scipy.sparse.linalg import splu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import fes_oracle as fo  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'stress_qoi.npz'))
CASES = {'mean': (False, None), 'mean_region_var': (True, None), 'soft': (False, 8.0), 'soft_region_var': (True, 6.0)}


def setup(name):
    verts, elems, bnd = G[f'{name}_vertices'], G[f'{name}_elements'], G[f'{name}_boundary']
    if name == 'p2':
        en, X, _ = fo.p2_nodes(verts, elems, bnd)
        kind = fo.P2_TRI
    else:
        en, X, kind = elems, verts, fo.P1_TRI
    return en, X, fo.geometry(kind, X[en])


@pytest.mark.parametrize('name', ['p1', 'p2'])
@pytest.mark.parametrize('tag', list(CASES))
def test_stress_qoi_matches_reference(name, tag):
    en, X, geo = setup(name)
    var, p = CASES[tag]
    E = G[f'{name}_E_el'] if var else 1.0
    region = G[f'{name}_region'] if var else None
    value, dJ = fo.stress_qoi(geo, G[f'{name}_u'], en, len(X), E, 0.3, region, p)
    np.testing.assert_allclose(value, G[f'{name}_{tag}_value'], rtol=1e-12)
    ref = G[f'{name}_{tag}_dJ_du']
    np.testing.assert_allclose(dJ, ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize('name', ['p1', 'p2'])
def test_von_mises_and_density_gradient_match_reference(name):
    en, X, geo = setup(name)
    u, rho = G[f'{name}_u'], G[f'{name}_rho']
    np.testing.assert_allclose(fo.von_mises_gradient(geo, u, en, 1.0, 0.3)[0], G[f'{name}_vm_u'], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(fo.von_mises_gradient(geo, u, en, G[f'{name}_E_el'], 0.3)[0], G[f'{name}_vm_v'], rtol=1e-11, atol=1e-14)
    # the adjoint gradient of the mean stress with respect to the SIMP densities
    K0 = fo.ke_elastic(geo, 1.0, 0.3)
    A = fo.assemble(en, 2, len(X), (rho ** 3)[:, None, None] * K0)
    fixed = np.flatnonzero(np.isclose(np.repeat(X[:, 0], 2), 0.0))
    free = np.setdiff1d(np.arange(2 * len(X)), fixed)
    _, dJ = fo.stress_qoi(geo, u, en, len(X), 1.0, 0.3)
    lam = fo.constrained_solve(A, free, fixed, np.zeros(len(fixed)), dJ, lambda M, b: splu(M.tocsc()).solve(b))
    grad = -3.0 * rho ** 2 * fo.element_quadratic(K0, u, en, 2, lam)
    ref = G[f'{name}_mean_grad']
    np.testing.assert_allclose(grad, ref, rtol=1e-8, atol=1e-10 * np.abs(ref).max())
