"""GPU parity of the algebra seam: Jacobi-PCG behind Backend.prepare/solve and the
Dirichlet elimination of DiscreteSystem, against the reference's direct solutions (1e-8)."""
import numpy as np
import pytest
from scipy.sparse import csr_array

import fes_oracle as fo
from conftest import load_golden

pytestmark = pytest.mark.gpu


def fes():
    import fes_b200
    return fes_b200


def _spd(n, seed=0):
    rng = np.random.default_rng(seed)
    M = rng.normal(size=(n, n))
    return M @ M.T + n * np.eye(n)


def test_backend_protocol_on_host_operators():
    """ref tests/test_backends.py:120-146 and tests/test_system.py:14-66."""
    F = fes()
    A = _spd(15, seed=4)
    solver = F.JacobiPCGBackend(rtol=1e-12).prepare(csr_array(A))
    for seed in range(3):
        b = np.random.default_rng(seed).normal(size=15)
        x = solver.solve(b)
        assert isinstance(x, np.ndarray) and x.dtype == np.float64
        np.testing.assert_allclose(x, np.linalg.solve(A, b), atol=1e-9)
    dense = F.JacobiPCGBackend().prepare(A)                 # a dense ndarray is a legal operator
    np.testing.assert_allclose(dense.solve(np.ones(15)), np.linalg.solve(A, np.ones(15)), atol=1e-8)
    np.testing.assert_array_equal(solver.solve(np.zeros(15)), np.zeros(15))


def test_discrete_system_matches_dense_elimination():
    F = fes()
    A = _spd(12, seed=3)
    b = np.linspace(-1, 1, 12)
    free, fixed, vals = np.arange(2, 12), np.array([0, 1]), np.array([0.3, -0.4])
    x = F.DiscreteSystem(A, (free, fixed, vals), F.JacobiPCGBackend(rtol=1e-13)).solve(b)
    expected = np.zeros(12)
    expected[fixed] = vals
    expected[free] = np.linalg.solve(A[np.ix_(free, free)], b[free] - A[np.ix_(free, fixed)] @ vals)
    np.testing.assert_allclose(x, expected, atol=1e-10)
    np.testing.assert_array_equal(x[fixed], vals)
    sys_ = F.DiscreteSystem(A, (free, fixed, vals), F.JacobiPCGBackend(rtol=1e-13))
    xh = sys_.solve_homogeneous(b)
    assert np.all(xh[fixed] == 0)
    np.testing.assert_allclose(xh[free], np.linalg.solve(A[np.ix_(free, free)], b[free]), atol=1e-10)
    # no fixed dofs: a plain solve; interleaved free/fixed sets
    x = F.DiscreteSystem(_spd(5, 3), (np.arange(5), np.array([], dtype=int), np.array([]))).solve(np.ones(5))
    np.testing.assert_allclose(x, np.linalg.solve(_spd(5, 3), np.ones(5)), atol=1e-8)
    A = _spd(8, seed=1)
    free, fixed = np.array([0, 1, 4, 5, 6]), np.array([2, 3, 7])
    x = F.DiscreteSystem(A, (free, fixed, np.array([1.0, 2.0, 3.0])), F.JacobiPCGBackend(rtol=1e-13)).solve(
        np.linspace(-1, 1, 8))
    np.testing.assert_allclose((A @ x - np.linspace(-1, 1, 8))[free], 0, atol=1e-9)


def test_non_convergence_raises():
    """ref tests/test_backends.py:195-207."""
    F = fes()
    A = _spd(40, seed=5)
    system = F.DiscreteSystem(A, (np.arange(40), np.array([], dtype=int), np.array([])),
                              F.JacobiPCGBackend(rtol=1e-14, maxiter=1))
    with pytest.raises(RuntimeError, match='CG failed'):
        system.solve(np.ones(40))
    indefinite = np.diag(np.r_[np.ones(5), -np.ones(5)]) + 0.01
    with pytest.raises(RuntimeError, match='CG failed'):
        F.JacobiPCGBackend().prepare(indefinite).solve(np.arange(10.0) - 4.5)


def test_iteration_count_matches_scipy_semantics():
    F = fes()
    g = load_golden('solve')
    v, e = g['c2d_vertices'], g['c2d_elements']
    A = fo.assemble(e, 2, len(v), fo.ke_elastic(fo.geometry(fo.P1_TRI, v[e]), 200.0, 0.3))
    free = g['c2d_free']
    Aff, b = A[np.ix_(free, free)], g['c2d_load'][free]
    x_o, its_o = fo.jacobi_pcg(Aff, b, rtol=1e-10)
    solver = F.JacobiPCGBackend(rtol=1e-10).prepare(Aff)
    x = solver.solve(b)
    assert abs(solver.last_iterations - its_o) <= 2
    assert solver.last_relres < 1e-10
    np.testing.assert_allclose(x, x_o, atol=1e-9 * np.abs(x_o).max())


def _problem(F, pre, g):
    mesh = F.Mesh(g[pre + '_vertices'], g[pre + '_elements'], g[pre + '_boundary'])
    from fes_b200.regions import on_plane
    bc = F.BoundaryConditions()
    if pre == 'c2d':
        bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0, 0])
        bc.add(F.BCType.NEUMANN, on_plane(0, 2.0), [50, -10])
        return F.linear_elastic(mesh, F.LinearElasticMaterial(200.0, 0.3), bc)
    if pre == 'c3d':
        bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0, 0, 0])
        bc.add(F.BCType.NEUMANN, on_plane(0, 1.0), [0, -5, 0])
        return F.linear_elastic(mesh, F.LinearElasticMaterial(200.0, 0.3), bc)
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), 1.5)
    bc.add(F.BCType.DIRICHLET, on_plane(0, 1.0), -0.5)
    bc.add_robin(on_plane(1, 1.0), 2.0, 0.75)
    return F.poisson(mesh, lambda p: [1.0 + p[0] * p[1]], bc)


@pytest.mark.parametrize('pre', ['c2d', 'c3d', 'poi'])
def test_full_solves_match_the_reference(pre):
    """Problem -> tangent (device CSR) -> DiscreteSystem -> PCG == the reference's direct solve."""
    F = fes()
    g = load_golden('solve')
    prob = _problem(F, pre, g)
    free, fixed, vals = prob.constraints
    np.testing.assert_array_equal(free, g[pre + '_free'])
    np.testing.assert_allclose(prob.load, g[pre + '_load'], rtol=1e-12, atol=1e-13 * np.abs(g[pre + '_load']).max())
    u = F.DiscreteSystem(prob.tangent(None), prob.constraints, F.JacobiPCGBackend(rtol=1e-12)).solve(prob.load)
    np.testing.assert_allclose(u, g[pre + '_u'], atol=1e-8 * np.abs(g[pre + '_u']).max())
    if pre == 'poi':   # Robin term: kept explicit zeros, so compare densely
        np.testing.assert_allclose(prob.tangent(None).toarray(), g['poi_A_dense'], atol=1e-12)
    # the reference's own DirectBackend-style object plugs into our DiscreteSystem too
    from scipy.sparse.linalg import splu
    from scipy.sparse import csc_array

    class Direct:
        def prepare(self, A):
            return splu(csc_array(A))
    u2 = F.DiscreteSystem(prob.tangent(None), prob.constraints, Direct()).solve(prob.load)
    np.testing.assert_allclose(u2, g[pre + '_u'], atol=1e-10 * np.abs(g[pre + '_u']).max())


def test_warm_start_and_device_vectors():
    import torch
    F = fes()
    g = load_golden('solve')
    prob = _problem(F, 'c3d', g)
    system = F.DiscreteSystem(prob.tangent(None), prob.constraints, F.JacobiPCGBackend(rtol=1e-12))
    u = system.solve(prob.load_d)
    assert isinstance(u, torch.Tensor)
    cold = system.solver.last_iterations
    u2 = system.solve(prob.load_d, x0=u.clone())
    assert system.solver.last_iterations <= 2 < cold
    np.testing.assert_allclose(u2.cpu().numpy(), g['c3d_u'], atol=1e-8 * np.abs(g['c3d_u']).max())


def test_cantilever_matches_a_direct_solve_at_127k_dofs():
    """The 4:1 cantilever of BASELINE.md 2.3 (504x126 nodes, 127 008 dofs): Jacobi-PCG at the
    reference's rtol=1e-10 against a host sparse LU of the same operator, 1e-8 relative."""
    from scipy.sparse import csc_array
    from scipy.sparse.linalg import splu
    F = fes()
    from fes_b200.regions import on_plane
    mesh = F.create_rect_mesh([[0, 0], [4, 1]], (504, 126))
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0, 0])
    bc.add(F.BCType.NEUMANN, on_plane(0, 4.0), [0, -1])
    prob = F.linear_elastic(mesh, F.LinearElasticMaterial(200.0, 0.3), bc)
    system = F.DiscreteSystem(prob.tangent(None), prob.constraints, F.JacobiPCGBackend(rtol=1e-10))
    u = system.solve(prob.load)
    free, fixed, vals = prob.constraints
    A = prob.tangent(None).to_scipy()
    direct = np.zeros_like(u)
    direct[free] = splu(csc_array(A[np.ix_(free, free)])).solve(prob.load[free])
    assert np.abs(u - direct).max() <= 1e-8 * np.abs(direct).max()
    assert 1000 < system.solver.last_iterations < 10000


def test_large_cantilever_residual():
    """A 1.0 M-triangle cantilever: the recurrence residual meets rtol (scipy's stopping rule) and
    the TRUE residual, measured with an independent scipy SpMV on the host, stays within the
    drift CG's recurrence accumulates over thousands of iterations at this conditioning."""
    F = fes()
    from fes_b200.regions import on_plane
    mesh = F.create_rect_mesh([[0, 0], [4, 1]], (1415, 354))
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0, 0])
    bc.add(F.BCType.NEUMANN, on_plane(0, 4.0), [0, -1])
    prob = F.linear_elastic(mesh, F.LinearElasticMaterial(200.0, 0.3), bc)
    system = F.DiscreteSystem(prob.tangent(None), prob.constraints, F.JacobiPCGBackend(rtol=1e-10))
    u = system.solve(prob.load)
    assert system.solver.last_relres < 1e-10
    free = prob.constraints[0]
    A = prob.tangent(None).to_scipy()
    r = (A @ u - prob.load)[free]
    assert np.linalg.norm(r) <= 1e-6 * np.linalg.norm(prob.load[free])
    assert system.solver.last_iterations > 100
    tip = u.reshape(-1, 2)[np.argmax(mesh.vertices[:, 0] + mesh.vertices[:, 1])]
    # Euler-Bernoulli tip deflection with the plane-strain modulus, within a few percent
    Eps, I, L, P = 200.0 / (1 - 0.09), 1.0 / 12, 4.0, 1.0
    assert tip[1] == pytest.approx(-P * L ** 3 / (3 * Eps * I), rel=0.05)
