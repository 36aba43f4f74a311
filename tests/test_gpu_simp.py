"""GPU parity of the SIMP inner loop against the reference's TopologyOptimizer trajectory
(teacher-forced per SURVEY.md section 7), the cone filter and the OC update."""
import numpy as np
import pytest

import fes_oracle as fo
from conftest import load_golden

pytestmark = pytest.mark.gpu


def fes():
    import fes_b200
    return fes_b200


def test_cone_filter_matches_reference():
    """ref tests/test_topology_optimization.py:129-174."""
    F = fes()
    g = load_golden('filter_oc')
    for pre in ('sq', 'box'):
        mesh = F.Mesh(g[pre + '_vertices'], g[pre + '_elements'])
        S = F.calculate_smoothing_matrix(mesh, float(g[pre + '_r']))
        np.testing.assert_array_equal(S.indptr, g[pre + '_indptr'])
        np.testing.assert_array_equal(S.indices, g[pre + '_indices'])
        np.testing.assert_allclose(S.data, g[pre + '_data'], rtol=1e-13, atol=1e-16)
        uniform = np.full(S.shape[0], 2.5)
        np.testing.assert_allclose(S @ uniform, uniform)
    mesh = F.create_rect_mesh([[0, 0], [1, 1]], (12, 12))
    S0 = F.calculate_smoothing_matrix(mesh, 0.0)            # r = 0: weightless rows stay zero
    assert S0.nnz == len(mesh.elements) and np.all(S0.data == 0.0)
    centers = mesh.vertices[mesh.elements].mean(axis=1)
    So = fo.cone_filter(centers, 0.31)
    S = F.calculate_smoothing_matrix(mesh, 0.31)
    np.testing.assert_array_equal(S.indices, So.indices)
    np.testing.assert_allclose(S.data, So.data, rtol=1e-13, atol=1e-16)


def test_oc_update_matches_reference():
    """ref tests/test_design.py:74-96 plus the golden outputs of fem/design.py:39-78."""
    F = fes()
    g = load_golden('filter_oc')
    out = F.optimality_criteria_update(g['oc_rho'], g['oc_sens'], g['oc_vol'], 0.4, move=0.1)
    np.testing.assert_allclose(out, g['oc_out_04'], rtol=1e-12, atol=1e-14)
    out = F.optimality_criteria_update(g['oc_rho'], g['oc_sens'], g['oc_vol'], 0.6, move=0.2)
    np.testing.assert_allclose(out, g['oc_out_06'], rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError, match='nonnegative'):
        F.optimality_criteria_update(np.full(10, 0.5), np.linspace(-1, 1, 10), np.ones(10), 0.4)
    upd = F.optimality_criteria_update(np.full(20, 0.5), np.linspace(1.0, 2.0, 20), np.ones(20), 0.4)
    assert abs(upd.mean() - 0.4) < 1e-3 and np.all(upd >= 1e-6) and np.all(upd <= 1.0)
    big = np.random.default_rng(1)
    rho, s, v = big.uniform(0.05, 1, 300001), big.uniform(0, 3, 300001), big.uniform(0.5, 1.5, 300001)
    np.testing.assert_allclose(F.optimality_criteria_update(rho, s, v, 0.5), fo.oc_update(rho, s, v, 0.5),
                               rtol=1e-10, atol=1e-12)


def _mbb(F, g, **kw):
    from fes_b200.regions import in_box, intersect, on_plane
    w, h = 3.0, 1.0
    mesh = F.Mesh(g['vertices'], g['elements'], g['boundary'])
    bc = F.BoundaryConditions()
    bottom, top = on_plane(1, 0.0), on_plane(1, h)
    bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
    bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
    bc.add(F.BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
    return F.TopologyOptimizer(mesh, F.LinearElastic(200.0, 0.4), bc, iters=5, volume_frac=0.5,
                               smoothing_radius=float(g['radius']), penalty=3.0,
                               backend=F.JacobiPCGBackend(rtol=1e-12), **kw)


def test_simp_teacher_forced_steps_match_reference():
    F = fes()
    g = load_golden('simp')
    topo = _mbb(F, g)
    np.testing.assert_allclose(topo._problem.load, g['load'], rtol=1e-12, atol=1e-15)
    S = topo.smoothing_matrix
    np.testing.assert_array_equal(S.indices, g['filter_indices'])
    np.testing.assert_allclose(S.data, g['filter_data'], rtol=1e-13, atol=1e-16)
    rhos = list(g['rho']) + [g['rho_final']]
    for k in range(5):
        topo.set_rho(rhos[k])
        sol = topo._solve()
        comp = g['compliance'][k]
        assert sol.total_compliance == pytest.approx(comp.sum(), rel=1e-8)
        np.testing.assert_allclose(sol.compliance, comp, atol=1e-8 * comp.max())
        np.testing.assert_allclose(sol.von_mises, g['von_mises'][k], atol=1e-8 * g['von_mises'][k].max())
        sens = topo._compliance_sensitivity(sol, 1.0)
        np.testing.assert_allclose(sens.cpu().numpy(), g['sens'][k], atol=1e-8 * g['sens'][k].max())
        rho_next = topo.oc_density(topo._filter.apply(sens), 0.5).cpu().numpy()
        np.testing.assert_allclose(rho_next, rhos[k + 1], rtol=1e-9, atol=1e-10)
        # u agrees where material is present (void-region displacements are ill-determined)
        solid = np.unique(g['elements'][rhos[k] > 0.1])
        du = (sol.u - g['u'][k]).reshape(-1, 2)[solid]
        assert np.abs(du).max() <= 1e-7 * np.abs(g['u'][k]).max()


@pytest.mark.parametrize('warm', [False, True])
def test_simp_free_running_trajectory(warm):
    F = fes()
    g = load_golden('simp')
    topo = _mbb(F, g, warm_start=warm)
    seen = []
    hist = topo.solve(on_iteration=lambda i, s: seen.append((i, s.total_compliance)))
    assert [i for i, _ in seen] == list(range(5))
    for k in range(5):
        np.testing.assert_allclose(hist.rho[k], g['rho'][k], rtol=1e-8, atol=1e-9)
        assert hist.compliance[k].sum() == pytest.approx(g['compliance'][k].sum(), rel=1e-8)
    np.testing.assert_allclose(topo.rho, g['rho_final'], rtol=1e-8, atol=1e-9)
    assert abs(topo._volume_fraction(topo.rho) - 0.5) < 1e-6
    assert len(hist.u) == 5 and hist.u[0].shape == (2 * len(g['vertices']),)


def test_simp_dilution_and_scaled_modulus():
    """ref tests/test_topology_optimization.py:43-86."""
    F = fes()
    g = load_golden('simp')
    topo = _mbb(F, g)
    topo.penalty = 2.0
    topo.set_rho(np.full(len(g['elements']), 0.5))
    assert np.allclose(topo.scaled_modulus, 0.5 ** 2.0 * 200.0)
    V = topo.space
    rho = np.linspace(0.2, 1.0, len(g['elements']))
    solid = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)).element_matrices(V).cpu().numpy()
    diluted = F.LinearElasticForm(F.LinearElasticMaterial(rho ** 3 * 200.0, 0.3)).element_matrices(V).cpu().numpy()
    np.testing.assert_allclose(rho[:, None, None] ** 3 * solid, diluted, rtol=1e-12, atol=1e-12 * np.abs(solid).max())


def _mbb100(F, g, **kw):
    from fes_b200.regions import in_box, intersect, on_plane
    w, h = (float(x) for x in g['width_height'])
    mesh = F.create_rect_mesh([[0, 0], [w, h]], tuple(int(r) for r in g['resolution']))
    bc = F.BoundaryConditions()
    bottom, top = on_plane(1, 0.0), on_plane(1, h)
    bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
    bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
    bc.add(F.BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
    return F.TopologyOptimizer(mesh, F.LinearElastic(float(g['E0']), float(g['nu'])), bc, iters=100,
                               volume_frac=float(g['volume_frac']), smoothing_radius=float(g['radius']),
                               penalty=float(g['penalty']), **kw)


def test_simp_100_iterations_against_the_reference_run():
    """Config 4 as BASELINE.json states it (100 design iterations, ref fem/topology.py:291-317) on the reduced MBB
    mesh of tests/golden/simp100.npz, against the reference's own run with its direct solver:
      (i)  FREE-RUNNING: total compliance of every one of the 100 iterations within 1e-8 relative, the volume
           fraction held, no solve at the iteration cap;
      (ii) TEACHER-FORCED at iterations 0 ... 99 (void elements at rho = 1e-6 from ~10 on): rho_{k+1} within 2e-8,
           per-element compliance within 1e-8 of its maximum (checks i-iv).
    The optimality-criteria update bisects its multiplier m to `hi - lo <= 1e-8 hi` (ref fem/design.py:39-78), so
    rho_{k+1} ~ m^(-1/2) is DEFINED only to ~5e-9 relative on the elements no bound or move limit clips: when the
    last bisection steps compare a volume sum at near equality, the device's fixed-tree sum and numpy's pairwise sum
    may take different branches.  Nine of the ten checkpoints agree to 1e-10 .. 1e-12; at iteration 75 the two
    unclipped elements (of 2400) differ by 7.5e-9, independently of the CG tolerance (1e-12 .. 1e-14 measured)."""
    F = fes()
    g = load_golden('simp100')
    topo = _mbb100(F, g, backend=F.JacobiPCGBackend(rtol=1e-12), history=None)
    free, fixed, vals = topo._problem.constraints                      # the same problem the reference resolved
    np.testing.assert_array_equal(free, g['free'])
    np.testing.assert_array_equal(fixed, g['fixed'])
    np.testing.assert_allclose(np.asarray(topo._problem.load), g['load'], rtol=1e-12, atol=1e-14)
    comps = [topo.step().total_compliance for _ in range(100)]
    np.testing.assert_allclose(comps, g['total_compliance'], rtol=1e-8)
    assert abs(topo._volume_fraction(topo.rho) - 0.5) < 1e-6
    assert max(topo.cg_iterations) < 10 * int(topo._mask.sum().item())
    forced = _mbb100(F, g, backend=F.JacobiPCGBackend(rtol=1e-12), warm_start=False, history=None)
    for j, k in enumerate(g['kept_iterations']):
        forced.set_rho(g['rho_kept'][j])
        sol = forced.step()
        assert sol.total_compliance == pytest.approx(g['total_compliance'][k], rel=1e-8)
        np.testing.assert_allclose(sol.compliance, g['compliance_kept'][j], atol=1e-8 * g['compliance_kept'][j].max())
        np.testing.assert_allclose(forced.rho, g['rho_next_kept'][j], rtol=2e-8, atol=2e-8)
