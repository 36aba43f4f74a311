"""On-disk formats (ref fem/io.py:52-147): the files under tests/golden/io_* were written by the reference's own
`save_mesh` / `save_solution` (oracle/gen_golden.py:io_cases); fes_b200.io must read them and write archives
with the same keys, dtypes and values, so either library loads what the other saved.  CPU only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')


def test_mesh_json_round_trip(tmp_path):
    import fes_b200 as F
    ref = json.load(open(os.path.join(GOLD, 'io_mesh.json')))
    mesh = F.load_mesh(os.path.join(GOLD, 'io_mesh.json'))
    assert np.array_equal(mesh.vertices, np.array(ref['vertices'])) and mesh.vertices.dtype == np.float64
    assert np.array_equal(mesh.elements, np.array(ref['elements'])) and np.array_equal(mesh.boundary, np.array(ref['boundary']))
    out = tmp_path / 'm.json'
    F.save_mesh(mesh, str(out))
    assert json.load(open(out)) == ref                                  # same document, bit-for-bit floats


@pytest.mark.parametrize('name,cls,etype', [('io_wave', 'WaveSolution', None),
                                             ('io_elastic_p2', 'ElasticSolution', 'QuadraticTriangleElement'),
                                             ('io_scalar', 'ScalarFieldSolution', 'LinearTriangleElement'),
                                             ('io_modal', 'ModalSolution', None)])
def test_solution_archives_interoperate(tmp_path, name, cls, etype):
    import fes_b200 as F
    path = os.path.join(GOLD, name + '.npz')
    sol = F.load_solution(path)
    assert type(sol).__name__ == cls
    assert (sol.element_type.__name__ if sol.element_type is not None else None) == etype
    out = tmp_path / 's.npz'
    sol.save(str(out))
    with np.load(path) as a, np.load(out) as b:                         # allow_pickle stays False
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert a[k].dtype.kind == b[k].dtype.kind and a[k].shape == b[k].shape, k
            assert np.array_equal(a[k], b[k]), k
    again = F.Solution.load(str(out))
    assert type(again) is type(sol) and again.n_components == sol.n_components


def test_elastic_solution_measures_and_refusals(tmp_path):
    import fes_b200 as F
    sol = F.load_solution(os.path.join(GOLD, 'io_elastic_p2.npz'))
    s = sol.stress
    dev = s - (np.trace(s, axis1=1, axis2=2) / 3.0)[:, None, None] * np.eye(3)
    np.testing.assert_allclose(sol.von_mises, np.sqrt(1.5 * np.einsum('eij,eij->e', dev, dev)), rtol=1e-14)
    np.testing.assert_allclose(sol.pressure, -np.trace(s, axis1=1, axis2=2) / 3.0, rtol=1e-14)
    assert sol.principal_stress.shape == (len(s), 3) and np.all(np.diff(sol.principal_stress, axis=1) >= 0)
    np.testing.assert_allclose(sol.max_shear, (sol.principal_stress[:, 2] - sol.principal_stress[:, 0]) / 2)
    with pytest.raises(ValueError):
        F.ElasticSolution(sol.mesh, 2, sol.u, sol.strain[:, 0], sol.stress, sol.compliance)
    ragged = F.TransientSolution(sol.mesh, 1, np.arange(2.0), [np.zeros(3), np.zeros(4)])
    with pytest.raises(ValueError, match='ragged'):
        ragged.save(str(tmp_path / 'r.npz'))
    bad = tmp_path / 'bad.npz'
    with np.load(os.path.join(GOLD, 'io_wave.npz')) as d:
        arrays = {k: d[k] for k in d.files}
    arrays['__solution_class__'] = np.array('os')
    np.savez(open(bad, 'wb'), **arrays)
    with pytest.raises(ValueError, match='Unknown solution class'):
        F.load_solution(str(bad))
    arrays['__solution_class__'] = np.array('WaveSolution')
    arrays['__element_type__'] = np.array('_lib')
    np.savez(open(bad, 'wb'), **arrays)
    with pytest.raises(ValueError, match='Unknown element type'):
        F.load_solution(str(bad))


@pytest.mark.gpu
def test_elastic_solution_from_solve_matches_the_reference_archive():
    import fes_b200 as F
    ref = F.load_solution(os.path.join(GOLD, 'io_elastic_p2.npz'))
    V = F.FunctionSpace(ref.mesh, F.QuadraticTriangleElement, n_components=2)
    sol = F.ElasticSolution.from_solve(V, ref.u, F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)))
    assert sol.element_type is F.QuadraticTriangleElement
    np.testing.assert_allclose(sol.strain, ref.strain, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(sol.stress, ref.stress, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(sol.compliance, ref.compliance, rtol=1e-12, atol=1e-16)
    nodal = sol.nodal_von_mises()
    assert nodal.shape == (V.n_nodes,) and np.all(nodal >= 0)


def test_modal_solution_units_and_mode_mesh():
    import fes_b200 as F
    sol = F.load_solution(os.path.join(GOLD, 'io_modal.npz'))
    np.testing.assert_allclose(sol.frequencies, sol.angular_frequencies / (2 * np.pi))
    np.testing.assert_allclose(sol.periods * sol.frequencies, 1.0)
    moved = sol.mode_mesh(1, scale=0.5)
    np.testing.assert_allclose(moved.vertices, sol.mesh.vertices + 0.5 * sol.modes[1].reshape(-1, 2))
    assert np.array_equal(moved.elements, sol.mesh.elements)


@pytest.mark.gpu
def test_scalar_solution_from_solve_matches_the_reference_archive():
    import fes_b200 as F
    ref = F.load_solution(os.path.join(GOLD, 'io_scalar.npz'))
    V = F.FunctionSpace(ref.mesh, n_components=1)
    sol = F.ScalarFieldSolution.from_solve(V, ref.u)
    assert sol.element_type is F.LinearTriangleElement
    np.testing.assert_allclose(sol.flux, ref.flux, rtol=1e-12, atol=1e-14)
    assert sol.nodal_flux().shape == (V.n_nodes, 2)
    Vq = F.FunctionSpace(ref.mesh, F.QuadraticTriangleElement, n_components=1)
    x = Vq.node_coords
    g = Vq.gradient(2.0 * x[:, 0] - 3.0 * x[:, 1] + 1.0)          # a linear field: exact gradient on P2 too
    np.testing.assert_allclose(g, np.tile([2.0, -3.0], (len(g), 1)), rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        V.gradient(np.zeros(V.n_nodes + 1))
