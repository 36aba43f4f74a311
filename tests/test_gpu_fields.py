"""Device paths of the SURVEY 8f rows 2 / 4 functions against outputs of the reference itself
(tests/golden/fields.npz): LinearForm loads, DiffusionForm and GeometricStiffnessForm operators, nodal
recovery (average: the reference's summation order; L2: mass solve by PCG) and LinearElasticForm.derived_fields."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'fields.npz'))


def spaces(name):
    import fes_b200 as F
    mesh = F.Mesh(G[f'{name}_vertices'], G[f'{name}_elements'], G[f'{name}_boundary'])
    etype = F.QuadraticTriangleElement if name == 'p2tri' else None
    d = mesh.spatial_dim
    V1, Vd = F.FunctionSpace(mesh, etype, n_components=1), F.FunctionSpace(mesh, etype, n_components=d)
    assert np.array_equal(V1.element_nodes, G[f'{name}_element_nodes'])
    return F, d, V1, Vd


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
def test_linear_form_loads(name):
    F, d, V1, Vd = spaces(name)
    f1 = (lambda p: [np.sin(p[0]) + p[1] ** 2]) if d == 2 else (lambda p: [np.sin(p[0]) + p[1] * p[2]])
    fd = (lambda p: [p[0] * p[1], 1.0 + p[0]]) if d == 2 else (lambda p: [p[0] * p[1], 1.0 + p[2], p[1] - p[0]])
    pts, wdet = V1.quadrature(2)
    np.testing.assert_allclose(pts.cpu().numpy(), G[f'{name}_points2'], rtol=0, atol=1e-14)
    np.testing.assert_allclose(wdet.cpu().numpy(), G[f'{name}_wdet2'], rtol=1e-13)
    np.testing.assert_allclose(V1.assemble_load(F.LinearForm(f1, 1, 2)), G[f'{name}_load1'], rtol=1e-12, atol=1e-14)
    hi = int(G[f'{name}_loadd_degree'])
    np.testing.assert_allclose(Vd.assemble_load(F.LinearForm(fd, d, hi)), G[f'{name}_loadd'], rtol=1e-12, atol=1e-14)
    # a constant field needs no host evaluation
    const = V1.assemble_load(F.LinearForm(2.5, 1, 2))
    np.testing.assert_allclose(const.sum(), 2.5 * V1.total_volume, rtol=1e-13)
    with pytest.raises(ValueError):
        Vd.assemble_load(F.LinearForm(f1, 1, 2))
    with pytest.raises(NotImplementedError):
        V1.quadrature(9)


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
def test_diffusion_and_geometric_stiffness(name):
    F, d, V1, Vd = spaces(name)
    kappa = (lambda p: [1.0 + p[0] ** 2 + 0.5 * p[1]]) if d == 2 else (lambda p: [1.0 + p[0] ** 2 + 0.5 * p[2]])
    A = V1.assemble(F.DiffusionForm(kappa))
    assert np.array_equal(A.indptr, G[f'{name}_diff_indptr']) and np.array_equal(A.indices, G[f'{name}_diff_indices'])
    np.testing.assert_allclose(A.data, G[f'{name}_diff_data'], rtol=1e-12, atol=1e-12 * np.abs(A.data).max())
    # kappa == 1 is the Laplacian
    L1 = V1.assemble(F.DiffusionForm(1.0)).data
    L0 = V1.assemble(F.LaplacianForm()).data
    np.testing.assert_allclose(L1, L0, rtol=1e-12, atol=1e-12 * np.abs(L0).max())
    Kg = Vd.assemble(F.GeometricStiffnessForm(G[f'{name}_prestress']))
    assert np.array_equal(Kg.indptr, G[f'{name}_geom_indptr']) and np.array_equal(Kg.indices, G[f'{name}_geom_indices'])
    np.testing.assert_allclose(Kg.data, G[f'{name}_geom_data'], rtol=1e-12, atol=1e-12 * np.abs(Kg.data).max())
    with pytest.raises(ValueError):
        Vd.assemble(F.GeometricStiffnessForm(G[f'{name}_prestress'][:-1]))


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
def test_nodal_recovery(name):
    F, d, V1, Vd = spaces(name)
    vals, tens = G[f'{name}_vals'], G[f'{name}_tens']
    # same summation order as the reference (test_node_scatter_is_bincount_bit_for_bit); the element volumes the
    # weights come from are computed on the device and differ from numpy's by an ulp
    np.testing.assert_allclose(V1.recover_nodal(vals), G[f'{name}_avg'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(V1.recover_nodal(tens), G[f'{name}_avg_tens'], rtol=1e-13, atol=1e-15)
    for V, v, key in ((V1, vals, 'l2'), (Vd, tens, 'l2_tens')):
        got = V.recover_nodal(v, method='l2')
        ref = G[f'{name}_{key}']
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-8 * np.abs(ref).max()
    with pytest.raises(ValueError):
        V1.recover_nodal(vals[:-1])
    with pytest.raises(ValueError):
        V1.recover_nodal(vals, method='median')
    # the L2 projection conserves the integral of the field
    q = V1.recover_nodal(vals, method='l2')
    M = V1.mass_matrix
    np.testing.assert_allclose(np.ones(V1.n_nodes) @ (M @ q), (vals * V1.element_volumes).sum(), rtol=1e-9)


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
@pytest.mark.parametrize('tag', ['uni', 'var'])
def test_derived_fields(name, tag):
    F, d, V1, Vd = spaces(name)
    E = 200.0 if tag == 'uni' else G[f'{name}_E_var']
    fields = F.LinearElasticForm(F.LinearElasticMaterial(E, 0.3)).derived_fields(Vd, G[f'{name}_u'])
    np.testing.assert_allclose(fields.strain.cpu().numpy(), G[f'{name}_strain_{tag}'], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(fields.stress.cpu().numpy(), G[f'{name}_stress_{tag}'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(fields.compliance.cpu().numpy(), G[f'{name}_compliance_{tag}'], rtol=1e-12, atol=1e-16)


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
@pytest.mark.parametrize('tag', ['uni', 'var'])
def test_stress_field_per_quadrature_point(name, tag):
    """LinearElasticForm.stress_field (ref fem/forms.py:376-398): (n_el, n_qp, d, d) at the space's rule and at
    the degree-2 rule, against the reference's outputs; the first point of the default rule is the in-plane
    block of derived_fields' stress."""
    F, d, V1, Vd = spaces(name)
    E = 200.0 if tag == 'uni' else G[f'{name}_E_var']
    form = F.LinearElasticForm(F.LinearElasticMaterial(E, 0.3))
    s0 = form.stress_field(Vd, G[f'{name}_u']).cpu().numpy()
    assert s0.shape == G[f'{name}_stressfield_{tag}'].shape
    np.testing.assert_allclose(s0, G[f'{name}_stressfield_{tag}'], rtol=1e-12, atol=1e-12)
    s2 = form.stress_field(Vd, G[f'{name}_u'], degree=2).cpu().numpy()
    np.testing.assert_allclose(s2, G[f'{name}_stressfield2_{tag}'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(s0[:, 0], G[f'{name}_stress_{tag}'][:, :d, :d], rtol=1e-12, atol=1e-12)


def test_node_scatter_is_bincount_bit_for_bit():
    """A million random element values on a shuffled mesh: the device gather equals np.bincount exactly."""
    import torch
    import fes_b200 as F
    mesh = F.create_rect_mesh([[0, 0], [1, 1]], (301, 201))
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(mesh.elements))
    mesh = F.Mesh(mesh.vertices, mesh.elements[perm], mesh.boundary)
    V = F.FunctionSpace(mesh, n_components=2)
    vals = rng.normal(size=(len(mesh.elements), 3, 2))
    got = V.node_scatter(torch.as_tensor(vals), 2).cpu().numpy()
    flat = mesh.elements.ravel()
    ref = np.column_stack([np.bincount(flat, weights=vals.reshape(-1, 2)[:, j], minlength=V.n_nodes) for j in range(2)])
    assert np.array_equal(got, ref)
