"""The oracle's restatement of the SURVEY 8f rows 2 / 4 functions (LinearForm loads, DiffusionForm,
GeometricStiffnessForm, nodal recovery, derived fields) against outputs of the reference itself
(tests/golden/fields.npz, written by oracle/gen_golden.py).  CPU only."""
import os
import sys

import numpy as np
import pytest
from scipy.sparse.linalg import spsolve

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
import fes_oracle as fo  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'fields.npz'))
KIND = {'tri': fo.P1_TRI, 'p2tri': fo.P2_TRI, 'tet': fo.P1_TET}


def fields_of(name):
    d = 3 if name == 'tet' else 2
    f1 = (lambda X: np.sin(X[..., 0]) + X[..., 1] ** 2) if d == 2 else (lambda X: np.sin(X[..., 0]) + X[..., 1] * X[..., 2])
    fd = (lambda X: np.stack([X[..., 0] * X[..., 1], 1.0 + X[..., 0]], -1)) if d == 2 else \
        (lambda X: np.stack([X[..., 0] * X[..., 1], 1.0 + X[..., 2], X[..., 1] - X[..., 0]], -1))
    kappa = (lambda X: 1.0 + X[..., 0] ** 2 + 0.5 * X[..., 1]) if d == 2 else (lambda X: 1.0 + X[..., 0] ** 2 + 0.5 * X[..., 2])
    return d, f1, fd, kappa


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
def test_loads_match_reference(name):
    d, f1, fd, _ = fields_of(name)
    en, X = G[f'{name}_element_nodes'], G[f'{name}_node_coords']
    geo = fo.geometry(KIND[name], X[en], degree=2)
    np.testing.assert_allclose(geo.points, G[f'{name}_points2'], rtol=0, atol=1e-14)
    np.testing.assert_allclose(geo.wdet, G[f'{name}_wdet2'], rtol=1e-13)
    b1 = fo.assemble_load(en, len(X), geo, f1(geo.points)[..., None])
    np.testing.assert_allclose(b1, G[f'{name}_load1'], rtol=1e-12, atol=1e-14)
    geo_hi = fo.geometry(KIND[name], X[en], degree=int(G[f'{name}_loadd_degree']))
    bd = fo.assemble_load(en, len(X), geo_hi, fd(geo_hi.points))
    np.testing.assert_allclose(bd, G[f'{name}_loadd'], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
def test_weighted_gradient_operators_match_reference(name):
    d, _, _, kappa = fields_of(name)
    en, X = G[f'{name}_element_nodes'], G[f'{name}_node_coords']
    geo2 = fo.geometry(KIND[name], X[en], degree=2)
    A = fo.assemble(en, 1, len(X), fo.ke_diffusion(geo2, kappa(geo2.points)))
    assert np.array_equal(A.indptr, G[f'{name}_diff_indptr']) and np.array_equal(A.indices, G[f'{name}_diff_indices'])
    np.testing.assert_allclose(A.data, G[f'{name}_diff_data'], rtol=1e-12, atol=1e-12 * np.abs(A.data).max())
    geo = fo.geometry(KIND[name], X[en])
    Kg = fo.assemble(en, d, len(X), fo.ke_geometric(geo, G[f'{name}_prestress']))
    assert np.array_equal(Kg.indptr, G[f'{name}_geom_indptr']) and np.array_equal(Kg.indices, G[f'{name}_geom_indices'])
    np.testing.assert_allclose(Kg.data, G[f'{name}_geom_data'], rtol=1e-12, atol=1e-12 * np.abs(Kg.data).max())


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
def test_nodal_recovery_matches_reference(name):
    d = 3 if name == 'tet' else 2
    en, X = G[f'{name}_element_nodes'], G[f'{name}_node_coords']
    geo = fo.geometry(KIND[name], X[en])
    avg = fo.recover_nodal_average(en, geo.volumes, G[f'{name}_vals'], len(X))
    assert np.array_equal(avg, G[f'{name}_avg'])                       # same bincount order: bit-identical
    avg_t = fo.recover_nodal_average(en, geo.volumes, G[f'{name}_tens'], len(X))
    np.testing.assert_allclose(avg_t, G[f'{name}_avg_tens'], rtol=1e-14, atol=1e-15)
    M = fo.assemble(en, 1, len(X), fo.ke_mass(geo, 1))
    l2 = fo.recover_nodal_l2(en, geo, G[f'{name}_vals'], len(X), M.tocsc(), spsolve)
    np.testing.assert_allclose(l2, G[f'{name}_l2'], rtol=1e-10, atol=1e-12)
    l2t = fo.recover_nodal_l2(en, geo, G[f'{name}_tens'], len(X), M.tocsc(), spsolve)
    np.testing.assert_allclose(l2t, G[f'{name}_l2_tens'], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
@pytest.mark.parametrize('tag', ['uni', 'var'])
def test_derived_fields_match_reference(name, tag):
    en, X = G[f'{name}_element_nodes'], G[f'{name}_node_coords']
    geo = fo.geometry(KIND[name], X[en])
    E = 200.0 if tag == 'uni' else G[f'{name}_E_var']
    eps, sig, comp = fo.elastic_fields(geo, G[f'{name}_u'], en, E, 0.3)
    np.testing.assert_allclose(eps, G[f'{name}_strain_{tag}'], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(sig, G[f'{name}_stress_{tag}'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(comp, G[f'{name}_compliance_{tag}'], rtol=1e-12, atol=1e-16)


@pytest.mark.parametrize('name', ['tri', 'p2tri', 'tet'])
@pytest.mark.parametrize('tag', ['uni', 'var'])
def test_stress_field_per_quadrature_point_matches_reference(name, tag):
    """LinearElasticForm.stress_field (ref fem/forms.py:376-398) at the space's rule and at degree 2."""
    en, X = G[f'{name}_element_nodes'], G[f'{name}_node_coords']
    E = 200.0 if tag == 'uni' else G[f'{name}_E_var']
    for key, geo in ((f'{name}_stressfield_{tag}', fo.geometry(KIND[name], X[en])),
                     (f'{name}_stressfield2_{tag}', fo.geometry(KIND[name], X[en], 2))):
        got = fo.stress_field(geo, G[f'{name}_u'], en, E, 0.3)
        assert got.shape == G[key].shape
        np.testing.assert_allclose(got, G[key], rtol=1e-12, atol=1e-12)
