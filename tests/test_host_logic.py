"""CPU-only checks of the host layer: mesh generators, P2 numbering, boundary-condition
resolution, the C-ABI surface (library loads and exports every declared symbol), and that the
product path refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

import fes_oracle as fo
from conftest import ROOT, load_golden

import fes_b200
from fes_b200 import _lib
from fes_b200.boundary import BCType, BoundaryConditions
from fes_b200.mesh import Mesh, boundary_facets, create_box_mesh, create_rect_mesh
from fes_b200.regions import everywhere, in_box, intersect, on_plane
from fes_b200.space import dof_indices, p2_connectivity


def test_structured_meshes_match_the_reference_fixtures():
    g = load_golden('assembly_tri')
    m = create_rect_mesh([[0, 0], [2, 1]], (6, 5))
    np.testing.assert_array_equal(m.vertices, g['vertices'])
    np.testing.assert_array_equal(m.elements, g['elements'])
    np.testing.assert_array_equal(m.boundary, g['boundary'])
    g = load_golden('assembly_tet')
    m = create_box_mesh([[0, 0, 0], [1, 2, 1.5]], (3, 4, 3))
    np.testing.assert_array_equal(m.vertices, g['vertices'])
    np.testing.assert_array_equal(m.elements, g['elements'])
    np.testing.assert_array_equal(m.boundary, g['boundary'])


@pytest.mark.parametrize('res', [(2, 2), (3, 7), (64, 65), (70, 61), (129, 5)])
def test_rect_mesh_fast_boundary_matches_the_generic_one(res):
    m = create_rect_mesh([[0, 0], [3, 1]], res)
    v, e = fo.rect_mesh([[0, 0], [3, 1]], res)
    np.testing.assert_array_equal(m.vertices, v)
    np.testing.assert_array_equal(m.elements, e)
    np.testing.assert_array_equal(m.boundary, fo.boundary_facets(e))
    np.testing.assert_array_equal(m.boundary, boundary_facets(e))


def test_p2_connectivity_matches_reference():
    g = load_golden('assembly_p2tri')
    en, X, bn = p2_connectivity(Mesh(g['vertices'], g['elements'], g['boundary']))
    np.testing.assert_array_equal(en, g['element_nodes'])
    np.testing.assert_array_equal(bn, g['boundary_nodes'])
    np.testing.assert_array_equal(X, g['node_coords'])


def test_dof_indices_interleave():
    np.testing.assert_array_equal(dof_indices([3, 0, 5], 2), [6, 7, 0, 1, 10, 11])
    assert dof_indices(np.arange(12).reshape(4, 3), 3).shape == (4, 9)


def test_boundary_resolution_matches_reference():
    g = load_golden('solve')
    mesh = Mesh(g['c2d_vertices'], g['c2d_elements'], g['c2d_boundary'])
    bc = BoundaryConditions()
    bc.add(BCType.DIRICHLET, on_plane(0, 0.0), [0, 0])
    bc.add(BCType.NEUMANN, on_plane(0, 2.0), [50, -10])
    r = bc.resolve(mesh, 2)
    np.testing.assert_array_equal(r.free_idxs, g['c2d_free'])
    np.testing.assert_array_equal(r.fixed_idxs, g['c2d_fixed'])
    np.testing.assert_array_equal(r.fixed_values, g['c2d_vals'])
    assert len(r.neumann) == 1 and r.neumann[0].facet_mask.sum() == 4

    mesh = Mesh(g['poi_vertices'], g['poi_elements'], g['poi_boundary'])
    bc = BoundaryConditions()
    bc.add(BCType.DIRICHLET, on_plane(0, 0.0), 1.5)
    bc.add(BCType.DIRICHLET, on_plane(0, 1.0), -0.5)
    bc.add_robin(on_plane(1, 1.0), 2.0, 0.75)
    r = bc.resolve(mesh, 1)
    np.testing.assert_array_equal(r.free_idxs, g['poi_free'])
    np.testing.assert_array_equal(r.fixed_idxs, g['poi_fixed'])
    np.testing.assert_array_equal(r.fixed_values, g['poi_vals'])

    s = load_golden('simp')
    mesh = Mesh(s['vertices'], s['elements'], s['boundary'])
    w, h = 3.0, 1.0
    bc = BoundaryConditions()
    bottom, top = on_plane(1, 0.0), on_plane(1, h)
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
    bc.add(BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
    r = bc.resolve(mesh, 2)
    np.testing.assert_array_equal(r.free_idxs, s['free'])
    np.testing.assert_array_equal(r.fixed_idxs, s['fixed'])
    np.testing.assert_array_equal(r.fixed_values, s['vals'])


def test_region_tolerance_matches_reference():
    """The reference's regions absorb round-off with atol = 1e-9, inclusive (ref fem/regions.py:26-58):
    0.4 * 3.0 = 1.2000000000000002 must still select the mesh vertex at x = 1.2.  The 61x21 MBB problem of
    tests/golden/simp100.npz (constraints and load written by the reference) has such edges."""
    g = load_golden('simp100')
    w, h = (float(x) for x in g['width_height'])
    mesh = create_rect_mesh([[0, 0], [w, h]], tuple(int(r) for r in g['resolution']))
    bc = BoundaryConditions()
    bottom, top = on_plane(1, 0.0), on_plane(1, h)
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
    bc.add(BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
    r = bc.resolve(mesh, 2)
    np.testing.assert_array_equal(r.free_idxs, g['free'])
    np.testing.assert_array_equal(r.fixed_idxs, g['fixed'])
    loaded_nodes = np.flatnonzero(g['load'].reshape(-1, 2)[:, 1])
    assert r.neumann[0].facet_mask.sum() == len(loaded_nodes) - 1     # a contiguous strip of facets
    assert in_box([1.2000000000000002], [None])(np.array([[1.2]]))[0] and on_plane(0, 1.0)(np.array([[1.0 + 5e-10]]))[0]


def test_boundary_condition_guardrails():
    mesh = create_rect_mesh([[0, 0], [1, 1]], (4, 4))
    bc = BoundaryConditions()
    bc.add('dirichlet', on_plane(0, 0.0), [0.0, 0.0])
    bc.add('dirichlet', on_plane(1, 0.0), [1.0, None])     # corner (0,0): x pinned to 0 and to 1
    with pytest.raises(ValueError, match='conflicting Dirichlet'):
        bc.resolve(mesh, 2)
    bc = BoundaryConditions()
    bc.add('dirichlet', everywhere(), [0.0, None])
    bc.add('neumann', on_plane(0, 1.0), [1.0, 0.0])
    with pytest.raises(ValueError, match='Dirichlet and a Neumann'):
        bc.resolve(mesh, 2)
    with pytest.raises(ValueError):
        BoundaryConditions().add('robin', everywhere(), 1.0)
    with pytest.raises(TypeError):
        BoundaryConditions().add('dirichlet', 'left', 1.0)
    # a callable source evaluated array-wise equals the per-point loop
    from fes_b200.regions import evaluate_field
    f = lambda p: [1.0 + p[0] * p[1], np.sin(p[0])]
    a = evaluate_field(f, mesh.vertices, 2)
    b = np.array([f(p) for p in mesh.vertices])
    np.testing.assert_array_equal(a, b)
    g = lambda p: [1.0 if p[0] > 0.5 else 0.0]              # not vectorisable -> loop fallback
    np.testing.assert_array_equal(evaluate_field(g, mesh.vertices, 1)[:, 0], (mesh.vertices[:, 0] > 0.5) * 1.0)


def test_mesh_validation():
    with pytest.raises(NotImplementedError):
        Mesh(np.zeros((6, 2)), np.arange(6).reshape(1, 6))
    with pytest.raises(ValueError):
        Mesh(np.zeros((3, 2)), [[0, 1, 7]])


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'fes_b200.h')).read()
    declared = set(re.findall(r'\b(fes_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert _lib.lib().fes_version() >= 100
    assert _lib.lib().fes_last_error() is not None


def test_product_path_never_touches_the_oracle():
    pkg = os.path.join(ROOT, 'finite-element-solver_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'fes_oracle' not in text and 'ref_shim' not in text, f


def test_kernels_are_hand_written_no_library_device_code():
    """The CUDA sources include no device library (CUB / Thrust / cuBLAS / cuSPARSE / CUTLASS): sort, scans,
    SpMV and reductions are this repo's own kernels."""
    csrc = os.path.join(ROOT, 'finite-element-solver_b200', 'csrc')
    for f in os.listdir(csrc):
        text = open(os.path.join(csrc, f)).read()
        for lib in ('cub/', 'cub::', 'thrust', 'cublas', 'cusparse', 'cutlass', 'cusolver'):
            assert lib not in text, (f, lib)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    mesh = create_rect_mesh([[0, 0], [1, 1]], (3, 3))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        fes_b200.FunctionSpace(mesh)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        fes_b200.JacobiPCGBackend().prepare(np.eye(3))
