"""GPU parity of the stencil-run assembly kernel (k_assemble_runs): structured meshes of the reference's
generators (ref fem/mesh/structured.py:17-96) with jittered coordinates and per-element moduli, against
the CPU oracle (structure array_equal, values 1e-12 with the floor 1e-12*max|A|) and, bit for
bit, against the tiled kernel on the same plan (FES_NO_RUNS_LAUNCH=1 switches the run tiles off per call:
both kernels sum a pair's contributions in the same tree)."""
import os

import numpy as np
import pytest
import torch

import fes_oracle as fo

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def close(a, b, rtol=RTOL):
    scale = float(np.abs(b).max()) if np.size(b) else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * scale)


def fes():
    import fes_b200
    return fes_b200


def both_kernels(V, form):
    """values from the run kernel (+ remainder tiles) and from the tiled kernel alone"""
    a = V.assemble_device(form).data_d.cpu().numpy()
    os.environ['FES_NO_RUNS_LAUNCH'] = '1'
    try:
        b = V.assemble_device(form).data_d.cpu().numpy()
    finally:
        del os.environ['FES_NO_RUNS_LAUNCH']
    return a, b


def jitter(v, res, seed, amount=0.2):
    h = np.min((v.max(0) - v.min(0)) / (np.asarray(res) - 1))
    return v + np.random.default_rng(seed).uniform(-amount * h, amount * h, v.shape)


@pytest.mark.parametrize('res', [(41, 23), (140, 12), (33, 17), (300, 9), (64, 64)])
def test_triangle_runs_match_oracle_and_tiled_kernel(res):
    F = fes()
    v, e = fo.rect_mesh([[0, 0], [3, 1]], res)
    v = jitter(v, res, 3)
    mesh = F.Mesh(v, e, np.zeros((0, 2), dtype=np.int64))
    geo = fo.geometry(fo.P1_TRI, v[e])
    E = np.random.default_rng(5).uniform(10.0, 300.0, len(e))
    V1 = F.FunctionSpace(mesh, n_components=1)
    assert V1.plan().n_run_tiles > 0 and V1.plan().run_nodes > 0.8 * len(v)
    a, b = both_kernels(V1, F.LaplacianForm())
    np.testing.assert_array_equal(a, b)
    close(a, fo.assemble(e, 1, len(v), fo.ke_laplacian(geo)).data)
    V2 = F.FunctionSpace(mesh, n_components=2)
    assert V2.plan().n_run_tiles > 0
    for form, ref in ((F.LinearElasticForm(F.LinearElasticMaterial(E, 0.3)), fo.ke_elastic(geo, E, 0.3)),
                      (F.LinearElasticForm(F.LinearElasticMaterial(70.0, 0.25)), fo.ke_elastic(geo, np.full(len(e), 70.0), 0.25))):
        a, b = both_kernels(V2, form)
        np.testing.assert_array_equal(a, b)
        A = V2.assemble(form)
        R = fo.assemble(e, 2, len(v), ref)
        np.testing.assert_array_equal(A.indptr, R.indptr)
        np.testing.assert_array_equal(A.indices, R.indices)
        close(A.data, R.data)
        np.testing.assert_array_equal(A.data, a)


@pytest.mark.parametrize('res', [(7, 6, 45), (5, 5, 21), (6, 5, 80), (12, 11, 34)])
def test_tet_runs_match_oracle_and_tiled_kernel(res):
    F = fes()
    v, e = fo.box_mesh([[0, 0, 0], [1, 1, 2]], res)
    v = jitter(v, res, 4, 0.15)
    mesh = F.Mesh(v, e, np.zeros((0, 3), dtype=np.int64))
    geo = fo.geometry(fo.P1_TET, v[e])
    E = np.random.default_rng(6).uniform(10.0, 300.0, len(e))
    V1 = F.FunctionSpace(mesh, n_components=1)
    assert V1.plan().n_run_tiles > 0 and V1.plan().run_nodes > 0.8 * len(v)
    a, b = both_kernels(V1, F.LaplacianForm())
    np.testing.assert_array_equal(a, b)
    close(a, fo.assemble(e, 1, len(v), fo.ke_laplacian(geo)).data)
    V3 = F.FunctionSpace(mesh, n_components=3)
    form = F.LinearElasticForm(F.LinearElasticMaterial(E, 0.3))
    a, b = both_kernels(V3, form)
    np.testing.assert_array_equal(a, b)
    A = V3.assemble(form)
    R = fo.assemble(e, 3, len(v), fo.ke_elastic(geo, E, 0.3))
    np.testing.assert_array_equal(A.indptr, R.indptr)
    np.testing.assert_array_equal(A.indices, R.indices)
    close(A.data, R.data)


def test_runs_with_per_element_scale_and_negative_factor():
    """SIMP-style rho^p scaling (d_scale) and a ScaledForm with a negative factor through the run kernel."""
    F = fes()
    from fes_b200.forms import Lowered
    from fes_b200 import _lib
    res = (90, 30)
    v, e = fo.rect_mesh([[0, 0], [3, 1]], res)
    mesh = F.Mesh(v, e, np.zeros((0, 2), dtype=np.int64))
    geo = fo.geometry(fo.P1_TRI, v[e])
    V = F.FunctionSpace(mesh, n_components=2)
    rho = np.random.default_rng(8).uniform(0.0, 1.0, len(e)) ** 3
    rho[::17] = 0.0   # void elements: weight 0
    scale = torch.as_tensor(rho).to(V.device)

    class Scaled:
        def lower(self, space, boundary=False):
            return Lowered(_lib.FORM_ELASTIC, E=5.0, nu=0.3, factor=-2.0, scale_elem=scale)

    a, b = both_kernels(V, Scaled())
    np.testing.assert_array_equal(a, b)
    ref = fo.assemble(e, 2, len(v), -2.0 * rho[:, None, None] * fo.ke_elastic(geo, np.full(len(e), 5.0), 0.3))
    close(a, ref.data)


def test_unstructured_numbering_has_no_runs():
    F = fes()
    rng = np.random.default_rng(2)
    v, e = fo.rect_mesh([[0, 0], [1, 1]], (40, 40))
    perm = rng.permutation(len(v))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(v))
    mesh = F.Mesh(v[perm], inv[e], np.zeros((0, 2), dtype=np.int64))
    V = F.FunctionSpace(mesh, n_components=2)
    assert V.plan().n_run_tiles == 0 and V.plan().n_tiles > 0


def test_more_plans_than_template_slots_and_graph_replay():
    """The template programs of a plan with runs live in one of three __constant__ slots.  A fourth live plan finds
    none free and keeps the tiled kernel -- same results; a destroyed plan frees its slot.  And a run-path assembly
    (with an irregular region, i.e. remainder tiles forked onto the plan's side stream) is capturable in a CUDA
    graph: the replay writes the same bits."""
    import ctypes as C
    F = fes()
    from fes_b200 import _lib
    spaces, datas = [], []
    form = F.LinearElasticForm(F.LinearElasticMaterial(70.0, 0.25))
    for k in range(5):
        v, e = fo.rect_mesh([[0, 0], [2, 1]], (40 + 3 * k, 21 + k))
        mesh = F.Mesh(jitter(v, (40 + 3 * k, 21 + k), 10 + k), e, np.zeros((0, 2), dtype=np.int64))
        V = F.FunctionSpace(mesh, n_components=2)
        a, b = both_kernels(V, form)
        np.testing.assert_array_equal(a, b)
        spaces.append(V)
    with_runs = [V.plan().n_run_tiles > 0 for V in spaces]
    assert with_runs[:3] == [True, True, True] and with_runs[3:] == [False, False]
    del spaces[0]                                    # frees a slot
    import gc
    gc.collect()
    nx, ny = 50, 30
    v, e = fo.rect_mesh([[0, 0], [2, 1]], (nx, ny))
    # an irregular patch: the other diagonal in a block of 9 x 9 cells, so > 64 nodes fall out of every run and the
    # plan keeps remainder tiles (the tiled kernel on the plan's side stream) next to the run tiles
    e = e.copy()
    for i in range(20, 29):
        for j in range(10, 19):
            a, b, c, d = j * nx + i, j * nx + i + 1, (j + 1) * nx + i + 1, (j + 1) * nx + i
            k = 2 * (i * (ny - 1) + j)
            if (i + j) % 2 == 0:
                e[k], e[k + 1] = [a, b, d], [b, c, d]
            else:
                e[k], e[k + 1] = [a, b, c], [a, c, d]
    V = F.FunctionSpace(F.Mesh(v, e, np.zeros((0, 2), dtype=np.int64)), n_components=2)
    assert 0 < V.plan().n_run_tiles < V.plan().n_tiles and V.plan().run_nodes < len(v) - 64
    A = V.assemble_device(form)
    want = A.data_d.clone()
    ref = fo.assemble(e, 2, len(v), fo.ke_elastic(fo.geometry(fo.P1_TRI, v[e]), np.full(len(e), 70.0), 0.25))
    close(want.cpu().numpy(), ref.data)
    low = form.lower(V, False)
    cf = low.coeffs()
    L = _lib.lib()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        A.data_d.zero_()
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            _lib.check(L.fes_assemble(V.plan().handle, V.element_type.KIND, low.form, V.spatial_dim, _lib.ptr(V._en_d),
                                      _lib.ptr(V._coords_d), C.byref(cf), _lib.ptr(A.data_d), _lib.stream()))
        A.data_d.zero_()
        g.replay()
        side.synchronize()
    torch.cuda.current_stream().wait_stream(side)
    assert torch.equal(A.data_d, want)
