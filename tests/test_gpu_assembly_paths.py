"""GPU parity of the fused assembly kernel's special paths against the CPU oracle: scattered node
numberings (tiles shrink until they reference <= 255 nodes), high-valence nodes (long pairs split
over lanes, and the untiled kernel beyond that), isolated nodes, duplicate and flipped elements,
a misaligned output pointer (plain-store path instead of the TMA bulk store), negative global
factors.  Tolerance: structure array_equal, values 1e-12 with the floor 1e-12*max|A|."""
import ctypes as C

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


def check_all_forms(v, e, okind, comps, seed=0, tiled=None):
    """Mass (every component count), Laplacian and elasticity with per-element moduli vs the oracle."""
    F = fes()
    d = v.shape[1]
    mesh = F.Mesh(v, e, np.zeros((0, e.shape[1] - 1), dtype=np.int64))
    geo = fo.geometry(okind, v[e])
    E = np.random.default_rng(seed).uniform(10.0, 300.0, len(e))
    for c in comps:
        V = F.FunctionSpace(mesh, n_components=c)
        ref = fo.assemble(e, c, len(v), fo.ke_mass(geo, c))
        A = V.assemble(F.MassForm(c))
        np.testing.assert_array_equal(A.indptr, ref.indptr)
        np.testing.assert_array_equal(A.indices, ref.indices)
        close(A.data, ref.data)
    V1 = F.FunctionSpace(mesh, n_components=1)
    close(V1.assemble(F.LaplacianForm()).data, fo.assemble(e, 1, len(v), fo.ke_laplacian(geo)).data)
    Vd = F.FunctionSpace(mesh, n_components=d)
    form = F.LinearElasticForm(F.LinearElasticMaterial(E, 0.3))
    A = Vd.assemble(form)
    ref = fo.assemble(e, d, len(v), fo.ke_elastic(geo, E, 0.3))
    np.testing.assert_array_equal(A.indptr, ref.indptr)
    np.testing.assert_array_equal(A.indices, ref.indices)
    close(A.data, ref.data)
    np.testing.assert_array_equal(A.data, Vd.assemble(form).data)      # bit-reproducible
    dense = A.toarray()
    if tiled is not None:
        assert (Vd.plan().n_tiles > 0) == tiled
    if Vd.plan().n_tiles > 0:
        np.testing.assert_array_equal(dense, dense.T)                  # (v,w) and (w,v) blocks: bitwise transposes
    else:
        close(dense, dense.T, rtol=1e-14)                              # untiled kernel: reference order, not mirrored
    return Vd, form, A


@pytest.mark.parametrize('kind,res', [('tri', (61, 47)), ('tet', (13, 12, 11))])
def test_scattered_numbering_shrinks_tiles(kind, res):
    """A random node permutation: consecutive nodes share nothing, so the first tiling attempt touches
    more than 255 distinct nodes per tile and the plan retries with smaller tiles."""
    rng = np.random.default_rng(17)
    if kind == 'tri':
        v, e = fo.rect_mesh([[0, 0], [3, 1]], res)
        okind = fo.P1_TRI
    else:
        v, e = fo.box_mesh([[0, 0, 0], [1, 1, 1]], res)
        okind = fo.P1_TET
    perm = rng.permutation(len(v))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(v))
    v, e = v[perm], inv[e][rng.permutation(len(e))]
    check_all_forms(v, e, okind, (1, v.shape[1]), seed=1, tiled=True)


@pytest.mark.parametrize('n_fan', [5, 40, 70, 300])
def test_high_valence_fan(n_fan):
    """A fan of n triangles around one node: its diagonal pair has n contributions, split over up to
    32 lanes (n <= 64) and handed to the untiled kernel beyond that; both must match the oracle."""
    ang = np.linspace(0.0, 2 * np.pi, n_fan, endpoint=False)
    rim = np.stack([np.cos(ang) * (1 + 0.2 * np.sin(3 * ang)), np.sin(ang)], axis=1)
    v = np.concatenate([[[0.05, -0.02]], rim])
    e = np.array([[0, 1 + k, 1 + (k + 1) % n_fan] for k in range(n_fan)], dtype=np.int64)
    check_all_forms(v, e, fo.P1_TRI, (1, 2), seed=n_fan, tiled=n_fan <= 64)


def test_isolated_nodes_and_duplicate_and_flipped_elements():
    v, e = fo.rect_mesh([[0, 0], [2, 1]], (19, 11))
    rng = np.random.default_rng(3)
    # unused vertices at the start, in the middle and at the end
    extra = rng.uniform(5, 6, (7, 2))
    v = np.concatenate([extra[:2], v[:100], extra[2:5], v[100:], extra[5:]])
    shift = np.where(np.arange(len(v) - 7) < 100, 2, 5)
    e = e + shift[e]
    e[::3] = e[::3][:, [0, 2, 1]]                        # negative orientation: |det J| as the reference
    e = np.concatenate([e, e[5:9]])                      # four elements listed twice
    Vd, form, A = check_all_forms(v, e, fo.P1_TRI, (1, 2), seed=2, tiled=True)
    ip = A.indptr
    for node in (0, 1, 102, 103, 104, len(v) - 2, len(v) - 1):
        assert ip[2 * node] == ip[2 * node + 1] == ip[2 * node + 2]    # empty rows


def test_empty_mesh_assembles_an_empty_operator():
    F = fes()
    v = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    mesh = F.Mesh(v, np.zeros((0, 3), dtype=np.int64), np.zeros((0, 2), dtype=np.int64))
    A = F.FunctionSpace(mesh, n_components=2).assemble(F.LinearElasticForm(F.LinearElasticMaterial(1.0, 0.3)))
    assert A.shape == (6, 6) and A.nnz == 0 and np.array_equal(A.indptr, np.zeros(7, dtype=np.int64))


@pytest.mark.parametrize('kind', ['tri', 'tet'])
def test_misaligned_output_takes_the_plain_store_path(kind):
    """`values` 8 bytes off a 16-byte boundary: the TMA bulk store cannot be used; the kernel's
    coalesced plain-store path must produce the same bits."""
    F = fes()
    from fes_b200 import _lib
    if kind == 'tri':
        mesh, et, c = F.create_rect_mesh([[0, 0], [2, 1]], (70, 33)), _lib.P1_TRI, 2
    else:
        mesh, et, c = F.create_box_mesh([[0, 0, 0], [1, 1, 1]], (12, 9, 11)), _lib.P1_TET, 3
    V = F.FunctionSpace(mesh, n_components=c)
    plan = V.plan()
    cf = _lib.Coeffs(200.0, 0.3, 1.0, None, None)
    buf = torch.zeros(plan.nnz + 1, dtype=torch.float64, device=V.device)
    out = {}
    for off in (0, 1):
        view = buf[off:off + plan.nnz]
        assert view.data_ptr() % 16 == 8 * off
        view.zero_()
        _lib.check(_lib.lib().fes_assemble(plan.handle, et, _lib.FORM_ELASTIC, c, _lib.ptr(V._en_d),
                                           _lib.ptr(V._coords_d), C.byref(cf), C.c_void_p(view.data_ptr()),
                                           _lib.stream()))
        out[off] = view.cpu().numpy().copy()
    np.testing.assert_array_equal(out[0], out[1])
    geo = fo.geometry(fo.P1_TRI if kind == 'tri' else fo.P1_TET, mesh.vertices[mesh.elements])
    close(out[1], fo.assemble(mesh.elements, c, len(mesh.vertices), fo.ke_elastic(geo, 200.0, 0.3)).data)


def test_negative_and_zero_factors():
    F = fes()
    mesh = F.create_rect_mesh([[0, 0], [1, 1]], (31, 29))
    V = F.FunctionSpace(mesh, n_components=2)
    form = F.LinearElasticForm(F.LinearElasticMaterial(50.0, 0.25))
    A = V.assemble(form).data
    np.testing.assert_array_equal(V.assemble(F.ScaledForm(-1.0, form)).data, -A)
    close(V.assemble(F.ScaledForm(-2.5, form)).data, -2.5 * A)
    assert not np.any(V.assemble(F.ScaledForm(0.0, form)).data)
    # a per-element modulus that is exactly zero on some elements (a masked region): no NaN from the
    # rsqrt scaling, the same values as the oracle
    Ez = np.full(len(mesh.elements), 50.0)
    Ez[::4] = 0.0
    geo = fo.geometry(fo.P1_TRI, mesh.vertices[mesh.elements])
    B = V.assemble(F.LinearElasticForm(F.LinearElasticMaterial(Ez, 0.25))).data
    assert np.all(np.isfinite(B))
    close(B, fo.assemble(mesh.elements, 2, len(mesh.vertices), fo.ke_elastic(geo, Ez, 0.25)).data)
    K0 = form.element_matrices(V)
    close(V.assemble(F.PrecomputedForm(K0, scale=Ez / 50.0)).data, B)


@pytest.mark.parametrize('kind', ['tri', 'tet', 'p2tri'])
def test_two_hundred_reassemblies_are_bit_identical(kind):
    """Race hunting without a sanitizer: the persistent kernel recycles its shared-memory rings, the
    value image and the coordinate buffer across tiles; any missing barrier or early TMA reuse shows up
    as run-to-run differences.  Mid-size meshes so every CTA walks many tiles."""
    F = fes()
    if kind == 'tri':
        V = F.FunctionSpace(F.create_rect_mesh([[0, 0], [3, 1]], (901, 301)), n_components=2)
    elif kind == 'tet':
        V = F.FunctionSpace(F.create_box_mesh([[0, 0, 0], [1, 1, 1]], (41, 41, 41)), n_components=3)
    else:
        V = F.FunctionSpace(F.create_rect_mesh([[0, 0], [1, 1]], (301, 301)), F.QuadraticTriangleElement, n_components=2)
    n_el = V._n_elements()
    E = torch.as_tensor(np.random.default_rng(0).uniform(1.0, 9.0, n_el), device=V.device)
    form = F.LinearElasticForm(F.LinearElasticMaterial(E, 0.3))
    ref = V.assemble_device(form).data_d.clone()
    out = torch.empty_like(ref)
    for rep in range(200):
        out.fill_(float('nan'))
        V.assemble_device(form, out=out)
        assert torch.equal(out, ref), f'reassembly {rep} differs'
    assert bool(torch.isfinite(ref).all())
