"""GPU parity: CSR pattern bit-exact, values within 1e-12 (relative, with the absolute floor
1e-12*max|A| of SURVEY A.5) against the reference's golden fixtures and the CPU oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

import fes_oracle as fo
import window as ow
from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def close(a, b, rtol=RTOL):
    scale = float(np.abs(b).max()) if np.size(b) else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * scale)


def fes():
    import fes_b200
    return fes_b200


ETYPES = {'tri': (None, fo.P1_TRI), 'tet': (None, fo.P1_TET), 'p2tri': ('QuadraticTriangleElement', fo.P2_TRI)}


def make_space(g, name, c):
    F = fes()
    et = getattr(F, ETYPES[name][0]) if ETYPES[name][0] else None
    mesh = F.Mesh(g['vertices'], g['elements'], g['boundary'])
    return F.FunctionSpace(mesh, et, n_components=c)


@pytest.mark.parametrize('name', ['tri', 'tet', 'p2tri'])
def test_assembly_matches_reference_fixtures(name):
    F = fes()
    g = load_golden('assembly_' + name)
    d = g['vertices'].shape[1]
    V1, Vd = make_space(g, name, 1), make_space(g, name, d)
    np.testing.assert_array_equal(V1.element_nodes, g['element_nodes'])
    np.testing.assert_array_equal(V1.boundary_nodes, g['boundary_nodes'])
    np.testing.assert_allclose(V1.element_volumes, g['volumes'], rtol=1e-13)
    mat = F.LinearElasticMaterial
    cases = {
        'lap': (V1, F.LaplacianForm(), False),
        'mass1': (V1, F.MassForm(1), False),
        'bmass1': (V1, F.MassForm(1), True),
        'massd': (Vd, F.MassForm(d), False),
        'bmassd': (Vd, F.MassForm(d), True),
        'bmaskd': (Vd, F.MaskedMassForm(d, g['facet_mask']), True),
        'elast': (Vd, F.LinearElasticForm(mat(200.0, 0.3)), False),
        'elastvar': (Vd, F.LinearElasticForm(mat(g['E_var'], 0.3)), False),
        'lapscaled': (V1, F.ScaledForm(2.25, F.LaplacianForm()), False),
    }
    for key, (V, form, bnd) in cases.items():
        A = V.assemble(form, boundary=bnd)
        assert A.indptr.dtype == np.int64 and A.indices.dtype == np.int64 and A.data.dtype == np.float64
        np.testing.assert_array_equal(A.indptr, g[key + '_indptr'], err_msg=key)
        np.testing.assert_array_equal(A.indices, g[key + '_indices'], err_msg=key)
        close(A.data, g[key + '_data'])
        close(form.element_matrices(V, bnd).cpu().numpy(), g[key + '_ke'])
        # PrecomputedForm fed the reference's own element matrices: same slots, same sum order
        B = V.assemble(F.PrecomputedForm(g[key + '_ke']), boundary=bnd)
        np.testing.assert_array_equal(B.data, g[key + '_data'], err_msg=key + ' (bincount order)')
    K0 = F.LinearElasticForm(mat(200.0, 0.3)).element_matrices(Vd)
    A = Vd.assemble(F.PrecomputedForm(K0, scale=g['rho'] ** 3))
    close(A.data, g['precomp_data'])
    np.testing.assert_array_equal(Vd.mass_matrix.indices, g['massd_indices'])
    close(Vd.boundary_mass_matrix.data, g['bmassd_data'])


def jittered(kind, seed):
    """A seeded, shuffled, jittered mesh: no structure for the pattern builder to lean on."""
    rng = np.random.default_rng(seed)
    if kind == 'tri':
        v, e = fo.rect_mesh([[0, 0], [2, 1]], (23, 17))
        h = np.array([2 / 22, 1 / 16])
        interior = (v[:, 0] > 0) & (v[:, 0] < 2) & (v[:, 1] > 0) & (v[:, 1] < 1)
    else:
        v, e = fo.box_mesh([[0, 0, 0], [1, 1.5, 0.7]], (7, 8, 6))
        h = np.array([1 / 6, 1.5 / 7, 0.7 / 5])
        interior = np.all((v > 0) & (v < np.array([1, 1.5, 0.7])), axis=1)
    v = v + interior[:, None] * rng.uniform(-0.25, 0.25, v.shape) * h
    perm = rng.permutation(len(v))                       # renumber nodes
    inv = np.empty_like(perm); inv[perm] = np.arange(len(v))
    v, e = v[perm], inv[e]
    e = e[rng.permutation(len(e))]                       # shuffle elements
    return v, e


@pytest.mark.parametrize('kind', ['tri', 'tet'])
def test_assembly_matches_oracle_on_unstructured_meshes(kind):
    F = fes()
    v, e = jittered(kind, seed=11)
    d = v.shape[1]
    okind = fo.P1_TRI if kind == 'tri' else fo.P1_TET
    mesh = F.Mesh(v, e)
    np.testing.assert_array_equal(mesh.boundary, fo.boundary_facets(e))
    geo = fo.geometry(okind, v[e])
    bgeo = fo.geometry(fo.facet_kind(okind), v[mesh.boundary])
    E = np.random.default_rng(5).uniform(10.0, 300.0, len(e))
    for c in (1, 2, d):
        V = F.FunctionSpace(mesh, n_components=c)
        ref = fo.assemble(e, c, len(v), fo.ke_mass(geo, c))
        A = V.assemble(F.MassForm(c))
        np.testing.assert_array_equal(A.indptr, ref.indptr)
        np.testing.assert_array_equal(A.indices, ref.indices)
        close(A.data, ref.data)
        refb = fo.assemble(mesh.boundary, c, len(v), fo.ke_mass(bgeo, c))
        B = V.assemble(F.MassForm(c), boundary=True)
        np.testing.assert_array_equal(B.indptr, refb.indptr)
        np.testing.assert_array_equal(B.indices, refb.indices)
        close(B.data, refb.data)
    V = F.FunctionSpace(mesh, n_components=1)
    close(V.assemble(F.LaplacianForm()).data, fo.assemble(e, 1, len(v), fo.ke_laplacian(geo)).data)
    V = F.FunctionSpace(mesh, n_components=d)
    A = V.assemble(F.LinearElasticForm(F.LinearElasticMaterial(E, 0.3)))
    ref = fo.assemble(e, d, len(v), fo.ke_elastic(geo, E, 0.3))
    np.testing.assert_array_equal(A.indices, ref.indices)
    close(A.data, ref.data)
    # determinism: bit-identical on reassembly
    np.testing.assert_array_equal(A.data, V.assemble(F.LinearElasticForm(F.LinearElasticMaterial(E, 0.3))).data)


def test_p2_assembly_matches_oracle():
    F = fes()
    v, e = jittered('tri', seed=3)
    mesh = F.Mesh(v, e)
    en, X, bn = fo.p2_nodes(v, e, mesh.boundary)
    geo = fo.geometry(fo.P2_TRI, X[en])
    V = F.FunctionSpace(mesh, F.QuadraticTriangleElement, n_components=2)
    np.testing.assert_array_equal(V.element_nodes, en)
    ref = fo.assemble(en, 2, len(X), fo.ke_elastic(geo, 200.0, 0.3))
    A = V.assemble(F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)))
    np.testing.assert_array_equal(A.indptr, ref.indptr)
    np.testing.assert_array_equal(A.indices, ref.indices)
    close(A.data, ref.data)
    refm = fo.assemble(en, 2, len(X), fo.ke_mass(geo, 2))
    close(V.mass_matrix.data, refm.data)
    refb = fo.assemble(bn, 2, len(X), fo.ke_mass(fo.geometry(fo.P2_LINE, X[bn]), 2))
    close(V.boundary_mass_matrix.data, refb.data)


def test_known_answers_and_invariants():
    """ref tests/test_assembly.py:68-101,113-175 and tests/test_forms.py:43-51,101-113."""
    F = fes()
    V = F.FunctionSpace(F.create_rect_mesh([[0, 0], [1, 1]], (2, 2)))
    np.testing.assert_allclose(V.assemble(F.LaplacianForm()).toarray(),
                               [[1, -.5, -.5, 0], [-.5, 1, 0, -.5], [-.5, 0, 1, -.5], [0, -.5, -.5, 1]], atol=1e-12)
    np.testing.assert_allclose(V.mass_matrix.toarray(),
                               [[1 / 6, 1 / 24, 1 / 24, 1 / 12], [1 / 24, 1 / 12, 0, 1 / 24],
                                [1 / 24, 0, 1 / 12, 1 / 24], [1 / 12, 1 / 24, 1 / 24, 1 / 6]], atol=1e-12)
    np.testing.assert_allclose(V.boundary_mass_matrix.toarray(),
                               [[2 / 3, 1 / 6, 1 / 6, 0], [1 / 6, 2 / 3, 0, 1 / 6],
                                [1 / 6, 0, 2 / 3, 1 / 6], [0, 1 / 6, 1 / 6, 2 / 3]], atol=1e-12)
    cube = F.create_box_mesh([[0, 0, 0], [1, 1, 1]], (3, 3, 3))
    V3 = F.FunctionSpace(cube, n_components=3)
    K = V3.assemble(F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3))).toarray()
    assert int((np.abs(K) > 1e-12 * np.abs(K).max()).sum()) == 1659
    assert np.trace(K) == pytest.approx(10153.846153846, rel=1e-10)
    assert np.linalg.norm(K) == pytest.approx(1660.0583982503, rel=1e-10)
    np.testing.assert_allclose(K, K.T, atol=1e-9)
    for comp in range(3):
        t = np.zeros(81); t[comp::3] = 1.0
        np.testing.assert_allclose(K @ t, 0, atol=1e-9)
    M = V3.mass_matrix.toarray()
    assert M.sum() == pytest.approx(3.0) and np.trace(M) == pytest.approx(1.2)
    assert V3.boundary_mass_matrix.toarray().sum() == pytest.approx(18.0)
    tri = F.FunctionSpace(F.Mesh([[0, 0], [1, 0], [0, 1]], [[0, 1, 2]]), n_components=2)
    Ke = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)).element_matrices(tri).cpu().numpy()[0]
    assert Ke[0, 0] == pytest.approx(173.076923, abs=1e-5) and Ke[0, 2] == pytest.approx(-134.615385, abs=1e-5)
    assert Ke[5, 5] == pytest.approx(134.615385, abs=1e-5) and Ke[3, 4] == pytest.approx(38.461538, abs=1e-5)


def test_error_behaviour_matches_the_reference():
    F = fes()
    V = F.FunctionSpace(F.create_rect_mesh([[0, 0], [1, 1]], (6, 6)), n_components=2)
    other = F.FunctionSpace(F.create_rect_mesh([[0, 0], [1, 1]], (8, 8)), n_components=2)
    K0 = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)).element_matrices(V)
    with pytest.raises(ValueError, match='elements'):
        other.assemble(F.PrecomputedForm(K0))
    with pytest.raises(ValueError, match='2 entries but the mesh has'):
        V.assemble(F.LinearElasticForm(F.LinearElasticMaterial(np.array([100.0, 200.0]), 0.3)))
    with pytest.raises(ValueError):
        V.assemble(F.LaplacianForm())               # scalar form on a 2-component space
    with pytest.raises(NotImplementedError):
        F.FunctionSpace(F.Mesh([[0.0], [1.0]], [[0, 1]], np.zeros((0, 1), dtype=int)))


def test_c_abi_host_buffer_entry_point():
    """fes_assemble_host: host coordinates in, host CSR values out, through the raw C ABI."""
    F = fes()
    from fes_b200 import _lib
    g = load_golden('assembly_tet')
    V = make_space(g, 'tet', 3)
    plan = V.plan()
    vals = np.empty(plan.nnz)
    cf = _lib.Coeffs(1.0, 0.3, 1.0, None, None)
    E = np.ascontiguousarray(g['E_var'])
    coords = np.ascontiguousarray(g['vertices'])
    _lib.check(_lib.lib().fes_assemble_host(plan.handle, _lib.P1_TET, _lib.FORM_ELASTIC, 3, _lib.ptr(V._en_d),
                                            _lib.ptr(coords), len(coords), C.byref(cf), _lib.ptr(E), None,
                                            _lib.ptr(vals), None))
    close(vals, g['elastvar_data'])
    ip, ix = plan.host_pattern()
    np.testing.assert_array_equal(ip, g['elastvar_indptr'])
    np.testing.assert_array_equal(ix, g['elastvar_indices'])


def test_full_size_cantilever_properties():
    """Config 2 of BASELINE.json (2D P1 elasticity, 3 998 792 triangles): properties that hold at
    any size -- CSR well-formed, rigid translations in the null space, symmetry via x^T A y."""
    F = fes()
    mesh = F.create_rect_mesh([[0, 0], [4, 1]], (2829, 708))
    V = F.FunctionSpace(mesh, n_components=2)
    A = V.assemble_device(F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)))
    n = V.n_dofs
    assert A.nnz == 56026604 or A.nnz > 5.5e7
    ip = A.indptr_d
    assert int(ip[0]) == 0 and int(ip[-1]) == A.nnz and bool((ip[1:] > ip[:-1]).all())
    ix = A.indices_d.to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(n, device=ix.device), (ip[1:] - ip[:-1]).to(torch.int64))
    key = rows * n + ix
    assert bool((key[1:] > key[:-1]).all())               # columns strictly ascending in every row
    scale = float(A.data_d.abs().max())
    for comp in range(2):
        t = torch.zeros(n, dtype=torch.float64, device=ix.device); t[comp::2] = 1.0
        assert float(A.matvec(t).abs().max()) < 1e-9 * scale
    gen = torch.Generator(device='cpu').manual_seed(0)
    x = torch.rand(n, generator=gen, dtype=torch.float64).to(ix.device)
    y = torch.rand(n, generator=gen, dtype=torch.float64).to(ix.device)
    assert abs(float(x @ A.matvec(y)) - float(y @ A.matvec(x))) < 1e-9 * float(x @ A.matvec(x))
    M = V.mass_matrix_d
    assert float(M.data_d.sum()) == pytest.approx(2 * 4.0, rel=1e-12)
    # the same rows assembled by the oracle on a sub-mesh window agree with the full operator
    sub_v, sub_e = fo.rect_mesh([[0, 0], [4, 1]], (2829, 708))
    win = np.arange(0, 4000)                                # first 4000 elements
    Ke_o = fo.ke_elastic(fo.geometry(fo.P1_TRI, sub_v[sub_e[win]]), 200.0, 0.3)
    Ke_g = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)).element_matrices(V)[win].cpu().numpy()
    close(Ke_g, Ke_o)
    # the FUSED kernel's rows against the oracle at this size: every row of three node windows (the first
    # mesh rows with the y = 0 boundary, the middle, the last rows) -- row lengths and column ids
    # array_equal, values 1e-12 with the floor 1e-12 max|A| (ref tests/test_assembly.py:196-218)
    nx, n_nodes = 2829, 2829 * 708
    ke = lambda geo, ids: fo.ke_elastic(geo, 200.0, 0.3)
    for v0 in (0, (n_nodes // 2 // nx) * nx - 17, n_nodes - 3 * nx):
        ow.compare_window(A, fo.P1_TRI, 2, sub_v, sub_e, v0, min(v0 + 3 * nx, n_nodes), ke, scale=scale)


def test_full_size_p2_properties():
    """Config 3 of BASELINE.json (P2 elasticity + mass on 708^2 vertices, 999 698 six-node triangles): the
    fused kernel (two inverse-Jacobian rows per record contracted with the per-node-pair table) against
    size-independent properties -- rigid translations AND the rigid rotation (a linear field, exactly
    representable in P2) in the null space, symmetry via x^T A y, the mass of the domain -- and against the
    unfused element-matrix kernel and the oracle on a window of elements."""
    F = fes()
    mesh = F.create_rect_mesh([[0, 0], [1, 1]], (708, 708))
    V = F.FunctionSpace(mesh, F.QuadraticTriangleElement, n_components=2)
    form = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3))
    A = V.assemble_device(form)
    n = V.n_dofs
    assert V._n_elements() == 999698 and A.nnz == 4 * V.plan().n_pairs
    scale = float(A.data_d.abs().max())
    X = torch.as_tensor(V.node_coords, device=A.device)
    for t in (torch.stack([torch.ones_like(X[:, 0]), torch.zeros_like(X[:, 0])], 1),
              torch.stack([torch.zeros_like(X[:, 0]), torch.ones_like(X[:, 0])], 1),
              torch.stack([-X[:, 1], X[:, 0]], 1)):
        assert float(A.matvec(t.reshape(-1).contiguous()).abs().max()) < 1e-9 * scale
    gen = torch.Generator(device='cpu').manual_seed(1)
    x = torch.rand(n, generator=gen, dtype=torch.float64).to(A.device)
    y = torch.rand(n, generator=gen, dtype=torch.float64).to(A.device)
    assert abs(float(x @ A.matvec(y)) - float(y @ A.matvec(x))) < 1e-9 * float(x @ A.matvec(x))
    assert float(V.mass_matrix_d.data_d.sum()) == pytest.approx(2 * 1.0, rel=1e-12)
    # fused assembly == deterministic scatter of the unfused element matrices (different kernels, same operator)
    Ke = form.element_matrices(V)
    B = V.assemble_device(F.PrecomputedForm(Ke))
    assert float((A.data_d - B.data_d).abs().max()) < 1e-12 * scale
    win = np.arange(500000, 502000)
    geo = fo.geometry(fo.P2_TRI, V.node_coords[V.element_nodes[win]])
    close(Ke[win].cpu().numpy(), fo.ke_elastic(geo, 200.0, 0.3))
    # the fused kernel's rows (stiffness and mass) against the oracle: windows of vertex nodes (first rows,
    # middle) and of edge nodes (numbered after the 708^2 vertices)
    nv = 708 * 708
    Mm = V.mass_matrix_d
    for v0, v1 in ((0, 3 * 708), (nv // 2, nv // 2 + 2000), (nv + 700000, nv + 703000), (V.n_nodes - 2500, V.n_nodes)):
        ow.compare_window(A, fo.P2_TRI, 2, V.node_coords, V.element_nodes, v0, v1,
                          lambda g, ids: fo.ke_elastic(g, 200.0, 0.3), scale=scale)
        ow.compare_window(Mm, fo.P2_TRI, 2, V.node_coords, V.element_nodes, v0, v1, lambda g, ids: fo.ke_mass(g, 2))


def test_full_size_tet_pattern_properties():
    """Config 5 on one GPU at its full size, 151^3 nodes = 20 250 000 tets: pattern counts follow the closed
    form for a Kuhn mesh, the operator annihilates rigid translations, and the fused kernel's rows agree with
    the oracle on node windows (one node plane at x = 0 with its boundary stencils, half a plane in the
    middle of the box, the last lines of the box)."""
    F = fes()
    from fes_b200.mesh import box_arrays
    n = 151
    verts, elems = box_arrays([[0, 0, 0], [1, 1, 1]], (n, n, n))
    mesh = F.Mesh(verts, elems, np.zeros((0, 3), dtype=np.int64))   # boundary facets are not needed here
    V = F.FunctionSpace(mesh, n_components=3)
    A = V.assemble_device(F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3)))
    ke = lambda geo, ids: fo.ke_elastic(geo, 200.0, 0.3)
    smax = float(A.data_d.abs().max())
    for v0, v1 in ((0, n * n), (n ** 3 // 2, n ** 3 // 2 + n * n // 2), (n ** 3 - 20 * n, n ** 3)):
        ow.compare_window(A, fo.P1_TET, 3, verts, elems, v0, v1, ke, scale=smax)
    # Kuhn mesh: node pairs = nodes + 2*edges; edges = 3 axis + 3 face-diagonal + 1 body-diagonal families
    m = n - 1
    edges = 3 * n * n * m + 3 * n * m * m + m ** 3
    assert V.plan().n_pairs == n ** 3 + 2 * edges
    assert A.nnz == 9 * V.plan().n_pairs
    scale = smax
    for comp in range(3):
        t = torch.zeros(V.n_dofs, dtype=torch.float64, device=A.device); t[comp::3] = 1.0
        assert float(A.matvec(t).abs().max()) < 1e-9 * scale


@pytest.mark.parametrize('kind,world', [('tri', 2), ('tri', 3), ('tet', 4)])
def test_partitioned_row_blocks_concatenate_to_the_single_gpu_operator(kind, world):
    """every rank assembles the rows it owns from its own + ghost elements with no
    exchange; the blocks concatenate to the one-GPU CSR -- structure bit-exact, values identical
    (same contributions in the same order).  Ranks are emulated one after another on one GPU."""
    F = fes()
    from fes_b200.dist import NodePartition, owned_row_block
    v, e = jittered(kind, seed=21)
    d = v.shape[1]
    E = np.random.default_rng(2).uniform(10.0, 300.0, len(e))
    whole = F.FunctionSpace(F.Mesh(v, e), n_components=d).assemble(
        F.LinearElasticForm(F.LinearElasticMaterial(E, 0.3)))
    ips, ixs, datas = [], [], []
    for r in range(world):
        part = NodePartition(e, len(v), r, world)
        Vl = F.FunctionSpace(part.local_mesh(v), n_components=d)
        Al = Vl.assemble(F.LinearElasticForm(F.LinearElasticMaterial(E[part.local_elements], 0.3)))
        ip, ix, data = owned_row_block(part, Al, d)
        ips.append(ip); ixs.append(ix); datas.append(data)
    off = np.cumsum([0] + [ip[-1] for ip in ips[:-1]])
    indptr = np.concatenate([[0]] + [ip[1:] + o for ip, o in zip(ips, off)])
    np.testing.assert_array_equal(indptr, whole.indptr)
    np.testing.assert_array_equal(np.concatenate(ixs), whole.indices)
    np.testing.assert_array_equal(np.concatenate(datas), whole.data)


@pytest.mark.parametrize('case', ['tri_1001x251', 'p2_355', 'tet_41'])
def test_whole_operator_against_the_oracle_at_survey_sizes(case):
    """The largest meshes the oracle assembles whole in seconds (BASELINE.md 2.1 sizes): the complete
    CSR of the fused kernel -- indptr / indices array_equal, every value 1e-12 + floor."""
    F = fes()
    mat = F.LinearElasticMaterial(200.0, 0.3)
    if case == 'tri_1001x251':
        v, e = fo.rect_mesh([[0, 0], [4, 1]], (1001, 251))
        V = F.FunctionSpace(F.Mesh(v, e, np.zeros((0, 2), dtype=np.int64)), n_components=2)
        en, X, okind, c = e, v, fo.P1_TRI, 2
    elif case == 'p2_355':
        mesh = F.create_rect_mesh([[0, 0], [1, 1]], (355, 355))
        V = F.FunctionSpace(mesh, F.QuadraticTriangleElement, n_components=2)
        en, X, okind, c = V.element_nodes, V.node_coords, fo.P2_TRI, 2
    else:
        v, e = fo.box_mesh([[0, 0, 0], [1, 1, 1]], (41, 41, 41))
        V = F.FunctionSpace(F.Mesh(v, e, np.zeros((0, 3), dtype=np.int64)), n_components=3)
        en, X, okind, c = e, v, fo.P1_TET, 3
    geo = fo.geometry(okind, X[en])
    ref = fo.assemble(en, c, len(X), fo.ke_elastic(geo, 200.0, 0.3))
    A = V.assemble(F.LinearElasticForm(mat))
    np.testing.assert_array_equal(A.indptr, ref.indptr)
    np.testing.assert_array_equal(A.indices, ref.indices)
    close(A.data, ref.data)
    if case != 'tet_41':
        refm = fo.assemble(en, c, len(X), fo.ke_mass(geo, c))
        close(V.mass_matrix.data, refm.data)
