"""CPU oracle for the fes-b200 hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/scipy restatement of the algorithms the CUDA path replaces, written from
the behaviour of janetyq/Finite-Element-Solver (citations `ref:` are file:line into the
upstream tree).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module.  The product package
(`finite-element-solver_b200/`, importable as `fes_b200`) never does, and raises if its
CUDA library is missing instead of falling back to anything in here.

Parity status: PINNED.  `oracle/gen_golden.py` imports the real reference in the build
container and writes `tests/golden/*.npz`; `tests/test_oracle_vs_golden.py` checks every
function below against those fixtures and against the reference's own known-answer
tests (ref: tests/test_forms.py:43-113, tests/test_assembly.py:68-139,196-218,
tests/test_topology_optimization.py:21-31, tests/test_system.py:7-29).

Conventions (ref: SURVEY.md A.2): fp64 arithmetic, int64 indices, DOF of (node v,
component d) = c*v + d, element-matrix flat order (e, i, j) row-major.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from scipy.sparse import csr_array

# --------------------------------------------------------------------------------------
# element kinds
# --------------------------------------------------------------------------------------
P1_LINE, P1_TRI, P1_TET, P2_LINE, P2_TRI = 'p1_line', 'p1_tri', 'p1_tet', 'p2_line', 'p2_tri'

_N_NODES = {P1_LINE: 2, P1_TRI: 3, P1_TET: 4, P2_LINE: 3, P2_TRI: 6}
_REF_DIM = {P1_LINE: 1, P1_TRI: 2, P1_TET: 3, P2_LINE: 1, P2_TRI: 2}
_DEGREE = {P1_LINE: 1, P1_TRI: 1, P1_TET: 1, P2_LINE: 2, P2_TRI: 2}
_FACET = {P1_TRI: P1_LINE, P1_TET: P1_TRI, P2_TRI: P2_LINE}


def n_nodes(kind):
    return _N_NODES[kind]


def ref_dim(kind):
    return _REF_DIM[kind]


def facet_kind(kind):
    return _FACET[kind]


# --------------------------------------------------------------------------------------
# quadrature (ref: fem/quadrature.py:54-114)
# --------------------------------------------------------------------------------------
_S3, _S35 = 0.5 / math.sqrt(3.0), 0.5 * math.sqrt(3.0 / 5.0)
_DA, _DB = 0.44594849091596489, 0.10810301816807022
_DC, _DD = 0.09157621350977074, 0.81684757298045851
_KA, _KB = 0.5854101966249685, 0.1381966011250105

# (points, weights, exactness degree), cheapest first, per reference-simplex dimension
_QUAD = {
    1: [
        ([[0.5]], [1.0], 1),
        ([[0.5 - _S3], [0.5 + _S3]], [0.5, 0.5], 3),
        ([[0.5 - _S35], [0.5], [0.5 + _S35]], [5 / 18, 8 / 18, 5 / 18], 5),
    ],
    2: [
        ([[1 / 3, 1 / 3]], [0.5], 1),
        ([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]], [1 / 6] * 3, 2),
        ([[_DA, _DA], [_DB, _DA], [_DA, _DB], [_DC, _DC], [_DD, _DC], [_DC, _DD]],
         [0.11169079483900573] * 3 + [0.05497587182766094] * 3, 4),
    ],
    3: [
        ([[0.25, 0.25, 0.25]], [1 / 6], 1),
        ([[_KA, _KB, _KB], [_KB, _KA, _KB], [_KB, _KB, _KA], [_KB, _KB, _KB]], [1 / 24] * 4, 2),
    ],
}


def quadrature(dim, min_degree):
    """Cheapest tabulated rule on the `dim`-simplex exact to `min_degree`."""
    for pts, wts, deg in _QUAD.get(dim, []):
        if deg >= min_degree:
            return np.array(pts, dtype=float), np.array(wts, dtype=float), deg
    raise NotImplementedError(f'no rule of degree >= {min_degree} on the {dim}-simplex')


def stiffness_degree(kind):
    """Rule degree a plain stiffness uses (ref: fem/elements.py:78-86)."""
    return max(1, 2 * (_DEGREE[kind] - 1))


# --------------------------------------------------------------------------------------
# shape functions (ref: fem/elements.py:258-297 P1, 329-348 P2 line, 369-396 P2 tri)
# --------------------------------------------------------------------------------------
def shape_values(kind, pts):
    pts = np.atleast_2d(np.asarray(pts, dtype=float))
    if _DEGREE[kind] == 1:
        return np.hstack([1.0 - pts.sum(1, keepdims=True), pts])
    if kind == P2_LINE:
        a, b = 1.0 - pts[:, 0], pts[:, 0]
        return np.column_stack([a * (2 * a - 1), b * (2 * b - 1), 4 * a * b])
    x, y = pts[:, 0], pts[:, 1]
    l0 = 1.0 - x - y
    return np.column_stack([l0 * (2 * l0 - 1), x * (2 * x - 1), y * (2 * y - 1),
                            4 * x * y, 4 * l0 * y, 4 * l0 * x])


def shape_gradients(kind, pts):
    """(n_pts, N, ref_dim) reference gradients."""
    pts = np.atleast_2d(np.asarray(pts, dtype=float))
    npt, N, r = len(pts), _N_NODES[kind], _REF_DIM[kind]
    g = np.zeros((npt, N, r))
    if _DEGREE[kind] == 1:
        g[:, 0, :] = -1.0
        for i in range(r):
            g[:, i + 1, i] = 1.0
        return g
    if kind == P2_LINE:
        x = pts[:, 0]
        g[:, 0, 0], g[:, 1, 0], g[:, 2, 0] = 4 * x - 3, 4 * x - 1, 4 - 8 * x
        return g
    x, y = pts[:, 0], pts[:, 1]
    l0 = 1.0 - x - y
    g[:, 0, 0] = g[:, 0, 1] = -(4 * l0 - 1)
    g[:, 1, 0] = 4 * x - 1
    g[:, 2, 1] = 4 * y - 1
    g[:, 3, 0], g[:, 3, 1] = 4 * y, 4 * x
    g[:, 4, 0], g[:, 4, 1] = -4 * y, 4 * l0 - 4 * y
    g[:, 5, 0], g[:, 5, 1] = 4 * l0 - 4 * x, -4 * x
    return g


def reference_mass(kind):
    """Mass matrix per unit measure (ref: fem/elements.py:113-126, 267-274)."""
    N, r = _N_NODES[kind], _REF_DIM[kind]
    if _DEGREE[kind] == 1:
        return (np.ones((N, N)) + np.eye(N)) / (N * (N + 1))
    pts, wts, _ = quadrature(r, 2 * _DEGREE[kind])
    phi = shape_values(kind, pts)
    return (phi.T * wts) @ phi * math.factorial(r)


# --------------------------------------------------------------------------------------
# affine geometry (ref: fem/elements.py:128-212)
# --------------------------------------------------------------------------------------
@dataclass
class Geometry:
    kind: str
    shape: np.ndarray        # (n_qp, N)
    grad: np.ndarray         # (n_el, n_qp, N, sd)
    wdet: np.ndarray         # (n_el, n_qp)
    points: np.ndarray       # (n_el, n_qp, sd)

    @property
    def volumes(self):
        return self.wdet.sum(1)


def geometry(kind, X, degree=None):
    """Batched geometry of straight-sided elements with node coordinates X (n_el, N, sd)."""
    X = np.asarray(X, dtype=float)
    r = _REF_DIM[kind]
    if X.ndim != 3 or X.shape[1] != _N_NODES[kind]:
        raise ValueError(f'{kind} expects (n_el, {_N_NODES[kind]}, sd) coordinates, got {X.shape}')
    pts, wts, _ = quadrature(r, stiffness_degree(kind) if degree is None else degree)
    dref = shape_gradients(kind, pts)
    phi = shape_values(kind, pts)
    corners = X[:, :r + 1]
    J = np.transpose(corners[:, 1:] - corners[:, :1], (0, 2, 1))      # (n_el, sd, r)
    if J.shape[1] == J.shape[2]:
        Jinv = np.linalg.inv(J)
        scale = np.abs(np.linalg.det(J))
    else:                                                              # embedded facet
        Jinv = np.linalg.pinv(J)
        scale = np.sqrt(np.abs(np.linalg.det(np.transpose(J, (0, 2, 1)) @ J)))
    grad = np.einsum('qnr,ers->eqns', dref, Jinv)
    wdet = scale[:, None] * wts[None, :]
    return Geometry(kind, phi, grad, wdet, np.einsum('qn,ens->eqs', phi, X))


# --------------------------------------------------------------------------------------
# materials + element matrices (ref: fem/materials.py:24-28,38-66,110-128; fem/forms.py)
# --------------------------------------------------------------------------------------
def lame(E, nu):
    return E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))


def hooke(dim, mu, lam):
    """Voigt D (plane strain in 2D), strain order xx,yy,(zz),xy,(yz),(xz)."""
    ns = dim * (dim - 1) // 2
    D = np.zeros((dim + ns, dim + ns))
    D[:dim, :dim] = lam
    D[np.arange(dim), np.arange(dim)] += 2 * mu
    D[np.arange(dim, dim + ns), np.arange(dim, dim + ns)] = mu
    return D


def voigt_B(grad):
    """(..., N, dim) gradients -> (..., n_strain, N*dim) (ref: fem/forms.py:37-71)."""
    *lead, N, dim = grad.shape
    if dim == 2:
        B = np.zeros((*lead, 3, 2 * N))
        gx, gy = grad[..., 0], grad[..., 1]
        B[..., 0, 0::2] = gx
        B[..., 1, 1::2] = gy
        B[..., 2, 0::2] = gy
        B[..., 2, 1::2] = gx
        return B
    if dim == 3:
        B = np.zeros((*lead, 6, 3 * N))
        gx, gy, gz = grad[..., 0], grad[..., 1], grad[..., 2]
        B[..., 0, 0::3] = gx
        B[..., 1, 1::3] = gy
        B[..., 2, 2::3] = gz
        B[..., 3, 0::3], B[..., 3, 1::3] = gy, gx
        B[..., 4, 1::3], B[..., 4, 2::3] = gz, gy
        B[..., 5, 0::3], B[..., 5, 2::3] = gz, gx
        return B
    raise NotImplementedError(f'no Voigt B for dim={dim}')


def ke_laplacian(geo):
    """ref: fem/forms.py:254-258."""
    return np.einsum('eqid,eqjd,eq->eij', geo.grad, geo.grad, geo.wdet)


def ke_mass(geo, c=1, mask=None):
    """vol * kron(M_ref, I_c), optionally masked per element (ref: fem/forms.py:202-209,245-247)."""
    block = np.kron(reference_mass(geo.kind), np.eye(c))
    out = geo.volumes[:, None, None] * block
    if mask is not None:
        out = out * np.asarray(mask, dtype=float)[:, None, None]
    return out


def ke_elastic(geo, E, nu):
    """sum_q w_q B^T D B with scalar or per-element E (ref: fem/forms.py:323-333)."""
    dim = _REF_DIM[geo.kind]
    n_el = geo.grad.shape[0]
    B = voigt_B(geo.grad)
    if isinstance(E, np.ndarray):
        if len(E) != n_el:
            raise ValueError(f'per-element modulus has {len(E)} entries but the mesh has {n_el} elements')
        mu, lam = lame(E, nu)
        D = mu[:, None, None] * hooke(dim, 1.0, 0.0) + lam[:, None, None] * hooke(dim, 0.0, 1.0)
    else:
        mu, lam = lame(E, nu)
        D = np.broadcast_to(mu * hooke(dim, 1.0, 0.0) + lam * hooke(dim, 0.0, 1.0),
                            (n_el, dim + dim * (dim - 1) // 2, dim + dim * (dim - 1) // 2))
    return np.einsum('eqji,ejk,eqkl,eq->eil', B, D, B, geo.wdet, optimize=True)


def ke_elastic_closed_form(geo, E, nu):
    """The isotropic closed form the CUDA kernels evaluate:
    K[(a,i),(b,j)] = sum_q w_q (lam g_ai g_bj + mu g_aj g_bi + mu (g_a.g_b) delta_ij)."""
    mu, lam = lame(np.asarray(E, dtype=float), nu)
    mu = np.broadcast_to(mu, (geo.grad.shape[0],))
    lam = np.broadcast_to(lam, (geo.grad.shape[0],))
    g, w = geo.grad, geo.wdet
    n_el, _, N, d = g.shape
    t1 = np.einsum('eq,eqai,eqbj->eaibj', w, g, g) * lam[:, None, None, None, None]
    t2 = np.einsum('eq,eqaj,eqbi->eaibj', w, g, g) * mu[:, None, None, None, None]
    dot = np.einsum('eq,eqas,eqbs->eab', w, g, g) * mu[:, None, None]
    t3 = np.einsum('eab,ij->eaibj', dot, np.eye(d))
    return (t1 + t2 + t3).reshape(n_el, N * d, N * d)


# --------------------------------------------------------------------------------------
# DOF numbering + CSR scatter (ref: fem/space.py:55-66,107-145,757-772)
# --------------------------------------------------------------------------------------
def dof_map(elem_nodes, c):
    elem_nodes = np.asarray(elem_nodes)
    return (c * elem_nodes[..., None] + np.arange(c)).reshape(*elem_nodes.shape[:-1], -1)


@dataclass
class ScatterPlan:
    order: np.ndarray
    group: np.ndarray
    indices: np.ndarray
    indptr: np.ndarray
    n: int

    def scatter(self, Ke):
        if Ke.size != len(self.order):
            raise ValueError(f'expected element matrices covering {len(self.order)} entries, got {Ke.size}')
        data = np.bincount(self.group, weights=Ke.ravel()[self.order], minlength=len(self.indices))
        return csr_array((data, self.indices, self.indptr), shape=(self.n, self.n))


def scatter_plan(elem_nodes, c, n_nodes_total):
    """Every (row, col) any element touches, explicit zeros kept, columns ascending."""
    dofs = dof_map(elem_nodes, c)
    k = dofs.shape[1]
    n = c * n_nodes_total
    key = (np.repeat(dofs, k, axis=1) * n + np.tile(dofs, (1, k))).ravel()
    order = np.argsort(key, kind='stable')
    skey = key[order]
    head = np.ones(len(skey), dtype=bool)
    head[1:] = skey[1:] != skey[:-1]
    slots = skey[head]
    counts = np.bincount(slots // n, minlength=n)
    return ScatterPlan(order, np.cumsum(head) - 1, slots % n,
                       np.concatenate([[0], np.cumsum(counts)]), n)


def csr_pattern_nodewise(elem_nodes, c, n_nodes_total):
    """Same pattern via node-level adjacency + block expansion -- the
    construction the CUDA pattern builder uses; tested equal to `scatter_plan`."""
    en = np.asarray(elem_nodes, dtype=np.int64)
    N = en.shape[1]
    key = np.unique((np.repeat(en, N, axis=1) * n_nodes_total + np.tile(en, (1, N))).ravel())
    v, w = key // n_nodes_total, key % n_nodes_total
    deg = np.bincount(v, minlength=n_nodes_total)
    nptr = np.concatenate([[0], np.cumsum(deg)])
    indptr = np.zeros(c * n_nodes_total + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(np.repeat(c * deg, c))
    indices = np.empty(indptr[-1], dtype=np.int64)
    for d in range(c):
        # row c*v+d holds, for each neighbour w in ascending order, columns c*w .. c*w+c-1
        start = indptr[c * v + d] + c * (np.arange(len(key)) - nptr[v])
        for j in range(c):
            indices[start + j] = c * w + j
    return indptr, indices


def assemble(elem_nodes, c, n_nodes_total, Ke):
    return scatter_plan(elem_nodes, c, n_nodes_total).scatter(Ke)


# --------------------------------------------------------------------------------------
# structured meshes (ref: fem/mesh/structured.py:17-88, fem/geometry.py:131-144)
# --------------------------------------------------------------------------------------
def rect_mesh(corners, resolution):
    nx, ny = resolution
    xs = np.linspace(corners[0][0], corners[1][0], nx)
    ys = np.linspace(corners[0][1], corners[1][1], ny)
    verts = np.stack([np.tile(xs, ny), np.repeat(ys, nx)], axis=1)
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing='ij')
    i, j = i.ravel(), j.ravel()
    v00, v10, v11, v01 = j * nx + i, j * nx + i + 1, (j + 1) * nx + i + 1, (j + 1) * nx + i
    even = ((i + j) % 2 == 0)[:, None]
    t0 = np.where(even, np.stack([v00, v10, v11], 1), np.stack([v00, v10, v01], 1))
    t1 = np.where(even, np.stack([v00, v11, v01], 1), np.stack([v10, v11, v01], 1))
    elems = np.stack([t0, t1], axis=1).reshape(-1, 3)
    return verts, elems


_KUHN = np.array([(0, 1, 3, 7), (0, 1, 5, 7), (0, 2, 3, 7), (0, 2, 6, 7), (0, 4, 5, 7), (0, 4, 6, 7)])


def box_mesh(corners, resolution):
    nx, ny, nz = resolution
    xs = np.linspace(corners[0][0], corners[1][0], nx)
    ys = np.linspace(corners[0][1], corners[1][1], ny)
    zs = np.linspace(corners[0][2], corners[1][2], nz)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing='ij')
    verts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    i, j, k = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing='ij')
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    corner = np.stack([((i + (c >> 2 & 1)) * ny + (j + (c >> 1 & 1))) * nz + (k + (c & 1))
                       for c in range(8)], axis=1)
    return verts, corner[:, _KUHN].reshape(-1, 4)


def boundary_facets(elems):
    """Facets belonging to exactly one element, ids sorted, first-appearance order."""
    elems = np.sort(np.asarray(elems), axis=1)
    N = elems.shape[1]
    # itertools.combinations(sorted(el), N-1) drops node N-1, then N-2, ..., then 0
    facets = np.stack([np.delete(elems, N - 1 - m, axis=1) for m in range(N)], axis=1).reshape(-1, N - 1)
    _, first, inverse, counts = np.unique(facets, axis=0, return_index=True, return_inverse=True,
                                          return_counts=True)
    single = counts[inverse.ravel()] == 1
    return facets[np.sort(first[counts == 1])] if single.any() else np.zeros((0, N - 1), dtype=elems.dtype)


def p2_nodes(verts, elems, boundary):
    """P2 triangle connectivity (ref: fem/space.py:205-252, fem/mesh/mesh.py:152-183)."""
    verts, elems, boundary = np.asarray(verts, float), np.asarray(elems), np.asarray(boundary)
    nv = len(verts)
    pairs = np.sort(np.concatenate([elems[:, [1, 2]], elems[:, [0, 2]], elems[:, [0, 1]]]), axis=1)
    key = pairs[:, 0] * nv + pairs[:, 1]
    ukey = np.unique(key)
    edges = np.stack([ukey // nv, ukey % nv], axis=1)
    mids = nv + np.searchsorted(ukey, key).reshape(3, -1).T
    bsort = np.sort(boundary, axis=1)
    bmid = nv + np.searchsorted(ukey, bsort[:, 0] * nv + bsort[:, 1])
    coords = np.vstack([verts, verts[edges].mean(axis=1)])
    return np.hstack([elems, mids]), coords, np.column_stack([boundary, bmid])


# --------------------------------------------------------------------------------------
# Dirichlet elimination + Jacobi-PCG (ref: fem/system.py:26-65; scipy cg as called at
# fem/backends.py:202, semantics in SURVEY.md A.6)
# --------------------------------------------------------------------------------------
class CGFailure(RuntimeError):
    pass


def jacobi_pcg(A, b, rtol=1e-10, maxiter=None, x0=None):
    """scipy.sparse.linalg.cg's iteration with M = diag(A)^-1; returns (x, iterations)."""
    A = csr_array(A)
    b = np.asarray(b, dtype=float)
    n = len(b)
    maxiter = 10 * n if maxiter is None else maxiter
    bnorm = np.linalg.norm(b)
    if bnorm == 0:
        return np.zeros(n), 0
    atol = rtol * bnorm
    dinv = 1.0 / A.diagonal()
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=float)
    r = b - A @ x if x0 is not None else b.copy()
    p = np.zeros(n)
    rho_prev = 1.0
    for it in range(maxiter):
        if np.linalg.norm(r) < atol:
            return x, it
        z = dinv * r
        rho = r @ z
        p = z + (rho / rho_prev) * p if it > 0 else z.copy()
        q = A @ p
        alpha = rho / (p @ q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho
    raise CGFailure(f'CG failed to solve the free-free block: did not converge in {maxiter} iterations.')


def constrained_solve(A, free, fixed, fixed_values, b, solve):
    """x[fixed]=vals; x[free]=solve(A_ff, b_f - A_fx vals)  (ref: fem/system.py:46-52)."""
    A = csr_array(A)
    free, fixed = np.asarray(free, dtype=int), np.asarray(fixed, dtype=int)
    vals = np.asarray(fixed_values, dtype=float)
    x = np.zeros(A.shape[0])
    x[fixed] = vals
    rhs = b[free] - A[np.ix_(free, fixed)] @ vals
    x[free] = solve(A[np.ix_(free, free)], rhs)
    return x


# --------------------------------------------------------------------------------------
# SIMP pieces (ref: fem/topology.py:39-77,247,261-317; fem/sensitivity.py:41-49,305-348;
# fem/design.py:39-78; fem/forms.py:335-374; fem/invariants.py:48-62)
# --------------------------------------------------------------------------------------
def cone_filter(centers, r):
    """Row-normalised cone weights max(0, r-dist) over element centroids, CSR."""
    centers = np.asarray(centers, dtype=float)
    n = len(centers)
    # uniform-grid binning instead of a KD-tree; pairs strictly inside the closed ball
    rows, cols, wts = [np.arange(n)], [np.arange(n)], [np.full(n, float(r))]
    if r > 0:
        lo = centers.min(0)
        cell = np.floor((centers - lo) / r).astype(np.int64)
        dims = cell.max(0) + 1
        flat = np.ravel_multi_index(cell.T, dims)
        order = np.argsort(flat, kind='stable')
        sflat = flat[order]
        starts = np.searchsorted(sflat, np.arange(np.prod(dims) + 1))
        d = centers.shape[1]
        for off in np.ndindex(*([3] * d)):
            nb = cell + (np.array(off) - 1)
            ok = np.all((nb >= 0) & (nb < dims), axis=1)
            idx = np.flatnonzero(ok)
            nflat = np.ravel_multi_index(nb[ok].T, dims)
            cnt = starts[nflat + 1] - starts[nflat]
            src = np.repeat(idx, cnt)
            within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            dst = order[np.repeat(starts[nflat], cnt) + within]
            dist = np.linalg.norm(centers[src] - centers[dst], axis=1)
            keep = (dist <= r) & (src != dst)
            rows.append(src[keep]); cols.append(dst[keep]); wts.append(r - dist[keep])
    rows, cols, wts = np.concatenate(rows), np.concatenate(cols), np.concatenate(wts)
    rowsum = np.bincount(rows, weights=wts, minlength=n)
    return csr_array((wts / (rowsum[rows] + 1e-16), (rows, cols)), shape=(n, n))


def element_vectors(u, elem_nodes, c):
    return np.asarray(u, float).reshape(-1, c)[elem_nodes].reshape(len(elem_nodes), -1)


def element_quadratic(K0, u, elem_nodes, c, lam=None):
    ue = element_vectors(u, elem_nodes, c)
    le = ue if lam is None else element_vectors(lam, elem_nodes, c)
    return np.einsum('ei,eij,ej->e', le, K0, ue)


def oc_update(rho, sens, volumes, volume_frac, move=0.1, max_iters=100, tol=1e-8):
    if np.any(sens < 0.0):
        raise ValueError('the optimality-criteria update needs a nonnegative sensitivity')
    lo, hi = 0.0, 1e15
    new = rho
    vsum = volumes.sum()
    for _ in range(max_iters):
        m = 0.5 * (lo + hi)
        new = np.clip(np.clip(rho * np.sqrt(sens / m), rho - move, rho + move), 1e-6, 1.0)
        if float((volumes * new).sum() / vsum) < volume_frac:
            hi = m
        else:
            lo = m
        if hi - lo <= tol * hi:
            break
    return new


def elastic_fields(geo, u, elem_nodes, E, nu):
    """(strain, stress, compliance) per element at the first quadrature point, 3x3 tensors
    with the plane-strain lift in 2D (ref: fem/forms.py:335-374)."""
    dim = _REF_DIM[geo.kind]
    n_el = len(elem_nodes)
    ue = element_vectors(u, elem_nodes, dim)
    B = voigt_B(geo.grad[:, 0])
    mu, lam = lame(np.broadcast_to(np.asarray(E, float), (n_el,)), nu)
    ev = np.einsum('esk,ek->es', B, ue)
    D = mu[:, None, None] * hooke(dim, 1.0, 0.0) + lam[:, None, None] * hooke(dim, 0.0, 1.0)
    sv = np.einsum('est,et->es', D, ev)
    eps, sig = np.zeros((n_el, 3, 3)), np.zeros((n_el, 3, 3))
    shear = [(0, 1)] if dim == 2 else [(0, 1), (1, 2), (0, 2)]
    for i in range(dim):
        eps[:, i, i], sig[:, i, i] = ev[:, i], sv[:, i]
    for o, (i, j) in enumerate(shear):
        eps[:, i, j] = eps[:, j, i] = ev[:, dim + o] / 2.0
        sig[:, i, j] = sig[:, j, i] = sv[:, dim + o]
    if dim == 2:
        sig[:, 2, 2] = lam * (eps[:, 0, 0] + eps[:, 1, 1])
    comp = np.einsum('eij,eij,e->e', sig, eps, geo.volumes)
    return eps, sig, comp


def stress_field(geo, u, elem_nodes, E, nu):
    """(n_el, n_qp, d, d) in-plane Cauchy stress at every quadrature point of `geo`'s rule, no plane-strain
    lift (ref: fem/forms.py:376-398)."""
    dim = _REF_DIM[geo.kind]
    n_el = len(elem_nodes)
    ue = element_vectors(u, elem_nodes, dim)
    mu, lam = lame(np.broadcast_to(np.asarray(E, float), (n_el,)), nu)
    D = mu[:, None, None] * hooke(dim, 1.0, 0.0) + lam[:, None, None] * hooke(dim, 0.0, 1.0)
    n_qp = geo.grad.shape[1]
    out = np.zeros((n_el, n_qp, dim, dim))
    shear = [(0, 1)] if dim == 2 else [(0, 1), (1, 2), (0, 2)]
    for q in range(n_qp):
        sv = np.einsum('est,et->es', D, np.einsum('esk,ek->es', voigt_B(geo.grad[:, q]), ue))
        for i in range(dim):
            out[:, q, i, i] = sv[:, i]
        for o, (i, j) in enumerate(shear):
            out[:, q, i, j] = out[:, q, j, i] = sv[:, dim + o]
    return out


def von_mises(sig):
    dev = sig - (np.trace(sig, axis1=1, axis2=2) / 3.0)[:, None, None] * np.eye(3)
    return np.sqrt(1.5 * np.einsum('eij,eij->e', dev, dev))


def simp_iteration(elem_nodes, geo, K0, rho, penalty, E0, nu, free, fixed, load, filt, volume_frac,
                   move=0.1, solve=None, chain=lambda C: 1.0):
    """One TopologyOptimizer iteration (ref: fem/topology.py:291-317). Returns
    (u, compliance_e, sensitivity_e, rho_next)."""
    from scipy.sparse.linalg import splu
    from scipy.sparse import csc_array
    dim = _REF_DIM[geo.kind]
    n_nodes_total = len(load) // dim
    A = assemble(elem_nodes, dim, n_nodes_total, (rho ** penalty)[:, None, None] * K0)
    solve = solve or (lambda M, rhs: splu(csc_array(M)).solve(rhs))
    u = constrained_solve(A, free, fixed, np.zeros(len(fixed)), load, solve)
    q = element_quadratic(K0, u, elem_nodes, dim)
    comp = elastic_fields(geo, u, elem_nodes, rho ** penalty * E0, nu)[2]
    sens = penalty * rho ** (penalty - 1) * q * chain(comp.sum())
    rho_next = oc_update(rho, filt @ sens, geo.volumes, volume_frac, move=move)
    return u, comp, sens, rho_next


# ---------------------------------------------------------------------------------------------
# time integrators (ref fem/integrators.py:33-120)
# ---------------------------------------------------------------------------------------------
def theta_method(M, K, b, free, fixed, fixed_values, u0, dt, steps, theta, solve):
    """(M + theta dt K) u_{n+1} = (M - (1 - theta) dt K) u_n + dt b with the Dirichlet dofs eliminated
    (ref fem/integrators.py:48-64).  `solve(A_ff, rhs)` solves the free-free block.  Returns (t, [u_n])."""
    A = (M + theta * dt * K).tocsr()
    rhs_op = (M - (1.0 - theta) * dt * K).tocsr()
    u = np.asarray(u0, dtype=float).copy()
    ts, us = [0.0], [u.copy()]
    for i in range(steps):
        u = constrained_solve(A, free, fixed, fixed_values, rhs_op @ u + dt * b, solve)
        ts.append(dt * (i + 1))
        us.append(u.copy())
    return np.asarray(ts), us


def newmark_method(M, K, b, free, fixed, fixed_values, u0, v0, dt, steps, beta, gamma, solve):
    """Average-acceleration Newmark for M u'' + K u = b (ref fem/integrators.py:86-120): accelerations are
    solved against M (initial) and M + beta dt^2 K with the fixed dofs pinned to zero.  Returns (t, [u], [v])."""
    u, v = np.asarray(u0, dtype=float).copy(), np.asarray(v0, dtype=float).copy()
    if not np.allclose(u[fixed], fixed_values):
        raise ValueError('u0 disagrees with the Dirichlet values at fixed nodes')
    if not np.allclose(v[fixed], 0):
        raise ValueError('v0 must be zero at Dirichlet-fixed nodes')
    zeros = np.zeros(len(fixed))
    a = constrained_solve(M.tocsr(), free, fixed, zeros, b - K @ u, solve)
    Aeff = (M + beta * dt ** 2 * K).tocsr()
    ts, us, vs = [0.0], [u.copy()], [v.copy()]
    for i in range(steps):
        u_pred = u + dt * v + dt ** 2 / 2 * (1 - 2 * beta) * a
        v_pred = v + dt * (1 - gamma) * a
        a = constrained_solve(Aeff, free, fixed, zeros, b - K @ u_pred, solve)
        u = u_pred + beta * dt ** 2 * a
        v = v_pred + gamma * dt * a
        ts.append(dt * (i + 1))
        us.append(u.copy())
        vs.append(v.copy())
    return np.asarray(ts), us, vs


def wave_energy(M, K, u, v):
    """(u^T K u + v^T M v) / 2 (ref fem/integrators.py:123-133)."""
    return float(0.5 * (u @ (K @ u) + v @ (M @ v)))


# ---------------------------------------------------------------------------------------------
# adjoint sensitivity and the design loop (ref fem/sensitivity.py:288-416, fem/design.py:135-192)
# ---------------------------------------------------------------------------------------------
def adjoint_gradient(K0, elem_nodes, c, A, free, fixed, fixed_values, load, dJ_du, factor, self_adjoint, solve):
    """dJ/dp_e = -factor_e lam_e^T K0_e u_e with K^T lam = dJ/du under homogeneous supports; lam = u when the
    quantity is self-adjoint and the supports are homogeneous (ref fem/sensitivity.py:401-416).
    Returns (u, lam, gradient)."""
    u = constrained_solve(A, free, fixed, fixed_values, load, solve)
    if self_adjoint and np.allclose(fixed_values, 0.0):
        lam = u
    else:
        lam = constrained_solve(A, free, fixed, np.zeros(len(fixed)), dJ_du, solve)
    return u, lam, -factor * element_quadratic(K0, u, elem_nodes, c, lam)


def design_step(elem_nodes, c, n_nodes_total, K0, rho, penalty, free, fixed, fixed_values, load, dJ_du_of, self_adjoint,
                filt, volumes, volume_frac, move, solve):
    """One DesignOptimizer.step (ref fem/design.py:160-178): assemble rho^p K0, solve, adjoint gradient of the
    density parameterization, optional filter, OC update.  Returns (u, gradient, rho_next)."""
    A = assemble(elem_nodes, c, n_nodes_total, (rho ** penalty)[:, None, None] * K0)
    u = constrained_solve(A, free, fixed, fixed_values, load, solve)
    lam = u if (self_adjoint and np.allclose(fixed_values, 0.0)) else \
        constrained_solve(A, free, fixed, np.zeros(len(fixed)), dJ_du_of(u), solve)
    grad = -penalty * rho ** (penalty - 1.0) * element_quadratic(K0, u, elem_nodes, c, lam)
    sens = -grad
    if filt is not None:
        sens = filt @ sens
    return u, grad, oc_update(rho, sens, volumes, volume_frac, move=move)


# --------------------------------------------------------------------------------------
# quadrature-tabulated forms and nodal recovery
# --------------------------------------------------------------------------------------
def node_scatter(elem_nodes, values, n_nodes_total):
    """Sum (n_el, N, m) per-(element, local node) values into (n_nodes, m) with np.bincount, the
    reference's scatter (ref: fem/space.py:148-184 _VectorScatterPlan, :590-591 np.add.at)."""
    values = np.asarray(values, dtype=float)
    flat = np.asarray(elem_nodes).ravel()
    vals = values.reshape(len(flat), -1)
    return np.column_stack([np.bincount(flat, weights=vals[:, j], minlength=n_nodes_total)
                            for j in range(vals.shape[1])])


def element_loads(geo, f_qp):
    """b[e,n,c] = sum_q w_eq phi_n(q) f[e,q,c] (ref: fem/forms.py:310-315)."""
    return np.einsum('eq,qn,eqc->enc', geo.wdet, geo.shape, np.asarray(f_qp, dtype=float))


def assemble_load(elem_nodes, n_nodes_total, geo, f_qp):
    """LinearForm load vector, DOFs interleaved per node (ref: fem/space.py:698-708)."""
    return node_scatter(elem_nodes, element_loads(geo, f_qp), n_nodes_total).ravel()


def ke_diffusion(geo, kappa_qp):
    """ref: fem/forms.py:289-293."""
    return np.einsum('eqid,eqjd,eq->eij', geo.grad, geo.grad, geo.wdet * np.asarray(kappa_qp, dtype=float))


def ke_geometric(geo, prestress):
    """sum_q w_q G^T (I (x) sigma0) G: the scalar g_a^T sigma0 g_b on each component's diagonal
    (ref: fem/forms.py:74-94, 425-441)."""
    d = geo.grad.shape[-1]
    s = np.einsum('eqai,eij,eqbj,eq->eab', geo.grad, np.asarray(prestress, dtype=float), geo.grad, geo.wdet)
    n_el, N, _ = s.shape
    K = np.zeros((n_el, N * d, N * d))
    for c in range(d):
        K[:, c::d, c::d] = s
    return K


def recover_nodal_average(elem_nodes, volumes, values, n_nodes_total):
    """Volume-weighted nodal average (ref: fem/space.py:578-596)."""
    values = np.asarray(values, dtype=float)
    n_el, N = np.asarray(elem_nodes).shape
    trailing = values.shape[1:]
    flat = values.reshape(n_el, -1) * volumes[:, None]
    sums = node_scatter(elem_nodes, np.repeat(flat[:, None, :], N, axis=1), n_nodes_total)
    norms = node_scatter(elem_nodes, np.repeat(volumes[:, None, None], N, axis=1), n_nodes_total)
    norms = np.where(norms > 0, norms, 1.0)
    return (sums / norms).reshape((n_nodes_total,) + trailing)


def recover_nodal_l2(elem_nodes, geo, values, n_nodes_total, mass, solve):
    """Global L2 projection of an element-constant field: M q = sum_e f_e int_e phi (ref: fem/space.py:598-616)."""
    values = np.asarray(values, dtype=float)
    n_el = len(values)
    trailing = values.shape[1:]
    shape_integral = np.einsum('eq,qn->en', geo.wdet, geo.shape)
    contrib = shape_integral[:, :, None] * values.reshape(n_el, 1, -1)
    load = node_scatter(elem_nodes, contrib, n_nodes_total)
    q = np.column_stack([solve(mass, load[:, j]) for j in range(load.shape[1])])
    return q.reshape((n_nodes_total,) + trailing)


# --------------------------------------------------------------------------------------
# stress quantities of interest, plane strain (ref: fem/sensitivity.py:108-276)
# --------------------------------------------------------------------------------------
def von_mises_gradient(geo, u, elem_nodes, E, nu):
    """(vm (n_el,), dvm/du_e (n_el, N*2)) from the in-plane Voigt stress at the first quadrature point,
    sigma_zz = nu (s_xx + s_yy) (ref: fem/sensitivity.py:151-189)."""
    n_el = len(elem_nodes)
    B = voigt_B(geo.grad[:, 0])
    mu, lam = lame(np.broadcast_to(np.asarray(E, float), (n_el,)), nu)
    D = mu[:, None, None] * hooke(2, 1.0, 0.0) + lam[:, None, None] * hooke(2, 0.0, 1.0)
    DB = np.einsum('est,etk->esk', D, B)
    s = np.einsum('esk,ek->es', DB, element_vectors(u, elem_nodes, 2))
    sxx, syy, sxy = s[:, 0], s[:, 1], s[:, 2]
    szz = nu * (sxx + syy)
    Q = sxx ** 2 + syy ** 2 + szz ** 2 - sxx * syy - syy * szz - szz * sxx + 3.0 * sxy ** 2
    vm = np.sqrt(np.maximum(Q, 0.0))
    dQ = np.empty_like(s)
    dQ[:, 0] = 2 * sxx - syy - szz + nu * (2 * szz - syy - sxx)
    dQ[:, 1] = 2 * syy - sxx - szz + nu * (2 * szz - sxx - syy)
    dQ[:, 2] = 6.0 * sxy
    dvm_ds = np.zeros_like(s)
    safe = vm > 0
    dvm_ds[safe] = dQ[safe] / (2.0 * vm[safe, None])
    return vm, np.einsum('esk,es->ek', DB, dvm_ds)


def stress_qoi(geo, u, elem_nodes, n_nodes_total, E, nu, region=None, p=None):
    """(value, dJ/du) of MeanStress (p None) or SoftMaxStress (p-norm) with volume weights over `region`
    (ref: fem/sensitivity.py:191-276)."""
    vm, dvm_du = von_mises_gradient(geo, u, elem_nodes, E, nu)
    vol = geo.volumes
    w = np.where(np.ones(len(vol), bool) if region is None else np.asarray(region), vol, 0.0)
    w = w / w.sum()
    N = np.asarray(elem_nodes).shape[1]
    if p is None:
        value, per_element = float(w @ vm), w[:, None] * dvm_du
    else:
        total = float(w @ vm ** p)
        value = total ** (1.0 / p)
        per_element = (total ** (1.0 / p - 1.0) * w * vm ** (p - 1.0))[:, None] * dvm_du
    return value, node_scatter(elem_nodes, per_element.reshape(len(vol), N, 2), n_nodes_total).ravel()
