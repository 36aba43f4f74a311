"""FunctionSpace on the GPU: DOF numbering, the cached CSR pattern/scatter plan and `assemble`.

Mirrors the reference's FunctionSpace surface (ref fem/space.py:310-436,479-482,669-696,757-803):
same constructor, `element_nodes / node_coords / boundary_nodes / n_dofs / dof_indices`,
`assemble(form, boundary=False)`, `mass_matrix`, `boundary_mass_matrix`, `element_volumes`.
`assemble` returns a host scipy `csr_array` whose indptr/indices are bit-identical to the
reference's; `assemble_device` returns the same operator as a `DeviceCSR` left in HBM, which is
what the device-resident solve and SIMP paths consume.
"""
import ctypes as C
from functools import cached_property

import numpy as np
import torch
from scipy.sparse import csr_array

from . import _lib
from .csr import DeviceCSR
from .elements import Element, element_type_for
from .forms import MassForm, PrecomputedForm


def dof_indices(element, n_components):
    """Interleaved DOFs of an element's nodes, batched over leading axes (ref fem/space.py:55-66)."""
    element = np.asarray(element)
    return (n_components * element[..., None] + np.arange(n_components)).reshape(*element.shape[:-1], -1)


def p2_connectivity(mesh):
    """(element_nodes, node_coords, boundary_nodes) of the straight-sided P2 triangle space:
    vertices, then one midpoint node per unique edge in lexicographic order; element nodes
    [c0,c1,c2,m12,m02,m01] (ref fem/space.py:205-252), vectorised."""
    nv = len(mesh.vertices)
    el = mesh.elements
    opposite = np.concatenate([el[:, [1, 2]], el[:, [0, 2]], el[:, [0, 1]]])
    opposite = np.sort(opposite, axis=1)
    key = opposite[:, 0] * nv + opposite[:, 1]
    ukey = np.unique(key)
    mids = nv + np.searchsorted(ukey, key).reshape(3, -1).T
    edge_a, edge_b = ukey // nv, ukey % nv
    coords = np.vstack([mesh.vertices, (mesh.vertices[edge_a] + mesh.vertices[edge_b]) / 2.0])
    bs = np.sort(mesh.boundary, axis=1)
    bmid = nv + np.searchsorted(ukey, bs[:, 0] * nv + bs[:, 1])
    return np.hstack([el, mids]), coords, np.column_stack([mesh.boundary, bmid])


class NodeSet:
    """What BoundaryConditions.resolve reads off a space (ref fem/space.py:187-202)."""

    def __init__(self, vertices, boundary, spatial_dim):
        self.vertices, self.boundary, self.spatial_dim = vertices, boundary, spatial_dim
        self.boundary_idxs = np.unique(boundary)


class _DevView:
    """Zero-copy torch view of library-owned device memory."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 2}


class ScatterPlan:
    """Owner of one fes_plan: CSR pattern + contribution lists for a (connectivity, n_comp)."""

    def __init__(self, elem_nodes_d, n_nodes, n_comp):
        L = _lib.lib()
        n_el, N = elem_nodes_d.shape
        handle = C.c_void_p()
        _lib.check(L.fes_plan_create(_lib.ptr(elem_nodes_d), n_el, N, n_nodes, n_comp, _lib.stream(), C.byref(handle)))
        self._h = handle
        self.n_dofs, self.nnz, self.n_pairs = (int(L.fes_plan_n_dofs(handle)), int(L.fes_plan_nnz(handle)),
                                               int(L.fes_plan_n_pairs(handle)))
        self.n_entries = n_el * (N * n_comp) ** 2
        dev = elem_nodes_d.device
        self.indptr_d = torch.as_tensor(_DevView(L.fes_plan_indptr(handle), self.n_dofs + 1, '<i4'), device=dev)
        self.indices_d = (torch.as_tensor(_DevView(L.fes_plan_indices(handle), self.nnz, '<i4'), device=dev)
                          if self.nnz else torch.zeros(0, dtype=torch.int32, device=dev))
        self._host = None

    @property
    def handle(self):
        return self._h

    def host_pattern(self):
        """(indptr, indices) as int64 host arrays, exported once."""
        if self._host is None:
            indptr = np.empty(self.n_dofs + 1, dtype=np.int64)
            indices = np.empty(self.nnz, dtype=np.int64)
            _lib.check(_lib.lib().fes_plan_export_csr(self._h, _lib.ptr(indptr), _lib.ptr(indices)))
            self._host = (indptr, indices)
        return self._host

    @property
    def n_tiles(self):
        """Tiles of the fused kernel; 0 when the mesh needs the untiled fallback kernel."""
        return int(_lib.lib().fes_plan_n_tiles(self._h))

    @property
    def n_run_tiles(self):
        """Tiles of the stencil-run kernel (structured node ranges); 0 when no run was detected."""
        return int(_lib.lib().fes_plan_n_run_tiles(self._h))

    @property
    def run_nodes(self):
        """Nodes the stencil-run tiles cover."""
        return int(_lib.lib().fes_plan_run_nodes(self._h))

    @property
    def bytes(self):
        return int(_lib.lib().fes_plan_bytes(self._h))

    @property
    def read_bytes(self):
        return int(_lib.lib().fes_plan_read_bytes(self._h))

    def __del__(self):
        try:
            if self._h:
                _lib.lib().fes_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


class FunctionSpace:
    def __init__(self, mesh, element_type=None, n_components=1, device=None):
        element_type = element_type if element_type is not None else element_type_for(mesh)
        if not (isinstance(element_type, type) and issubclass(element_type, Element)):
            raise TypeError('element_type must be an fes_b200 element class')
        if element_type.SUB_TYPE is None:
            raise NotImplementedError(f'{element_type.__name__} has no boundary element type, so boundary '
                                      'integrals (and hence a FunctionSpace) are not defined for it yet')
        if n_components < 1:
            raise ValueError(f'n_components must be at least 1, got {n_components}')
        self.mesh, self.element_type, self.boundary_type = mesh, element_type, element_type.SUB_TYPE
        self.n_components = n_components
        self.device = device if device is not None else _lib.require_cuda()
        if element_type.SHAPE_DEGREE == 1:
            self.element_nodes, self.node_coords, self.boundary_nodes = mesh.elements, mesh.vertices, mesh.boundary
        else:
            self.element_nodes, self.node_coords, self.boundary_nodes = p2_connectivity(mesh)
        if self.element_nodes.shape[1] != element_type.N:
            raise ValueError(f'{element_type.__name__} needs {element_type.N} nodes per element')
        self._en_d = torch.as_tensor(self.element_nodes.astype(np.int32)).to(self.device).contiguous()
        self._bn_d = torch.as_tensor(np.ascontiguousarray(self.boundary_nodes).astype(np.int32)
                                     .reshape(-1, self.boundary_type.N)).to(self.device).contiguous()
        self._coords_d = torch.as_tensor(np.ascontiguousarray(self.node_coords, dtype=np.float64)).to(self.device)
        self._plans = {}

    def __repr__(self):
        return f'FunctionSpace({self.element_type.__name__}, n_components={self.n_components}, n_dofs={self.n_dofs})'

    # -- sizing and numbering ----------------------------------------------------------------
    @property
    def spatial_dim(self):
        return self.mesh.spatial_dim

    @property
    def n_nodes(self):
        return len(self.node_coords)

    @property
    def n_dofs(self):
        return self.n_nodes * self.n_components

    @cached_property
    def nodes(self):
        if self.element_type.SHAPE_DEGREE == 1:
            return self.mesh
        return NodeSet(self.node_coords, self.boundary_nodes, self.spatial_dim)

    def dof_indices(self, element):
        return dof_indices(element, self.n_components)

    def _n_elements(self, boundary=False):
        return len(self.boundary_nodes) if boundary else len(self.element_nodes)

    def _conn(self, boundary):
        return (self._bn_d, self.boundary_type) if boundary else (self._en_d, self.element_type)

    def plan(self, boundary=False):
        """The cached pattern/scatter plan (the reference's _volume_scatter / _boundary_scatter)."""
        if boundary not in self._plans:
            conn, _ = self._conn(boundary)
            self._plans[boundary] = ScatterPlan(conn, self.n_nodes, self.n_components)
        return self._plans[boundary]

    # -- geometry ----------------------------------------------------------------------------
    @cached_property
    def element_volumes_d(self):
        out = torch.empty(len(self.element_nodes), dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().fes_element_volumes(self.element_type.KIND, self.spatial_dim, _lib.ptr(self._en_d),
                                                  len(self.element_nodes), _lib.ptr(self._coords_d), _lib.ptr(out),
                                                  _lib.stream()))
        return out

    @cached_property
    def element_volumes(self):
        return self.element_volumes_d.cpu().numpy()

    @property
    def total_volume(self):
        return float(self.element_volumes.sum())

    # -- operators ---------------------------------------------------------------------------
    def _element_matrices(self, form, boundary=False):
        conn, etype = self._conn(boundary)
        low = form.lower(self, boundary)
        n_el, k = conn.shape[0], etype.N * self.n_components
        out = torch.empty((n_el, k, k), dtype=torch.float64, device=self.device)
        cf = low.coeffs()
        _lib.check(_lib.lib().fes_element_matrices(etype.KIND, low.form, self.n_components, self.spatial_dim,
                                                   _lib.ptr(conn), n_el, _lib.ptr(self._coords_d), C.byref(cf),
                                                   _lib.ptr(out), _lib.stream()))
        return out

    def _assemble_values(self, form, boundary=False, out=None):
        plan = self.plan(boundary)
        conn, etype = self._conn(boundary)
        values = out if out is not None else torch.empty(plan.nnz, dtype=torch.float64, device=self.device)
        L = _lib.lib()
        if isinstance(form, PrecomputedForm) or getattr(form, 'precomputed', False):
            Ke = form.element_matrices(self, boundary)
            if not isinstance(Ke, torch.Tensor):
                Ke = torch.as_tensor(np.ascontiguousarray(Ke, dtype=np.float64))
            Ke = Ke.to(self.device).contiguous()
            if Ke.numel() != plan.n_entries:
                raise ValueError(f'expected element matrices covering {plan.n_entries} entries, got '
                                 f'{Ke.numel()} (shape {tuple(Ke.shape)})')
            scale = getattr(form, 'scale', None)
            if scale is not None and not isinstance(scale, torch.Tensor):
                scale = torch.as_tensor(np.ascontiguousarray(scale, dtype=np.float64)).to(self.device)
            _lib.check(L.fes_assemble_precomputed(plan.handle, _lib.ptr(Ke), _lib.ptr(scale), 1.0,
                                                  _lib.ptr(values), _lib.stream()))
            return values
        low = form.lower(self, boundary)
        cf = low.coeffs()
        _lib.check(L.fes_assemble(plan.handle, etype.KIND, low.form, self.spatial_dim, _lib.ptr(conn),
                                  _lib.ptr(self._coords_d), C.byref(cf), _lib.ptr(values), _lib.stream()))
        return values

    def assemble_device(self, form, boundary=False, out=None):
        plan = self.plan(boundary)
        values = self._assemble_values(form, boundary, out)
        A = DeviceCSR((self.n_dofs, self.n_dofs), plan.indptr_d, plan.indices_d, values, plan.host_pattern)
        A._plan = plan   # keeps the library-owned index arrays alive
        return A

    def assemble(self, form, boundary=False):
        """Drop-in for the reference: a host csr_array (int64 indices, fp64 data)."""
        plan = self.plan(boundary)
        values = self._assemble_values(form, boundary)
        indptr, indices = plan.host_pattern()
        return csr_array((values.cpu().numpy(), indices, indptr), shape=(self.n_dofs, self.n_dofs))

    # -- quadrature-tabulated loads and nodal recovery -----------------
    def quadrature(self, min_degree):
        """(points (n_el, n_qp, sd), weight_detJ (n_el, n_qp)) of the volume elements at the cheapest
        tabulated rule of degree >= min_degree, cached per rule (ref fem/space.py:412-425)."""
        cache = self.__dict__.setdefault('_quad_cache', {})
        L = _lib.lib()
        nq = int(L.fes_quadrature_size(self.element_type.KIND, int(min_degree)))
        if nq == 0:
            raise NotImplementedError(f'no quadrature rule of degree >= {min_degree} for '
                                      f'reference_dim={self.element_type.REF_DIM}')
        if nq not in cache:
            n_el = len(self.element_nodes)
            pts = torch.empty((n_el, nq, self.spatial_dim), dtype=torch.float64, device=self.device)
            wdet = torch.empty((n_el, nq), dtype=torch.float64, device=self.device)
            _lib.check(L.fes_quadrature_points(self.element_type.KIND, self.spatial_dim, int(min_degree),
                                               _lib.ptr(self._en_d), n_el, _lib.ptr(self._coords_d), _lib.ptr(pts),
                                               _lib.ptr(wdet), _lib.stream()))
            cache[nq] = (pts, wdet)
        return cache[nq]

    def gradient(self, u):
        """(n_elements, spatial_dim) gradient of a scalar nodal field, one value per element, taken at the first
        quadrature point (ref fem/space.py:511-517).  numpy in -> numpy out, device tensor in -> device tensor out."""
        host = not isinstance(u, torch.Tensor)
        ud = torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64)) if host else u
        ud = ud.to(device=self.device, dtype=torch.float64).contiguous().reshape(-1)
        if ud.numel() != self.n_nodes:
            raise ValueError(f'expected one value per node ({self.n_nodes}), got {ud.numel()}')
        n_el = len(self.element_nodes)
        out = torch.empty((n_el, self.spatial_dim), dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().fes_scalar_gradient(self.element_type.KIND, self.spatial_dim,
                                                  self.element_type.default_quadrature_degree(), _lib.ptr(self._en_d), n_el,
                                                  _lib.ptr(self._coords_d), _lib.ptr(ud), _lib.ptr(out), _lib.stream()))
        return out.cpu().numpy() if host else out

    def node_scatter(self, element_values, m):
        """Sum (n_el, N, m) per-(element, local node) values into (n_nodes, m), each node in ascending
        element id: bit-identical to the reference's bincount / np.add.at scatters
        (ref fem/space.py:148-184, 578-596)."""
        vals = element_values.to(device=self.device, dtype=torch.float64).contiguous()
        n_el, N = self._en_d.shape
        if vals.numel() != n_el * N * m:
            raise ValueError(f'expected element vectors covering {n_el * N * m} entries, got {vals.numel()} '
                             f'(shape {tuple(vals.shape)})')
        out = torch.empty((self.n_nodes, m), dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().fes_node_scatter(self.plan().handle, _lib.ptr(vals), int(m), _lib.ptr(out), _lib.stream()))
        return out

    def assemble_load_d(self, form):
        if form.n_components != self.n_components:
            raise ValueError(f'LinearForm({form.n_components}) on a space with {self.n_components} components')
        return self.node_scatter(form.element_vectors(self), self.n_components).reshape(-1)

    def assemble_load(self, form):
        """Scatter a LinearForm's element vectors into the global load vector (ref fem/space.py:698-708)."""
        return self.assemble_load_d(form).cpu().numpy()

    def recover_nodal(self, values, method='average'):
        """A continuous nodal field from a per-element one (ref fem/space.py:545-616): 'average' is the
        volume-weighted nodal average (summed in the reference's order), 'l2' the global L2 projection
        (mass solve by Jacobi-PCG to rtol 1e-13 instead of a sparse LU)."""
        host = not isinstance(values, torch.Tensor)
        v = torch.as_tensor(np.asarray(values, dtype=np.float64)) if host else values
        v = v.to(device=self.device, dtype=torch.float64)
        n_el, N = self._en_d.shape
        if v.shape[0] != n_el:
            raise ValueError(f'expected one value per element ({n_el}), got {v.shape[0]}')
        trailing = tuple(v.shape[1:])
        flat = v.reshape(n_el, -1)
        m = flat.shape[1]
        if method == 'average':
            w = self.element_volumes_d
            weighted = (flat * w[:, None])[:, None, :].expand(n_el, N, m)
            sums = self.node_scatter(weighted, m)
            norms = self.node_scatter(w[:, None, None].expand(n_el, N, 1), 1)
            norms = torch.where(norms > 0, norms, torch.ones_like(norms))
            out = (sums / norms).reshape((self.n_nodes,) + trailing)
        elif method == 'l2':
            ones = self._element_loads(torch.ones((n_el, self._default_nq(), 1), dtype=torch.float64, device=self.device),
                                       self.element_type.default_quadrature_degree())   # (n_el, N, 1): int_e phi_n
            contrib = ones * flat[:, None, :]                      # (n_el, N, m): f_e * int_e phi_n
            load = self.node_scatter(contrib, m)
            out = self._nodal_mass_solve(load).reshape((self.n_nodes,) + trailing)
        else:
            raise ValueError(f"unknown recovery method {method!r}; use 'average' or 'l2'")
        return out.cpu().numpy() if host else out

    def _default_nq(self):
        return int(_lib.lib().fes_quadrature_size(self.element_type.KIND, self.element_type.default_quadrature_degree()))

    def _element_loads(self, f_qp, degree):
        """(n_el, N, m) = sum_q w_eq phi_n(q) f[e, q, :] (ref fem/forms.py:310-315)."""
        n_el, N = self._en_d.shape
        m = f_qp.shape[2]
        out = torch.empty((n_el, N, m), dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().fes_element_loads(self.element_type.KIND, self.spatial_dim, int(degree), _lib.ptr(self._en_d),
                                                n_el, _lib.ptr(self._coords_d), _lib.ptr(f_qp.contiguous()), m,
                                                _lib.ptr(out), _lib.stream()))
        return out

    def project_to_nodal(self, values_qp, degree=None):
        """L2-project a per-quadrature-point field (n_el, n_qp, *shape) sampled at the rule of `degree`
        (default: the space's own) onto the nodal space (ref fem/space.py:618-640)."""
        host = not isinstance(values_qp, torch.Tensor)
        v = torch.as_tensor(np.asarray(values_qp, dtype=np.float64)) if host else values_qp
        v = v.to(device=self.device, dtype=torch.float64)
        degree = self.element_type.default_quadrature_degree() if degree is None else degree
        trailing = tuple(v.shape[2:])
        flat = v.reshape(v.shape[0], v.shape[1], -1)
        load = self.node_scatter(self._element_loads(flat, degree), flat.shape[2])
        out = self._nodal_mass_solve(load).reshape((self.n_nodes,) + trailing)
        return out.cpu().numpy() if host else out

    def _nodal_mass_solve(self, load):
        """M q = load for every column, M the scalar consistent mass matrix (ref fem/space.py:655-667)."""
        from .backends import JacobiPCGBackend
        if '_nodal_mass_solver' not in self.__dict__:
            scalar = self if self.n_components == 1 else FunctionSpace(self.mesh, self.element_type, 1, self.device)
            self.__dict__['_nodal_mass_solver'] = JacobiPCGBackend(rtol=1e-13).prepare(scalar.mass_matrix_d)
        solver = self.__dict__['_nodal_mass_solver']
        cols = []
        for j in range(load.shape[1]):
            x = torch.zeros(self.n_nodes, dtype=torch.float64, device=self.device)
            solver.solve_device(load[:, j].contiguous(), x, warm_start=False)
            cols.append(x)
        return torch.stack(cols, dim=1)

    @cached_property
    def mass_matrix_d(self):
        return self.assemble_device(MassForm(self.n_components))

    @cached_property
    def mass_matrix(self):
        return self.mass_matrix_d.to_scipy()

    @cached_property
    def boundary_mass_matrix(self):
        return self.assemble(MassForm(self.n_components), boundary=True)
