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
        if isinstance(form, PrecomputedForm):
            Ke = form.element_matrices(self, boundary)
            if not isinstance(Ke, torch.Tensor):
                Ke = torch.as_tensor(np.ascontiguousarray(Ke, dtype=np.float64))
            Ke = Ke.to(self.device).contiguous()
            if Ke.numel() != plan.n_entries:
                raise ValueError(f'expected element matrices covering {plan.n_entries} entries, got '
                                 f'{Ke.numel()} (shape {tuple(Ke.shape)})')
            scale = form.scale
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

    @cached_property
    def mass_matrix_d(self):
        return self.assemble_device(MassForm(self.n_components))

    @cached_property
    def mass_matrix(self):
        return self.mass_matrix_d.to_scipy()

    @cached_property
    def boundary_mass_matrix(self):
        return self.assemble(MassForm(self.n_components), boundary=True)
