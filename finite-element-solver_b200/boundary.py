"""Boundary-condition specification and its resolution to DOF arrays (ref fem/boundary.py:
131-343), with the per-vertex Python loops replaced by array operations so resolving a
4-20 M node problem costs milliseconds.  Produces the same `free_idxs` (ascending),
`fixed_idxs` (vertex-major, component-minor, free components skipped), `fixed_values`,
per-condition facet masks and nodal traction fields as the reference.
"""
from dataclasses import dataclass
from enum import Enum

import numpy as np

from .regions import coerce_components, evaluate_field, is_mesh_bound


class BCType(Enum):
    DIRICHLET = 'dirichlet'
    NEUMANN = 'neumann'
    ROBIN = 'robin'


@dataclass(frozen=True)
class NeumannContribution:
    facet_mask: np.ndarray
    traction: np.ndarray      # (n_nodes, c), nonzero on the region's nodes


@dataclass(frozen=True)
class RobinContribution:
    facet_mask: np.ndarray
    kappa: float
    g: np.ndarray


@dataclass(frozen=True)
class ResolvedBC:
    n_vertices: int
    n_components: int
    fixed_idxs: np.ndarray
    free_idxs: np.ndarray
    fixed_values: np.ndarray
    neumann_load: np.ndarray
    dirichlet_vertices: np.ndarray
    neumann_vertices: np.ndarray
    robin: tuple = ()
    neumann: tuple = ()


class BoundaryConditions:
    def __init__(self):
        self.conditions = []
        self.robin_conditions = []

    @staticmethod
    def _check_region(region):
        if not callable(region):
            raise TypeError('region must be a callable over point coordinates; pass a helper from '
                            f'fes_b200.regions. Got {type(region).__name__}.')

    def add(self, bc_type, region, value):
        bc_type = BCType(bc_type)
        if bc_type is BCType.ROBIN:
            raise ValueError('use add_robin(region, kappa, g) for Robin conditions')
        self._check_region(region)
        self.conditions.append((bc_type, region, value))

    def add_robin(self, region, kappa, g):
        self._check_region(region)
        self.robin_conditions.append((region, float(kappa), g))

    def select(self, nodes, region):
        selected = np.flatnonzero(region(nodes.vertices))
        boundary = np.asarray(nodes.boundary_idxs, dtype=int)
        if is_mesh_bound(region):
            interior = np.setdiff1d(selected, boundary)
            if len(interior):
                raise ValueError(f'boundary conditions on non-boundary nodes: {sorted(interior)}')
            return selected
        return np.intersect1d(selected, boundary)

    @staticmethod
    def _facet_mask(nodes, idxs):
        member = np.zeros(len(nodes.vertices), dtype=bool)
        member[idxs] = True
        return member[np.asarray(nodes.boundary)].all(axis=1)

    def resolve(self, nodes, n_components):
        n, c = len(nodes.vertices), n_components
        pinned = np.full((n, c), np.nan)          # Dirichlet value per (node, component); NaN = free
        touched = np.zeros(n, dtype=bool)
        neumann = np.zeros((n, c))
        neumann_parts, robin_parts = [], []
        d_nodes, n_nodes = [], []
        for bc_type, region, value in self.conditions:
            idxs = self.select(nodes, region)
            if bc_type is BCType.DIRICHLET:
                vals = coerce_components(value, nodes.vertices[idxs], c)
                if vals.shape != (len(idxs), c):
                    raise ValueError(f'field must give {c} component(s) per point, got shape {vals.shape} '
                                     f'for {len(idxs)} point(s)')
                old = pinned[idxs]
                both = ~np.isnan(old) & ~np.isnan(vals)
                clash = both & ~np.isclose(np.where(both, old, 0.0), np.where(both, vals, 0.0))
                if clash.any():
                    v = int(idxs[np.flatnonzero(clash.any(axis=1))[0]])
                    raise ValueError(f'conflicting Dirichlet values at vertex {v}: {pinned[v]} and '
                                     f'{vals[np.flatnonzero(idxs == v)[0]]}')
                pinned[idxs] = np.where(np.isnan(vals), old, vals)
                touched[idxs] = True
                d_nodes.append(idxs)
            else:
                vals = evaluate_field(value, nodes.vertices[idxs], c)
                neumann[idxs] += vals
                traction = np.zeros((n, c))
                traction[idxs] = vals
                neumann_parts.append(NeumannContribution(self._facet_mask(nodes, idxs), traction))
                n_nodes.append(idxs)
        for region, kappa, g_field in self.robin_conditions:
            idxs = self.select(nodes, region)
            g = np.zeros((n, c))
            g[idxs] = evaluate_field(g_field, nodes.vertices[idxs], c)
            robin_parts.append(RobinContribution(self._facet_mask(nodes, idxs), kappa, g))

        is_fixed = ~np.isnan(pinned)
        conflict = is_fixed & (neumann != 0.0)
        if conflict.any():
            raise ValueError('vertices carry a Dirichlet and a Neumann condition on the same '
                             f'component: {sorted(int(v) for v in np.flatnonzero(conflict.any(axis=1)))}')
        flat_fixed = is_fixed.ravel()
        fixed_idxs = np.flatnonzero(flat_fixed)           # vertex-major, component-minor, ascending
        cat = (lambda parts: np.unique(np.concatenate(parts)).astype(int) if parts else np.zeros(0, dtype=int))
        return ResolvedBC(
            n_vertices=n, n_components=c, fixed_idxs=fixed_idxs, free_idxs=np.flatnonzero(~flat_fixed),
            fixed_values=pinned.ravel()[fixed_idxs], neumann_load=neumann,
            dirichlet_vertices=cat(d_nodes), neumann_vertices=cat(n_nodes),
            robin=tuple(robin_parts), neumann=tuple(neumann_parts))
