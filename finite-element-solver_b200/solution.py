"""Typed solutions (ref fem/solution.py:28-147, 265-275; fem/invariants.py:31-79): what a solve returns and what
`fes_b200.io` writes.  Same class names, field names and field order as the reference, so an archive
written by either library loads in the other.  Arrays are host numpy (a solution is an output format);
`ElasticSolution.from_solve` recovers the stress state with the device kernel `fes_elastic_fields`.
"""
from dataclasses import dataclass, field
from functools import cached_property

import numpy as np


def _host(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)


@dataclass(frozen=True, eq=False)
class Solution:
    mesh: object
    n_components: int
    element_type: object = field(default=None, kw_only=True)

    @cached_property
    def space(self):
        """The FunctionSpace the DOFs live on, rebuilt from (mesh, element_type, n_components)."""
        from .space import FunctionSpace
        return FunctionSpace(self.mesh, self.element_type, n_components=self.n_components)

    def save(self, path):
        from .io import save_solution
        save_solution(self, path)

    @staticmethod
    def load(path):
        from .io import load_solution
        return load_solution(path)


def _moved(mesh, displacement, n_components, scale=1.0):
    """The mesh displaced by the leading vertex DOFs of a (P1 or P2) displacement vector."""
    from .mesh import Mesh
    n_vertices = len(mesh.vertices)
    d = np.asarray(displacement, dtype=float).reshape(-1, n_components)[:n_vertices]
    return Mesh(mesh.vertices + scale * d, mesh.elements, mesh.boundary)


@dataclass(frozen=True, eq=False)
class FieldSolution(Solution):
    """A single steady field u (ref fem/solution.py:65-80)."""
    u: object = None

    def deformed_mesh(self):
        return _moved(self.mesh, self.u, self.n_components)


@dataclass(frozen=True, eq=False)
class ScalarFieldSolution(FieldSolution):
    """A scalar field plus its per-element flux grad u (ref fem/solution.py:83-104)."""
    flux: object = None        # (n_elements, spatial_dim)

    @classmethod
    def from_solve(cls, space, u):
        """Package a scalar solve with its per-element flux grad u (device kernel `fes_scalar_gradient`)."""
        return cls(space.mesh, space.n_components, _host(u), flux=_host(space.gradient(u)), element_type=space.element_type)

    def nodal_flux(self, method='average'):
        return self.space.recover_nodal(self.flux, method=method)


def trace(t):
    return np.trace(t, axis1=-2, axis2=-1)


def deviatoric(t):
    d = t.shape[-1]
    return t - (trace(t) / d)[:, None, None] * np.eye(d)


def pressure(stress):
    return -trace(stress) / stress.shape[-1]


def von_mises(stress):
    s = deviatoric(np.asarray(stress, dtype=float))
    return np.sqrt(1.5 * np.einsum('eij,eij->e', s, s))


def principal(t):
    return np.asarray(np.linalg.eigvalsh(t), dtype=np.float64)


def max_shear(t):
    v = principal(t)
    return (v[:, -1] - v[:, 0]) / 2.0


@dataclass(frozen=True, eq=False)
class ElasticSolution(FieldSolution):
    """A displacement field plus the stress state recovered from it (ref fem/solution.py:107-186)."""
    strain: object = None       # (n_elements, 3, 3)
    stress: object = None       # (n_elements, 3, 3)
    compliance: object = None   # (n_elements,)

    def __post_init__(self):
        for name in ('strain', 'stress'):
            value = getattr(self, name)
            if np.ndim(value) != 3:
                raise ValueError(f'{type(self).__name__}.{name} must be an (n_elements, 3, 3) tensor field, '
                                 f'got shape {np.shape(value)}')

    @classmethod
    def from_solve(cls, space, u, form):
        fields = form.derived_fields(space, u)
        et = space.element_type if space.element_type.SHAPE_DEGREE > 1 else None
        return cls(space.mesh, space.n_components, _host(u), _host(fields.strain), _host(fields.stress),
                   _host(fields.compliance), element_type=et)

    @property
    def von_mises(self):
        return von_mises(self.stress)

    def nodal_stress(self, method='average'):
        return self.space.recover_nodal(self.stress, method=method)

    def nodal_strain(self, method='average'):
        return self.space.recover_nodal(self.strain, method=method)

    def nodal_von_mises(self, method='average'):
        return von_mises(self.nodal_stress(method=method))

    @property
    def pressure(self):
        return pressure(self.stress)

    @property
    def principal_stress(self):
        return principal(self.stress)

    @property
    def max_shear(self):
        return max_shear(self.stress)


@dataclass(frozen=True, eq=False)
class BucklingSolution(Solution):
    """Linearised-buckling eigenpairs (ref fem/solution.py:189-222); carried for file interoperability --
    the eigen solve itself is outside this library's path."""
    load_factors: object = None   # (n_modes,) ascending
    modes: object = None          # (n_modes, n_dofs)

    @property
    def critical_load_factor(self):
        return float(self.load_factors[0])

    def mode_mesh(self, i, scale=1.0):
        return _moved(self.mesh, self.modes[i], self.n_components, scale)


@dataclass(frozen=True, eq=False)
class ModalSolution(Solution):
    """Free-vibration eigenpairs (ref fem/solution.py:225-262); carried for file interoperability."""
    angular_frequencies: object = None   # (n_modes,) ascending, rad/s
    modes: object = None                 # (n_modes, n_dofs)

    @property
    def frequencies(self):
        return np.asarray(self.angular_frequencies) / (2 * np.pi)

    @property
    def periods(self):
        with np.errstate(divide='ignore'):
            return 1.0 / self.frequencies

    def mode_mesh(self, i, scale=1.0):
        return _moved(self.mesh, self.modes[i], self.n_components, scale)


@dataclass(frozen=True, eq=False)
class TransientSolution(Solution):
    """A time series: the times t and the field u at each step (ref fem/solution.py:265-269)."""
    t: object = None
    u: object = None


@dataclass(frozen=True, eq=False)
class WaveSolution(TransientSolution):
    """A time series that also carries du/dt at each step (ref fem/solution.py:272-275)."""
    dudt: object = None
