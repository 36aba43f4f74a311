"""Adjoint sensitivity on the device (ref fem/sensitivity.py): the same three parts -- a
`QuantityOfInterest` (adjoint load), a `Parameterization` (dK/dp turned into a gradient) and
`SensitivityAnalysis` (forward + adjoint solves on one prepared `DiscreteSystem`).

The element bilinear form lam_e^T K0_e u_e (ref fem/sensitivity.py:305-309) is never formed from
(n_el, k, k) matrices: the strain-energy kernel `fes_elastic_energy` gives Q(w) = w_e^T K0_e w_e
per element, and K0_e being symmetric, lam^T K0 u = (Q(u + lam) - Q(u - lam)) / 4; the self-adjoint
case (lam = u) is one call.  Vectors may be numpy arrays or device tensors; results come back in kind.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .system import DiscreteSystem


def _dev(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device)


def _like(result, *inputs):
    """Device tensor if any input was one, numpy otherwise."""
    return result if any(isinstance(i, torch.Tensor) for i in inputs) else result.cpu().numpy()


# -- quantities of interest --------------------------------------------------------------------
@dataclass(frozen=True)
class Compliance:
    """J = f^T u; self-adjoint under homogeneous supports (ref fem/sensitivity.py:72-87)."""
    self_adjoint: bool = True

    def value(self, problem, u):
        return float(torch.dot(problem.load_d, _dev(u, problem.load_d.device)))

    def dJ_du(self, problem, u):
        return problem.load_d


@dataclass(frozen=True)
class PointValue:
    """J = u[dof]: the unit-load method (ref fem/sensitivity.py:90-107)."""
    dof: int
    self_adjoint: bool = False

    def value(self, problem, u):
        return float(u[self.dof])

    def dJ_du(self, problem, u):
        e = torch.zeros(problem.space.n_dofs, dtype=torch.float64, device=problem.space.device)
        e[self.dof] = 1.0
        return e


# -- stress-based quantities of interest (plane strain; ref fem/sensitivity.py:108-276) --------
class _VonMisesStress:
    """Per-element von Mises stress and d(von Mises)/du_e from `fes_von_mises_gradient`; the adjoint
    load is scattered with the reference's own summation order (`fes_node_scatter`)."""

    def __init__(self, space, material, region=None):
        if space.spatial_dim != 2 or space.n_components != 2:
            raise NotImplementedError('stress quantities of interest handle the plane-strain 2D case only')
        self.space, self.material = space, material
        self.region = None if region is None else np.asarray(region, dtype=bool)

    def evaluate(self, u, gradient=True):
        sp, m = self.space, self.material
        n_el, N = sp._en_d.shape
        ud = _dev(u, sp.device).reshape(-1)
        E_elem = None
        if m.per_element():
            E_elem = _dev(m.E, sp.device)
            if E_elem.numel() != n_el:
                raise ValueError(f'per-element modulus has {E_elem.numel()} entries but the mesh has {n_el} elements')
        vm = torch.empty(n_el, dtype=torch.float64, device=sp.device)
        dvm = torch.empty((n_el, N, 2), dtype=torch.float64, device=sp.device) if gradient else None
        _lib.check(_lib.lib().fes_von_mises_gradient(
            sp.element_type.KIND, _lib.ptr(sp._en_d), n_el, _lib.ptr(sp._coords_d), _lib.ptr(ud), _lib.ptr(E_elem),
            0.0 if m.per_element() else float(m.E), float(m.nu), sp.element_type.default_quadrature_degree(),
            _lib.ptr(vm), _lib.ptr(dvm), _lib.stream()))
        return vm, dvm

    def weights(self):
        vol = self.space.element_volumes_d
        if self.region is not None:
            vol = torch.where(torch.as_tensor(self.region, device=vol.device), vol, torch.zeros_like(vol))
        return vol / vol.sum()

    def scatter(self, per_element):
        return self.space.node_scatter(per_element, 2).reshape(-1)


@dataclass(frozen=True, eq=False)
class MeanStress:
    """Volume-weighted mean von Mises stress over a region (ref fem/sensitivity.py:202-227)."""
    space: object
    material: object
    region: object = None
    self_adjoint: bool = False

    def _stress(self):
        return _VonMisesStress(self.space, self.material, self.region)

    def value(self, problem, u):
        s = self._stress()
        return float(torch.dot(s.weights(), s.evaluate(u, gradient=False)[0]))

    def dJ_du(self, problem, u):
        s = self._stress()
        _, dvm = s.evaluate(u)
        return s.scatter(s.weights()[:, None, None] * dvm)


@dataclass(frozen=True, eq=False)
class SoftMaxStress:
    """The volume-weighted p-norm of the von Mises stress, a smooth peak (ref fem/sensitivity.py:230-276)."""
    space: object
    material: object
    region: object = None
    p: float = 8.0
    self_adjoint: bool = False

    def _stress(self):
        return _VonMisesStress(self.space, self.material, self.region)

    def value(self, problem, u):
        s = self._stress()
        vm = s.evaluate(u, gradient=False)[0]
        return float(torch.dot(s.weights(), vm ** self.p) ** (1.0 / self.p))

    def dJ_du(self, problem, u):
        s = self._stress()
        w = s.weights()
        vm, dvm = s.evaluate(u)
        total = float(torch.dot(w, vm ** self.p))
        if total <= 0.0:
            return torch.zeros(self.space.n_dofs, dtype=torch.float64, device=self.space.device)
        dJ_dvm = total ** (1.0 / self.p - 1.0) * w * vm ** (self.p - 1.0)
        return s.scatter(dJ_dvm[:, None, None] * dvm)


# -- parameterizations -------------------------------------------------------------------------
class _ElementModulusParameterization:
    """Per-element scaling of a solid stiffness K0_e at modulus `_E0` (ref fem/sensitivity.py:288-309)."""

    def _energy(self, w):
        sp = self.space
        n_el = len(sp.element_nodes)
        q = torch.empty(n_el, dtype=torch.float64, device=sp.device)
        _lib.check(_lib.lib().fes_elastic_energy(sp.element_type.KIND, _lib.ptr(sp._en_d), n_el, _lib.ptr(sp._coords_d),
                                                 _lib.ptr(w), float(self._E0), float(self.nu), None, _lib.ptr(q), None,
                                                 None, _lib.stream()))
        return q

    def _element_quadratic(self, u, lam):
        """lam_e^T K0_e u_e, one scalar per element (device tensor)."""
        dev = self.space.device
        ud = _dev(u, dev)
        if lam is u:
            return self._energy(ud)
        ld = _dev(lam, dev)
        # polarization of the element strain energy, on vectors NORMALISED to unit size: with |u| and |lam| many
        # orders apart (a stress adjoint at E ~ 1e11 scales like E/h against 1/E) the plain difference
        # Q(u+lam) - Q(u-lam) cancels catastrophically; a b (Q(u/a + lam/b) - Q(u/a - lam/b)) / 4 does not.
        a, b = float(ud.abs().max()), float(ld.abs().max())
        if a == 0.0 or b == 0.0:
            return torch.zeros(len(self.space.element_nodes), dtype=torch.float64, device=dev)
        un, ln = ud / a, ld / b
        return (self._energy(un + ln) - self._energy(un - ln)) * (0.25 * a * b)


class DensityField(_ElementModulusParameterization):
    """SIMP density E(rho) = rho^p E0: gradient -p rho^(p-1) lam_e^T K0_e u_e (ref fem/sensitivity.py:312-348)."""

    def __init__(self, space, nu, base_E, rho, penalty=3.0):
        self.space, self.nu, self._E0, self.penalty = space, float(nu), float(base_E), float(penalty)
        self.rho = _dev(rho, space.device)
        if self.rho.numel() != len(space.element_nodes):
            raise ValueError(f'rho has {self.rho.numel()} entries but the mesh has {len(space.element_nodes)} elements')

    @classmethod
    def create(cls, space, rho, base_E, nu, penalty=3.0):
        return cls(space, nu, base_E, rho, penalty)

    def with_density(self, rho):
        return DensityField(self.space, self.nu, self._E0, rho, self.penalty)

    @property
    def size(self):
        return int(self.rho.numel())

    def gradient(self, problem, u, lam):
        quad = self._element_quadratic(u, lam)
        return _like(-self.penalty * self.rho ** (self.penalty - 1.0) * quad, u, lam)


class ModulusField(_ElementModulusParameterization):
    """A per-element Young's modulus differentiated directly: gradient -lam_e^T K0_e u_e at unit modulus
    (ref fem/sensitivity.py:351-375)."""

    def __init__(self, space, nu, E):
        self.space, self.nu, self._E0 = space, float(nu), 1.0
        self.E = _dev(E, space.device)

    @classmethod
    def create(cls, space, E, nu):
        return cls(space, nu, E)

    def with_modulus(self, E):
        return ModulusField(self.space, self.nu, E)

    @property
    def size(self):
        return int(self.E.numel())

    def gradient(self, problem, u, lam):
        return _like(-self._element_quadratic(u, lam), u, lam)


# -- the driver --------------------------------------------------------------------------------
class SensitivityAnalysis:
    """Forward solve, adjoint solve and gradient on one prepared system (ref fem/sensitivity.py:380-416)."""

    def __init__(self, problem, backend=None):
        self.problem = problem
        _, _, fixed_values = problem.constraints
        self._homogeneous_supports = bool(np.allclose(fixed_values, 0.0))
        self._system = DiscreteSystem(problem.tangent(None), problem.constraints, backend)

    def solve_forward(self, device=False, x0=None):
        """u of K u = f (numpy like the reference; `device=True` keeps it in HBM)."""
        u = self._system.solve(self.problem.load_d, x0=x0)
        return u if device else u.cpu().numpy()

    def adjoint(self, qoi, u):
        if qoi.self_adjoint and self._homogeneous_supports:
            return u
        lam = self._system.solve_homogeneous(_dev(qoi.dJ_du(self.problem, u), self.problem.space.device))
        return _like(lam, u)

    def gradient(self, qoi, parameterization, u):
        lam = self.adjoint(qoi, u)
        return parameterization.gradient(self.problem, u, lam)
