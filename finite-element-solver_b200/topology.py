"""SIMP topology optimization with the whole design iteration resident on the GPU.

Same constructor, `solve(on_iteration) -> TopologyHistory`, `set_rho`, `dilution`,
`scaled_modulus`, `oc_density` and `_solve` as the reference's TopologyOptimizer
(ref fem/topology.py:143-330); `calculate_smoothing_matrix` returns the same row-normalised
cone filter (ref fem/topology.py:39-77).  Per iteration: rho^p-scaled fused reassembly into the
cached CSR pattern -> Jacobi-PCG on the free block -> per-element u_e^T K0_e u_e, compliance and
von Mises -> sensitivity -> cone-filter SpMV -> OC bisection; rho, K, u never leave HBM.
"""
import ctypes as C
import logging
from dataclasses import dataclass

import numpy as np
import torch
from scipy.sparse import csr_array

from . import _lib
from .backends import JacobiPCGBackend
from .boundary import BoundaryConditions
from .csr import DeviceCSR
from .design import optimality_criteria_update
from .equations import LinearElastic
from .forms import Lowered
from .materials import LinearElasticMaterial
from .forms import LinearElasticForm
from .problem import LinearProblem
from .space import FunctionSpace

logger = logging.getLogger(__name__)


class ConeFilter:
    """Device-resident cone filter over element centroids."""

    def __init__(self, space, radius):
        self.n = len(space.element_nodes)
        n_corner = space.mesh.elements.shape[1]
        en = space._en_d if space.element_type.N == n_corner else \
            torch.as_tensor(space.mesh.elements.astype(np.int32)).to(space.device).contiguous()
        coords = space._coords_d
        h = C.c_void_p()
        _lib.check(_lib.lib().fes_filter_create(_lib.ptr(en), self.n, n_corner, _lib.ptr(coords), space.spatial_dim,
                                                float(radius), _lib.stream(), C.byref(h)))
        self._h = h
        self.nnz = int(_lib.lib().fes_filter_nnz(h))

    def apply(self, x):
        y = torch.empty_like(x)
        _lib.check(_lib.lib().fes_filter_apply(self._h, _lib.ptr(x.contiguous()), _lib.ptr(y), _lib.stream()))
        return y

    def to_scipy(self):
        indptr = np.empty(self.n + 1, dtype=np.int64)
        indices = np.empty(self.nnz, dtype=np.int64)
        data = np.empty(self.nnz, dtype=np.float64)
        _lib.check(_lib.lib().fes_filter_export_csr(self._h, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data)))
        return csr_array((data, indices, indptr), shape=(self.n, self.n))

    def __del__(self):
        try:
            if self._h:
                _lib.lib().fes_filter_destroy(self._h)
                self._h = None
        except Exception:
            pass


def calculate_smoothing_matrix(mesh, r):
    """Host csr_array of the row-normalised cone weights, built on the device."""
    space = FunctionSpace(mesh, n_components=mesh.spatial_dim)
    return ConeFilter(space, r).to_scipy()


@dataclass(frozen=True)
class MinCompliance:
    def gradient(self, compliance, rho, penalty):
        return compliance * penalty / rho

    def chain(self, total_compliance):
        return 1.0


@dataclass(frozen=True)
class TargetCompliance:
    target: float

    def gradient(self, compliance, rho, penalty):
        return (compliance * penalty / rho) * 2 * (compliance.sum() - self.target)

    def chain(self, total_compliance):
        return 2.0 * (total_compliance - self.target)


@dataclass(frozen=True, eq=False)
class TopologyHistory:
    rho: list
    u: list
    von_mises: list
    compliance: list


class SimpSolution:
    """What one design iteration produced; device tensors with lazy host views."""

    def __init__(self, u_d, compliance_d, von_mises_d, q_d, total_compliance, cg_iterations):
        self.u_d, self.compliance_d, self.von_mises_d, self.q_d = u_d, compliance_d, von_mises_d, q_d
        self.total_compliance, self.cg_iterations = total_compliance, cg_iterations

    @property
    def u(self):
        return self.u_d.cpu().numpy()

    @property
    def compliance(self):
        return self.compliance_d.cpu().numpy()

    @property
    def von_mises(self):
        return self.von_mises_d.cpu().numpy()


class TopologyOptimizer:
    def __init__(self, mesh, equation, boundary_conditions=None, iters=10, volume_frac=1.0, smoothing_radius=0.1,
                 penalty=3.0, objective=None, strategy=None, backend=None, warm_start=True, history='host'):
        assert isinstance(equation, LinearElastic), 'TopologyOptimizer only supports LinearElastic equations'
        self.mesh = mesh
        self.bc = boundary_conditions if boundary_conditions is not None else BoundaryConditions()
        self.base_E, self.nu, self.source = equation.E, equation.nu, equation.source
        self.iters, self.volume_frac, self.penalty = iters, volume_frac, float(penalty)
        self.objective = objective if objective is not None else MinCompliance()
        self.strategy = strategy   # stored and never used, exactly as the reference does (fem/topology.py:181)
        self.backend = backend if backend is not None else JacobiPCGBackend()
        self.warm_start, self.history_mode = warm_start, history
        self.space = FunctionSpace(mesh, n_components=mesh.spatial_dim)
        dev = self.space.device
        n_el = len(mesh.elements)
        self._rho = torch.full((n_el,), float(volume_frac), dtype=torch.float64, device=dev)
        self._filter = ConeFilter(self.space, smoothing_radius)
        self._E_elem = None
        if not np.isscalar(self.base_E):
            self._E_elem = torch.as_tensor(np.asarray(self.base_E, dtype=np.float64)).to(dev)
            if self._E_elem.numel() != n_el:
                raise ValueError(f'per-element modulus has {self._E_elem.numel()} entries but the mesh has '
                                 f'{n_el} elements')
        # constraints + load once (the density never reaches them, ref fem/topology.py:199-201)
        self._problem = LinearProblem(self.space, LinearElasticForm(LinearElasticMaterial(1.0, self.nu)),
                                      self.source, self.bc)
        free, fixed, vals = self._problem.constraints
        plan = self.space.plan()
        self._values = torch.zeros(plan.nnz, dtype=torch.float64, device=dev)
        self._A = DeviceCSR((self.space.n_dofs,) * 2, plan.indptr_d, plan.indices_d, self._values, plan.host_pattern)
        self._A._plan = plan
        self._mask = torch.zeros(self.space.n_dofs, dtype=torch.uint8, device=dev)
        self._mask[torch.as_tensor(free, device=dev)] = 1
        self._x_fixed = torch.zeros(self.space.n_dofs, dtype=torch.float64, device=dev)
        if len(fixed):
            self._x_fixed[torch.as_tensor(fixed, device=dev)] = torch.as_tensor(np.asarray(vals, float), device=dev)
        self._lifting = bool(np.any(np.asarray(vals) != 0.0))
        self._dilution = torch.empty(n_el, dtype=torch.float64, device=dev)
        self._scale = torch.empty(n_el, dtype=torch.float64, device=dev) if self._E_elem is not None else None
        self._u = None
        self._solver = None
        self._last = None
        self.history = None
        self.cg_iterations = []

    # -- the reference's small surface ------------------------------------------------------------
    @property
    def rho(self):
        return self._rho.cpu().numpy()

    @rho.setter
    def rho(self, value):   # a plain attribute in the reference (fem/topology.py:203)
        self.set_rho(value)

    @property
    def rho_d(self):
        return self._rho

    def set_rho(self, rho):
        self._rho = (rho if isinstance(rho, torch.Tensor) else
                     torch.as_tensor(np.ascontiguousarray(rho, dtype=np.float64))).to(self.space.device).contiguous()

    @property
    def dilution(self):
        return (self._rho ** self.penalty).cpu().numpy()

    @property
    def scaled_modulus(self):
        return self.dilution * self.base_E

    @property
    def smoothing_matrix(self):
        return self._filter.to_scipy()

    def _volume_fraction(self, rho):
        v = self.space.element_volumes
        return float((v * rho).sum() / v.sum())

    # -- one design iteration's solve ---------------------------------------------------------------
    def _solve(self):
        L, sp = _lib.lib(), self.space
        n_el = len(sp.element_nodes)
        _lib.check(L.fes_pow(_lib.ptr(self._rho), self.penalty, n_el, _lib.ptr(self._dilution), _lib.stream()))
        E0 = 1.0 if self._E_elem is not None else float(self.base_E)
        low = Lowered(_lib.FORM_ELASTIC, E=E0, nu=float(self.nu), E_elem=self._E_elem, scale_elem=self._dilution)
        cf = low.coeffs()
        _lib.check(L.fes_assemble(sp.plan().handle, sp.element_type.KIND, low.form, sp.spatial_dim,
                                  _lib.ptr(sp._en_d), _lib.ptr(sp._coords_d), C.byref(cf), _lib.ptr(self._values),
                                  _lib.stream()))
        robin = self._problem._robin_operator
        if robin is not None:
            # the tangent of problem.with_operator(stiffness) includes the Robin term (ref fem/problem.py:138-140,
            # fem/topology.py:252): add kappa * M_boundary on its sub-pattern, in place
            _lib.check(L.fes_csr_add_subpattern(self.space.n_dofs, _lib.ptr(self._A.indptr_d), _lib.ptr(self._A.indices_d),
                                                _lib.ptr(self._values), _lib.ptr(robin.indptr_d), _lib.ptr(robin.indices_d),
                                                _lib.ptr(robin.data_d), 1.0, _lib.stream()))
        if self._solver is None:
            self._solver = self.backend.prepare(self._A, self._mask)
        else:
            self._solver.refresh()
        b = self._problem.load_d
        rhs = self._solver.lift(b, self._x_fixed) if self._lifting else b
        if self._u is None or not self.warm_start:
            u = self._x_fixed.clone()
            self._solver.solve_device(rhs, u, warm_start=False)
        else:
            u = self._u.clone()
            self._solver.solve_device(rhs, u, warm_start=True)
        self._u = u
        self.cg_iterations.append(self._solver.last_iterations)

        q = torch.empty(n_el, dtype=torch.float64, device=sp.device)
        comp, vm = torch.empty_like(q), torch.empty_like(q)
        scale = self._dilution
        if self._E_elem is not None:
            torch.mul(self._dilution, self._E_elem, out=self._scale)
            scale = self._scale
        _lib.check(L.fes_elastic_energy(sp.element_type.KIND, _lib.ptr(sp._en_d), n_el, _lib.ptr(sp._coords_d),
                                        _lib.ptr(u), E0, float(self.nu), _lib.ptr(scale), _lib.ptr(q), _lib.ptr(comp),
                                        _lib.ptr(vm), _lib.stream()))
        if self._E_elem is not None:
            q = q * self._E_elem
        total = C.c_double(0.0)
        _lib.check(L.fes_dot(_lib.ptr(comp), None, n_el, C.byref(total), _lib.stream()))
        self._last = SimpSolution(u, comp, vm, q, total.value, self._solver.last_iterations)
        return self._last

    def _compliance_sensitivity(self, solution, chain=1.0):
        """+p rho^(p-1) u_e^T K0_e u_e (the negated adjoint gradient, ref fem/topology.py:261-274)."""
        sens = torch.empty_like(self._rho)
        _lib.check(_lib.lib().fes_simp_sensitivity(_lib.ptr(self._rho), _lib.ptr(solution.q_d), self.penalty,
                                                   float(chain), self._rho.numel(), _lib.ptr(sens), _lib.stream()))
        return sens

    def oc_density(self, sensitivity, volume_frac, max_iters=100, tol=1e-8):
        return optimality_criteria_update(self._rho, sensitivity, self.space.element_volumes_d, volume_frac,
                                          move=0.1, max_iters=max_iters, tol=tol)

    def step(self):
        """Solve at the current density, then advance it.  Returns the SimpSolution."""
        solution = self._solve()
        sens = self._compliance_sensitivity(solution, self.objective.chain(solution.total_compliance))
        self._rho = self.oc_density(self._filter.apply(sens), self.volume_frac)
        return solution

    def solve(self, on_iteration=None):
        keep = self.history_mode
        rho_s, u_s, vm_s, c_s = [], [], [], []
        grab = (lambda t: t.cpu().numpy()) if keep == 'host' else (lambda t: t.clone())
        for i in range(self.iters):
            rho_before = self._rho
            solution = self._solve()
            if keep:
                rho_s.append(grab(rho_before)); u_s.append(grab(solution.u_d))
                vm_s.append(grab(solution.von_mises_d)); c_s.append(grab(solution.compliance_d))
            logger.info('Iteration %d: total compliance = %.4f, cg iterations = %d', i,
                        solution.total_compliance, solution.cg_iterations)
            if on_iteration is not None:
                on_iteration(i, solution)
            sens = self._compliance_sensitivity(solution, self.objective.chain(solution.total_compliance))
            self._rho = self.oc_density(self._filter.apply(sens), self.volume_frac)
        self.history = TopologyHistory(rho_s, u_s, vm_s, c_s)
        return self.history
