"""Design optimization on the device (ref fem/design.py): the optimality-criteria update (the whole
bisection on the Lagrange multiplier in one persistent cooperative kernel), `SIMPModel` and
`DesignOptimizer` with the reference's constructors, `step()` and `solve()`."""
import ctypes as C
import logging
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

logger = logging.getLogger(__name__)


def _dev(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(device)


def optimality_criteria_update(rho, sensitivity, volumes, volume_frac, move=0.1, max_iters=100, tol=1e-8,
                               return_bisections=False):
    """One OC step.  numpy in -> numpy out, device tensors in -> device tensor out.
    Raises ValueError (message contains 'nonnegative') on a signed sensitivity."""
    host = not isinstance(rho, torch.Tensor)
    device = _lib.require_cuda() if host else rho.device
    r, s, v = _dev(rho, device), _dev(sensitivity, device), _dev(volumes, device)
    if not (r.numel() == s.numel() == v.numel()):
        raise ValueError('rho, sensitivity and volumes must have one entry per element')
    out = torch.empty_like(r)
    n_bis = C.c_int(0)
    _lib.check(_lib.lib().fes_oc_update(_lib.ptr(r), _lib.ptr(s), _lib.ptr(v), r.numel(), float(volume_frac),
                                        float(move), int(max_iters), float(tol), _lib.ptr(out), C.byref(n_bis),
                                        _lib.stream()))
    res = out.cpu().numpy() if host else out
    return (res, n_bis.value) if return_bisections else res


class SIMPModel:
    """A SIMP density design of a linear-elastic problem (ref fem/design.py:81-124): the diluted problem at a
    density (E(rho) = rho^p E0 as a per-element modulus of the fused assembly; constraints and load are
    resolved once) and the `DensityField` that differentiates it.  `sensitivity_filter` is a `ConeFilter`,
    a scipy matrix or None."""

    def __init__(self, space, base_E, nu, bc, source=None, penalty=3.0, sensitivity_filter=None):
        from .forms import LinearElasticForm
        from .materials import LinearElasticMaterial
        from .problem import LinearProblem
        from .sensitivity import DensityField
        self.space, self.base_E, self.nu, self.bc, self.source = space, float(base_E), float(nu), bc, source
        self.penalty, self.sensitivity_filter = float(penalty), sensitivity_filter
        n_el = len(space.element_nodes)
        self._density = DensityField.create(space, np.ones(n_el), base_E, nu, penalty)
        self._base = LinearProblem(space, LinearElasticForm(LinearElasticMaterial(self.base_E, self.nu)), source, bc)

    @property
    def volumes(self):
        return self.space.element_volumes

    def parameterization(self, rho):
        return self._density.with_density(rho)

    def problem(self, rho):
        from .forms import LinearElasticForm
        from .materials import LinearElasticMaterial
        r = _dev(rho, self.space.device)
        return self._base.with_operator(LinearElasticForm(LinearElasticMaterial(self.base_E * r ** self.penalty, self.nu)))


@dataclass(frozen=True, eq=False)
class DesignHistory:
    """The per-iteration series a design optimization produces (ref fem/design.py:127-132)."""
    rho: list
    u: list
    objective: list


class DesignOptimizer:
    """Minimize a QuantityOfInterest over a SIMP density under a volume constraint (ref fem/design.py:135-192).
    The density, the state and the sensitivities stay on the device between steps."""

    def __init__(self, model, objective=None, volume_frac=0.5, iters=30, move=0.2, backend=None):
        from .sensitivity import Compliance
        self.model = model
        self.objective = objective if objective is not None else Compliance()
        self.volume_frac, self.iters, self.move, self.backend = volume_frac, iters, move, backend
        self._rho = torch.full((len(model.space.element_nodes),), float(volume_frac), dtype=torch.float64,
                               device=model.space.device)
        self.history = None

    @property
    def rho(self):
        return self._rho.cpu().numpy()

    def step(self):
        """One iteration: solve, score, advance the density.  Returns (u, J) like the reference."""
        from .sensitivity import SensitivityAnalysis
        problem = self.model.problem(self._rho)
        analysis = SensitivityAnalysis(problem, self.backend)
        u = analysis.solve_forward(device=True)
        value = self.objective.value(problem, u)
        gradient = analysis.gradient(self.objective, self.model.parameterization(self._rho), u)
        sensitivity = -gradient
        f = self.model.sensitivity_filter
        if f is not None:
            sensitivity = f.apply(sensitivity) if hasattr(f, 'apply') else _dev(f @ sensitivity.cpu().numpy(), sensitivity.device)
        self._rho = optimality_criteria_update(self._rho, sensitivity, self.model.space.element_volumes_d,
                                               self.volume_frac, self.move)
        return u.cpu().numpy(), value

    def solve(self, on_iteration=None):
        rho_s, u_s, obj_s = [], [], []
        vol = self.model.volumes
        for i in range(self.iters):
            rho_before = self.rho
            u, value = self.step()
            rho_s.append(rho_before); u_s.append(u); obj_s.append(value)
            logger.info('Design iteration %d: objective = %.6g, volume fraction = %.4f', i, value,
                        float((vol * rho_before).sum() / vol.sum()))
            if on_iteration is not None:
                on_iteration(i, rho_before, value)
        self.history = DesignHistory(rho_s, u_s, obj_s)
        return self.history
