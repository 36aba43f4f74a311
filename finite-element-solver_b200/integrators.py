"""Time integrators on the device (ref fem/integrators.py:33-133): same classes, constructor
arguments and `run` signatures as the reference.

The effective operator (M + theta dt K, or M + beta dt^2 K) and the right-hand-side operator are
formed once on the SHARED CSR pattern (mass and stiffness come from the same scatter plan, so the
sum is an axpy of two value arrays), the system goes through `DiscreteSystem` once, and every step
is device work only: one SpMV for the right-hand side and one Jacobi-PCG solve warm-started from the
previous state.  Histories are returned as host arrays like the reference's solutions.
"""
import numpy as np
import torch

from .system import DiscreteSystem


from .solution import TransientSolution, WaveSolution  # noqa: E402,F401  (typed solutions live in solution.py)


def _dev(x, device):
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device)


def _solve(system, rhs, x0=None):
    """system.solve on device vectors; a foreign backend (host path of DiscreteSystem) gets host arrays."""
    if system._device_path:
        return system.solve(rhs, x0=x0)
    return _dev(system.solve(rhs.cpu().numpy()), rhs.device)


class ThetaMethod:
    """First-order integrator for M u' + K u = b: (M + theta dt K) u_{n+1} = (M - (1-theta) dt K) u_n + dt b
    (ref fem/integrators.py:33-64).  theta = 1/2 is Crank-Nicolson, theta = 1 backward Euler."""

    def __init__(self, dt, steps, theta=0.5, backend=None):
        self.dt, self.steps, self.theta, self.backend = dt, steps, theta, backend

    def run(self, problem, u0):
        space = problem.space
        M, K, b = space.mass_matrix_d, problem.tangent(None), problem.load_d
        dt, theta = self.dt, self.theta
        system = DiscreteSystem(M.add_subpattern(K, theta * dt), problem.constraints, self.backend)
        rhs_operator = M.add_subpattern(K, -(1.0 - theta) * dt)
        u = _dev(u0, space.device)
        if u.numel() != space.n_dofs:
            raise ValueError(f'u0 has {u.numel()} entries but the space has {space.n_dofs} dofs')
        t_values, u_values = [0.0], [u.cpu().numpy()]
        for i in range(self.steps):
            u = _solve(system, rhs_operator.matvec(u) + dt * b, x0=u.clone())
            t_values.append(dt * (i + 1))
            u_values.append(u.cpu().numpy())
        return TransientSolution(space.mesh, space.n_components, np.asarray(t_values), u_values)


class NewmarkMethod:
    """Second-order integrator for M u'' + K u = b (ref fem/integrators.py:67-120): beta = 1/4, gamma = 1/2 is
    the average-acceleration scheme.  Accelerations are solved against M + beta dt^2 K with the
    Dirichlet dofs pinned to zero."""

    def __init__(self, dt, steps, beta=0.25, gamma=0.5, backend=None):
        self.dt, self.steps, self.beta, self.gamma, self.backend = dt, steps, beta, gamma, backend

    def run(self, problem, u0, v0):
        space = problem.space
        M, K, b = space.mass_matrix_d, problem.tangent(None), problem.load_d
        free, fixed, fixed_values = problem.constraints
        u_h, v_h = np.asarray(u0, dtype=float), np.asarray(v0, dtype=float)
        if not np.allclose(u_h[fixed], fixed_values):
            raise ValueError('u0 disagrees with the Dirichlet values at fixed nodes')
        if not np.allclose(v_h[fixed], 0):
            raise ValueError('v0 must be zero at Dirichlet-fixed nodes')
        dt, beta, gamma = self.dt, self.beta, self.gamma
        accel_constraints = (free, fixed, np.zeros(len(fixed)))
        u, v = _dev(u_h, space.device), _dev(v_h, space.device)
        a = _solve(DiscreteSystem(M, accel_constraints, self.backend), b - K.matvec(u))
        effective = DiscreteSystem(M.add_subpattern(K, beta * dt ** 2), accel_constraints, self.backend)
        t_values, u_values, dudt_values = [0.0], [u.cpu().numpy()], [v.cpu().numpy()]
        for i in range(self.steps):
            u_pred = u + dt * v + (dt ** 2 / 2 * (1 - 2 * beta)) * a
            v_pred = v + (dt * (1 - gamma)) * a
            a = _solve(effective, b - K.matvec(u_pred), x0=a.clone())
            u = u_pred + (beta * dt ** 2) * a
            v = v_pred + (gamma * dt) * a
            t_values.append(dt * (i + 1))
            u_values.append(u.cpu().numpy())
            dudt_values.append(v.cpu().numpy())
        return WaveSolution(space.mesh, space.n_components, np.asarray(t_values), u_values, dudt_values)


def wave_energy(problem, u, v):
    """Total wave energy (u^T K u + v^T M v) / 2 (ref fem/integrators.py:123-133)."""
    dev = problem.space.device
    ud, vd = _dev(u, dev), _dev(v, dev)
    K, M = problem.tangent(None), problem.space.mass_matrix_d
    return float(0.5 * (torch.dot(ud, K.matvec(ud)) + torch.dot(vd, M.matvec(vd))))
