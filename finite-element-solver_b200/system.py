"""A x = b with the Dirichlet DOFs eliminated (ref fem/system.py:23-65), on the device.

Same constructor and `solve` / `solve_homogeneous` as the reference's DiscreteSystem.  With a
B200 backend the free-free block is never formed: the operator stays a full n x n CSR in HBM
and a 0/1 free mask restricts the PCG iteration to the free rows and columns, which is
arithmetically the reference's A[ix_(free, free)] solve.  A foreign backend (any object with
`prepare(A_ff)`) gets the reference's host slicing instead, so the seam stays a drop-in both ways.
"""
import numpy as np
import torch
from scipy.sparse import csr_array

from . import _lib
from .backends import JacobiPCGBackend
from .csr import DeviceCSR


class DiscreteSystem:
    def __init__(self, A, constraints, backend=None):
        free, fixed, fixed_values = constraints
        self.n_dofs = A.shape[0]
        self.free = np.asarray(free, dtype=int)
        self.fixed = np.asarray(fixed, dtype=int)
        self.fixed_values = np.asarray(fixed_values, dtype=float)
        backend = backend if backend is not None else JacobiPCGBackend()
        self._device_path = isinstance(backend, JacobiPCGBackend)
        if self._device_path:
            dev = A.device if isinstance(A, DeviceCSR) else _lib.require_cuda()
            self._A = A if isinstance(A, DeviceCSR) else DeviceCSR.from_scipy(A, dev)
            mask = torch.zeros(self.n_dofs, dtype=torch.uint8, device=dev)
            mask[torch.as_tensor(self.free, device=dev)] = 1
            self._mask = mask
            self._x_fixed = torch.zeros(self.n_dofs, dtype=torch.float64, device=dev)
            if len(self.fixed):
                self._x_fixed[torch.as_tensor(self.fixed, device=dev)] = torch.as_tensor(self.fixed_values, device=dev)
            self._lifting = bool(np.any(self.fixed_values != 0.0))
            self._solver = backend.prepare(self._A, mask)
        else:
            A_host = A.to_scipy() if isinstance(A, DeviceCSR) else (A if isinstance(A, np.ndarray) else csr_array(A))
            self._free_fixed = A_host[np.ix_(self.free, self.fixed)]
            self._solver = backend.prepare(A_host[np.ix_(self.free, self.free)])

    @property
    def solver(self):
        return self._solver

    def _solve_device(self, b, lift, x0=None):
        host = not isinstance(b, torch.Tensor)
        bd = torch.as_tensor(np.ascontiguousarray(b, dtype=np.float64)).to(self._mask.device) if host else b
        rhs = self._solver.lift(bd, self._x_fixed) if (lift and self._lifting) else bd
        if x0 is not None:
            x = x0
            if lift:
                x[self._mask == 0] = self._x_fixed[self._mask == 0]
            self._solver.solve_device(rhs, x, warm_start=True)
        else:
            x = self._x_fixed.clone() if lift else torch.zeros_like(self._x_fixed)
            self._solver.solve_device(rhs, x, warm_start=False)
        return x.cpu().numpy() if host else x

    def solve(self, b, x0=None):
        """x with x[fixed] = fixed_values and the free block solved; reuses the prepared solver."""
        if self._device_path:
            return self._solve_device(b, lift=True, x0=x0)
        x = np.zeros(self.n_dofs)
        x[self.fixed] = self.fixed_values
        x[self.free] = self._solver.solve(b[self.free] - self._free_fixed @ self.fixed_values)
        return x

    def solve_homogeneous(self, b):
        """Same solve with the fixed DOFs pinned to zero (the adjoint's supports)."""
        if self._device_path:
            return self._solve_device(b, lift=False)
        x = np.zeros(self.n_dofs)
        x[self.free] = self._solver.solve(b[self.free])
        return x
