"""The linear-algebra backend slot (ref fem/backends.py:62-116,186-213): Jacobi-PCG on the GPU.

`JacobiPCGBackend().prepare(A)` satisfies the reference's `Backend` protocol: it accepts the
host `csr_array` (or dense ndarray) `DiscreteSystem` hands over and returns a `LinearSolver`
whose `solve(b)` takes and returns host float64 arrays.  Non-convergence raises
RuntimeError('CG failed ...') exactly where the reference's `_CGSolver` does.  The same object
also prepares a `DeviceCSR` (optionally with a free-DOF mask) for the device-resident path.
"""
import ctypes as C

import numpy as np
import torch
from scipy.sparse import csr_array

from . import _lib
from .csr import DeviceCSR


class PCGSolver:
    """One operator, preconditioned once (Jacobi diagonal), solved against many right-hand sides."""

    def __init__(self, handle, n, rtol, maxiter, keepalive=None, device=None):
        self._h, self.n, self.rtol, self.maxiter = handle, n, rtol, maxiter
        self._keep = keepalive
        self.device = device
        self.last_iterations = 0
        self.last_relres = 0.0

    def _finish(self, status, iters, relres):
        self.last_iterations, self.last_relres = int(iters.value), float(relres.value)
        _lib.check(status)

    def solve(self, b):
        """Host drop-in: float64 ndarray in, fresh float64 ndarray out."""
        if isinstance(b, torch.Tensor):
            return self.solve_device(b)
        b = np.ascontiguousarray(b, dtype=np.float64)
        if b.shape != (self.n,):
            raise ValueError(f'right-hand side has shape {b.shape}, the operator is {self.n} x {self.n}')
        x = np.empty(self.n, dtype=np.float64)
        iters, relres = C.c_int64(0), C.c_double(0.0)
        status = _lib.lib().fes_solver_solve_host(self._h, _lib.ptr(b), _lib.ptr(x), self.rtol,
                                                  int(self.maxiter or 0), C.byref(iters), C.byref(relres))
        self._finish(status, iters, relres)
        return x

    def solve_device(self, b, x=None, warm_start=False):
        """Device path: b (n) fp64 tensor; x is updated in place on the free rows."""
        if x is None:
            x = torch.zeros(self.n, dtype=torch.float64, device=b.device)
            warm_start = False
        for name, t in (('b', b), ('x', x)):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.numel() == self.n):
                raise ValueError(f'{name} must be a CUDA float64 tensor of {self.n} entries')
        if not x.is_contiguous() or x.data_ptr() % 16:
            # the block kernels read the vector with 16-byte loads: solve on an aligned copy, write back
            xa = x.contiguous().clone()
            self.solve_device(b, xa, warm_start)
            x.copy_(xa)
            return x
        iters, relres = C.c_int64(0), C.c_double(0.0)
        status = _lib.lib().fes_solver_solve(self._h, _lib.ptr(b.contiguous()), _lib.ptr(x), self.rtol,
                                             int(self.maxiter or 0), 1 if warm_start else 0, C.byref(iters),
                                             C.byref(relres), _lib.stream())
        self._finish(status, iters, relres)
        return x

    @property
    def operator_bytes(self):
        """Bytes of the block-SELL operator copy one product reads (0 before the first solve)."""
        return int(_lib.lib().fes_solver_operator_bytes(self._h))

    def refresh(self):
        """The operator's values changed in place (same pattern): rebuild the Jacobi diagonal."""
        _lib.check(_lib.lib().fes_solver_refresh(self._h, _lib.stream()))

    def lift(self, b, x_fixed):
        """b - A x_fixed on the free rows, 0 on the fixed ones (ref fem/system.py:49-50)."""
        rhs = torch.empty_like(b)
        _lib.check(_lib.lib().fes_lift_rhs(self._h, _lib.ptr(b.contiguous()), _lib.ptr(x_fixed.contiguous()),
                                           _lib.ptr(rhs), _lib.stream()))
        return rhs

    def __del__(self):
        try:
            if self._h:
                _lib.lib().fes_solver_destroy(self._h)
                self._h = None
        except Exception:
            pass


class JacobiPCGBackend:
    """Immutable config; `prepare` binds it to an operator (ref fem/backends.py:85-116)."""

    def __init__(self, rtol=1e-10, maxiter=None):
        self.rtol, self.maxiter = rtol, maxiter

    def prepare(self, A, free_mask=None):
        L = _lib.lib()
        handle = C.c_void_p()
        if isinstance(A, DeviceCSR):
            if A.shape[0] != A.shape[1]:
                raise ValueError(f'operator must be square, got {A.shape}')
            if free_mask is not None and not (isinstance(free_mask, torch.Tensor) and free_mask.dtype == torch.uint8
                                              and free_mask.device == A.device and free_mask.is_contiguous()
                                              and free_mask.numel() == A.shape[0]):
                raise ValueError(f'free_mask must be a contiguous uint8 tensor of {A.shape[0]} entries on {A.device}')
            _lib.check(L.fes_solver_create(A.shape[0], _lib.ptr(A.indptr_d), _lib.ptr(A.indices_d),
                                           _lib.ptr(A.data_d), _lib.ptr(free_mask), _lib.stream(), C.byref(handle)))
            plan = getattr(A, '_plan', None)
            if plan is not None and A.indptr_d.data_ptr() == plan.indptr_d.data_ptr():
                _lib.check(L.fes_solver_use_plan(handle, plan.handle))
            return PCGSolver(handle, A.shape[0], self.rtol, self.maxiter, keepalive=(A, free_mask, plan),
                             device=A.device)
        _lib.require_cuda()
        if free_mask is not None:
            raise ValueError('a free mask needs a DeviceCSR operator')
        A = csr_array(A)
        if A.shape[0] != A.shape[1]:
            raise ValueError(f'operator must be square, got {A.shape}')
        if not A.has_sorted_indices:
            A = A.sorted_indices()
        indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(A.indices, dtype=np.int64)
        data = np.ascontiguousarray(A.data, dtype=np.float64)
        _lib.check(L.fes_solver_create_host(A.shape[0], _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data),
                                            C.byref(handle)))
        return PCGSolver(handle, A.shape[0], self.rtol, self.maxiter)


B200Backend = JacobiPCGBackend
