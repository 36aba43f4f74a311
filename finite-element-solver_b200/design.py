"""Optimality-criteria density update on the device (ref fem/design.py:39-78): the whole
bisection on the Lagrange multiplier runs in one persistent cooperative kernel."""
import ctypes as C

import numpy as np
import torch

from . import _lib


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
