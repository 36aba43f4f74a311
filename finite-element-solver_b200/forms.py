"""Bilinear forms as descriptors (ref fem/forms.py:190-258,318-333,444-480).

In the reference a Form computes (n_el,k,k) element matrices with numpy and the space scatters
them.  Here a Form only *describes* the integrand -- `lower(space, boundary)` returns the
(form code, coefficients) the fused assembly kernel evaluates per element block -- so element
matrices never exist in HBM.  `element_matrices(space)` is still offered (unfused kernel) for
parity checks and for building the K0 of a PrecomputedForm.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .materials import LinearElasticMaterial


@dataclass
class Lowered:
    form: int
    E: float = 1.0
    nu: float = 0.0
    factor: float = 1.0
    E_elem: object = None      # device fp64 (n_el) or None
    scale_elem: object = None  # device fp64 (n_el) or None

    def coeffs(self):
        return _lib.Coeffs(self.E, self.nu, self.factor, _lib.ptr(self.E_elem).value, _lib.ptr(self.scale_elem).value)


def _per_element(values, n_el, device, what):
    t = values if isinstance(values, torch.Tensor) else torch.as_tensor(np.asarray(values, dtype=np.float64))
    if t.numel() != n_el:
        raise ValueError(f'{what} has {t.numel()} entries but the mesh has {n_el} elements')
    t = t.to(device=device, dtype=torch.float64).contiguous()
    if t.numel() and float(t.min()) < 0.0:
        # the fused kernel folds sqrt(weight) into the stored gradients
        raise ValueError(f'{what} must be nonnegative')
    return t


class _Form:
    def element_matrices(self, space, boundary=False):
        """(n_el, k, k) device tensor from the unfused kernel."""
        return space._element_matrices(self, boundary)


@dataclass(frozen=True)
class LaplacianForm(_Form):
    def lower(self, space, boundary=False):
        return Lowered(_lib.FORM_LAPLACIAN)


@dataclass(frozen=True)
class MassForm(_Form):
    n_components: int = 1

    def lower(self, space, boundary=False):
        if self.n_components != space.n_components:
            raise ValueError(f'MassForm({self.n_components}) on a space with {space.n_components} components')
        return Lowered(_lib.FORM_MASS)


@dataclass(frozen=True, eq=False)
class MaskedMassForm(_Form):
    n_components: int
    mask: object   # one 0/1 entry per facet

    def lower(self, space, boundary=False):
        n_el = space._n_elements(boundary)
        mask = self.mask if isinstance(self.mask, torch.Tensor) else np.asarray(self.mask, dtype=np.float64)
        return Lowered(_lib.FORM_MASS, scale_elem=_per_element(mask, n_el, space.device, 'facet mask'))


@dataclass(frozen=True, eq=False)
class LinearElasticForm(_Form):
    material: LinearElasticMaterial

    def lower(self, space, boundary=False):
        m = self.material
        if m.per_element():
            E = _per_element(m.E, space._n_elements(boundary), space.device, 'per-element modulus')
            return Lowered(_lib.FORM_ELASTIC, E=1.0, nu=float(m.nu), E_elem=E)
        return Lowered(_lib.FORM_ELASTIC, E=float(m.E), nu=float(m.nu))


@dataclass(frozen=True, eq=False)
class ScaledForm(_Form):
    factor: float
    form: object

    def lower(self, space, boundary=False):
        low = self.form.lower(space, boundary)
        low.factor *= float(self.factor)
        return low


@dataclass(frozen=True, eq=False)
class PrecomputedForm:
    """Element matrices computed elsewhere (device tensor or ndarray), optionally with a
    per-element scale applied on the fly (SIMP rho^p)."""
    matrices: object
    scale: object = None

    def lower(self, space, boundary=False):
        return None  # handled by FunctionSpace._assemble_values

    def element_matrices(self, space, boundary=False):
        n_el = space._n_elements(boundary)
        if len(self.matrices) != n_el:
            raise ValueError(f'precomputed matrices cover {len(self.matrices)} elements but the '
                             f'geometry has {n_el}')
        return self.matrices
