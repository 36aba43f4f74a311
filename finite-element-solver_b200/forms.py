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

    def derived_fields(self, space, u):
        """Element strain, stress (full 3x3 tensors, plane-strain lift in 2D) and compliance from the
        DOF vector `u`, at the first quadrature point of the space's default rule
        (ref fem/forms.py:335-374).  Device tensors."""
        m = self.material
        n_el = len(space.element_nodes)
        ud = u if isinstance(u, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64))
        ud = ud.to(device=space.device, dtype=torch.float64).contiguous().reshape(-1)
        if ud.numel() != space.n_nodes * space.spatial_dim:
            raise ValueError(f'expected {space.n_nodes * space.spatial_dim} displacement DOFs, got {ud.numel()}')
        E_elem = _per_element(m.E, n_el, space.device, 'per-element modulus') if m.per_element() else None
        eps = torch.empty((n_el, 3, 3), dtype=torch.float64, device=space.device)
        sig = torch.empty_like(eps)
        comp = torch.empty(n_el, dtype=torch.float64, device=space.device)
        _lib.check(_lib.lib().fes_elastic_fields(
            space.element_type.KIND, _lib.ptr(space._en_d), n_el, _lib.ptr(space._coords_d), _lib.ptr(ud),
            _lib.ptr(E_elem), 0.0 if m.per_element() else float(m.E), float(m.nu),
            space.element_type.default_quadrature_degree(), _lib.ptr(eps), _lib.ptr(sig), _lib.ptr(comp), _lib.stream()))
        return ElasticFields(eps, sig, comp)

    def stress_field(self, space, u, degree=None):
        """(n_elements, n_qp, d, d) in-plane Cauchy stress at every quadrature point of the space's rule
        (ref fem/forms.py:376-398; `degree` picks another tabulated rule, as geometry_at does).  Device tensor."""
        m = self.material
        n_el, sd = len(space.element_nodes), space.spatial_dim
        ud = u if isinstance(u, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64))
        ud = ud.to(device=space.device, dtype=torch.float64).contiguous().reshape(-1)
        if ud.numel() != space.n_nodes * sd:
            raise ValueError(f'expected {space.n_nodes * sd} displacement DOFs, got {ud.numel()}')
        degree = space.element_type.default_quadrature_degree() if degree is None else int(degree)
        n_qp = int(_lib.lib().fes_quadrature_size(space.element_type.KIND, degree))
        if n_qp == 0:
            raise NotImplementedError(f'no tabulated rule of degree >= {degree} for {space.element_type.__name__}')
        E_elem = _per_element(m.E, n_el, space.device, 'per-element modulus') if m.per_element() else None
        sig = torch.empty((n_el, n_qp, sd, sd), dtype=torch.float64, device=space.device)
        _lib.check(_lib.lib().fes_elastic_stress_field(
            space.element_type.KIND, _lib.ptr(space._en_d), n_el, _lib.ptr(space._coords_d), _lib.ptr(ud),
            _lib.ptr(E_elem), 0.0 if m.per_element() else float(m.E), float(m.nu), degree, _lib.ptr(sig), _lib.stream()))
        return sig


@dataclass(frozen=True, eq=False)
class ScaledForm(_Form):
    factor: float
    form: object

    def lower(self, space, boundary=False):
        low = self.form.lower(space, boundary)
        low.factor *= float(self.factor)
        return low


def _sample_field(field, space, n_components, degree):
    """The field at every quadrature point, (n_el, n_qp, n_components) on the device: constants are
    broadcast, callables are evaluated on the host over the flattened points exactly as the reference's
    `_sample_field` does (ref fem/forms.py:261-271)."""
    from .regions import evaluate_field
    points, _ = space.quadrature(degree)
    n_el, n_qp, sd = points.shape
    if callable(field):
        values = evaluate_field(field, points.reshape(n_el * n_qp, sd).cpu().numpy(), n_components)
    else:
        values = evaluate_field(field, np.zeros((1, sd)), n_components)
        values = np.broadcast_to(values, (n_el * n_qp, n_components))
    return torch.as_tensor(np.ascontiguousarray(values, dtype=np.float64)).to(space.device).reshape(n_el, n_qp, n_components)


class _MatrixForm:
    """Forms whose element matrices are computed by their own (unfused) kernel and scattered by the
    bit-exact gather of `fes_assemble_precomputed`."""
    precomputed = True

    def lower(self, space, boundary=False):
        return None


@dataclass(frozen=True, eq=False)
class DiffusionForm(_MatrixForm):
    """Variable-coefficient diffusion, kappa sampled at the quadrature points (ref fem/forms.py:273-293)."""
    coefficient: object
    quadrature_degree: int = 2

    def element_matrices(self, space, boundary=False):
        if boundary:
            raise NotImplementedError('DiffusionForm is a volume form')
        kappa = _sample_field(self.coefficient, space, 1, self.quadrature_degree)[..., 0].contiguous()
        n_el, k = len(space.element_nodes), space.element_type.N * space.n_components
        out = torch.empty((n_el, k, k), dtype=torch.float64, device=space.device)
        _lib.check(_lib.lib().fes_weighted_gradient_matrices(
            space.element_type.KIND, space.spatial_dim, space.n_components, self.quadrature_degree,
            _lib.ptr(space._en_d), n_el, _lib.ptr(space._coords_d), _lib.ptr(kappa), None, _lib.ptr(out), _lib.stream()))
        return out


@dataclass(frozen=True, eq=False)
class GeometricStiffnessForm(_MatrixForm):
    """Initial-stress stiffness from a per-element prestress (n_el, d, d) (ref fem/forms.py:401-441)."""
    prestress: object

    def element_matrices(self, space, boundary=False):
        if boundary:
            raise NotImplementedError('GeometricStiffnessForm is a volume form')
        d, n_el = space.spatial_dim, len(space.element_nodes)
        sigma = self.prestress if isinstance(self.prestress, torch.Tensor) else \
            torch.as_tensor(np.ascontiguousarray(self.prestress, dtype=np.float64))
        if tuple(sigma.shape) != (n_el, d, d):
            raise ValueError(f'prestress must be ({n_el}, {d}, {d}) for this mesh, got shape {tuple(sigma.shape)}')
        sigma = sigma.to(device=space.device, dtype=torch.float64).contiguous()
        k = space.element_type.N * space.n_components
        out = torch.empty((n_el, k, k), dtype=torch.float64, device=space.device)
        _lib.check(_lib.lib().fes_weighted_gradient_matrices(
            space.element_type.KIND, d, space.n_components, space.element_type.default_quadrature_degree(),
            _lib.ptr(space._en_d), n_el, _lib.ptr(space._coords_d), None, _lib.ptr(sigma), _lib.ptr(out), _lib.stream()))
        return out


@dataclass(frozen=True, eq=False)
class LinearForm:
    """L(v) = int f.v with f sampled at the quadrature points (ref fem/forms.py:296-315)."""
    field: object
    n_components: int = 1
    quadrature_degree: int = 2

    def element_vectors(self, space):
        """(n_el, N * n_components) element load vectors on the device, DOFs interleaved per node."""
        f = _sample_field(self.field, space, self.n_components, self.quadrature_degree)
        n_el, N = len(space.element_nodes), space.element_type.N
        out = torch.empty((n_el, N * self.n_components), dtype=torch.float64, device=space.device)
        _lib.check(_lib.lib().fes_element_loads(
            space.element_type.KIND, space.spatial_dim, self.quadrature_degree, _lib.ptr(space._en_d), n_el,
            _lib.ptr(space._coords_d), _lib.ptr(f), self.n_components, _lib.ptr(out), _lib.stream()))
        return out


@dataclass(frozen=True)
class ElasticFields:
    """Per-element strain / stress as (n_el, 3, 3) tensors and compliance (ref fem/forms.py:153-173)."""
    strain: object
    stress: object
    compliance: object


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
