"""Isotropic linear-elastic material (ref fem/materials.py:24-28,81-128): E scalar or
per-element, nu uniform, plane strain in 2D.  The Voigt D matrix is never formed; the device
kernels contract the closed form (csrc/elements.cuh)."""
from dataclasses import dataclass

import numpy as np


def Enu_to_Lame(E, nu):
    return E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))


@dataclass(frozen=True)
class LinearElasticMaterial:
    E: object   # float | (n_elements,) array (numpy or torch)
    nu: float

    def per_element(self):
        return not np.isscalar(self.E)
