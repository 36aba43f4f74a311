"""LinearProblem: constraints, load and a lazily assembled operator (ref fem/problem.py:86-179).

Same composition as the reference -- load = M f + sum of per-region masked boundary masses
times their tractions + Robin loads; tangent = assemble(operator) [+ Robin operator] -- with
every matrix assembled by the fused CUDA kernel and every matrix-vector product done on the
device.  `tangent()` returns a DeviceCSR.
"""
import copy

import numpy as np
import torch

from .boundary import BoundaryConditions
from .forms import LaplacianForm, LinearElasticForm, MaskedMassForm, MassForm, ScaledForm
from .regions import evaluate_field
from .space import FunctionSpace


class LinearProblem:
    def __init__(self, space, operator, source=None, bc=None):
        self.space, self.operator = space, operator
        bc = bc if bc is not None else BoundaryConditions()
        self._resolved = bc.resolve(space.nodes, space.n_components)
        dev = space.device
        load = torch.zeros(space.n_dofs, dtype=torch.float64, device=dev)
        self._robin_operator = None
        for robin in self._resolved.robin:
            bm = space.assemble_device(MaskedMassForm(space.n_components, robin.facet_mask), boundary=True)
            term = bm.scaled(robin.kappa)
            self._robin_operator = term if self._robin_operator is None else \
                self._robin_operator.add_subpattern(term)
            load += bm.matvec(torch.as_tensor(robin.g.ravel(), device=dev))
        for neumann in self._resolved.neumann:
            bm = space.assemble_device(MaskedMassForm(space.n_components, neumann.facet_mask), boundary=True)
            load += bm.matvec(torch.as_tensor(neumann.traction.ravel(), device=dev))
        # the reference always forms M f, even for source=None; a zero source adds 0
        if source is not None:
            f = evaluate_field(source, space.node_coords, space.n_components)
            load += space.mass_matrix_d.matvec(torch.as_tensor(f.ravel(), device=dev))
        self._b_d = load
        self._A = None

    def with_operator(self, operator):
        derived = copy.copy(self)
        derived.operator = operator
        derived._A = None
        return derived

    @property
    def constraints(self):
        r = self._resolved
        return (r.free_idxs, r.fixed_idxs, r.fixed_values)

    @property
    def load_d(self):
        return self._b_d

    @property
    def load(self):
        return self._b_d.cpu().numpy()

    def tangent(self, u=None):
        if self._A is None:
            A = self.space.assemble_device(self.operator)
            self._A = A if self._robin_operator is None else A.add_subpattern(self._robin_operator)
        return self._A

    def residual(self, u):
        return self.tangent().matvec(np.asarray(u, dtype=np.float64)) - self.load


def poisson(mesh, source=None, bc=None):
    return LinearProblem(FunctionSpace(mesh, n_components=1), LaplacianForm(), source, bc)


def linear_elastic(mesh, material, bc=None, source=None):
    space = FunctionSpace(mesh, n_components=mesh.spatial_dim)
    return LinearProblem(space, LinearElasticForm(material), source, bc)


def projection(mesh, source, bc=None):
    """L2 projection of `source` onto the P1 space, M u = M f (ref fem/problem.py:236-239)."""
    space = FunctionSpace(mesh, n_components=1)
    return LinearProblem(space, MassForm(space.n_components), source, bc)


def heat(mesh, source=None, bc=None):
    """Transient heat: Poisson's operator, to be stepped by a ThetaMethod (ref fem/problem.py:274-282)."""
    return LinearProblem(FunctionSpace(mesh, n_components=1), LaplacianForm(), source, bc)


def wave(mesh, c, bc=None, source=None):
    """Transient wave with speed c: the Laplacian scaled by c^2, to be Newmark-stepped (ref fem/problem.py:285-294)."""
    return LinearProblem(FunctionSpace(mesh, n_components=1), ScaledForm(c ** 2, LaplacianForm()), source, bc)
