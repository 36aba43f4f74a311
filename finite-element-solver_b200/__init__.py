"""fes_b200: the B200-native hot path of janetyq/Finite-Element-Solver.

Python host code with the reference's API (FunctionSpace.assemble, DiscreteSystem/Backend,
TopologyOptimizer) driving hand-written sm_100a kernels through the C ABI of libfes_b200.so.
Importable as `fes_b200` (the directory name carries a hyphen).
"""
from . import _lib  # noqa: F401
from .backends import B200Backend, JacobiPCGBackend  # noqa: F401
from .boundary import BCType, BoundaryConditions  # noqa: F401
from .csr import DeviceCSR  # noqa: F401
from .design import DesignHistory, DesignOptimizer, SIMPModel, optimality_criteria_update  # noqa: F401
from .elements import (LinearLineElement, LinearTetrahedralElement, LinearTriangleElement,  # noqa: F401
                       QuadraticLineElement, QuadraticTriangleElement)
from .equations import LinearElastic, Poisson  # noqa: F401
from .forms import (DiffusionForm, ElasticFields, GeometricStiffnessForm, LaplacianForm,  # noqa: F401
                    LinearElasticForm, LinearForm, MaskedMassForm, MassForm, PrecomputedForm, ScaledForm)
from .materials import LinearElasticMaterial  # noqa: F401
from .mesh import Mesh, create_box_mesh, create_rect_mesh  # noqa: F401
from .integrators import NewmarkMethod, ThetaMethod, wave_energy  # noqa: F401
from .io import load_mesh, load_solution, save_mesh, save_solution  # noqa: F401
from .solution import (BucklingSolution, ElasticSolution, FieldSolution, ModalSolution,  # noqa: F401
                       ScalarFieldSolution, Solution, TransientSolution, WaveSolution)
from .problem import LinearProblem, heat, linear_elastic, poisson, projection, wave  # noqa: F401
from .sensitivity import (Compliance, DensityField, MeanStress, ModulusField, PointValue,  # noqa: F401
                          SensitivityAnalysis, SoftMaxStress)
from .space import FunctionSpace  # noqa: F401
from .system import DiscreteSystem  # noqa: F401
from .topology import TopologyOptimizer, calculate_smoothing_matrix  # noqa: F401

__all__ = [n for n in dir() if not n.startswith('_')]
