"""Element type descriptors: the names and class attributes of the reference's element
hierarchy (ref fem/elements.py:31-126,249-413), reduced to what the device kernels dispatch on.
The shape functions, quadrature tables and affine geometry themselves live in csrc/elements.cuh."""
from . import _lib


class Element:
    N = None              # nodes per element
    SHAPE_DEGREE = None
    SUB_TYPE = None       # element type of a boundary facet
    GEOMETRY_DEGREE = 1
    KIND = None           # FES_* code of include/fes_b200.h
    REF_DIM = None

    @classmethod
    def reference_dim(cls):
        return cls.REF_DIM

    @classmethod
    def n_corners(cls):
        return cls.REF_DIM + 1

    @classmethod
    def default_quadrature_degree(cls):
        return max(1, 2 * (cls.SHAPE_DEGREE - 1))


class LinearLineElement(Element):
    N, SHAPE_DEGREE, SUB_TYPE, KIND, REF_DIM = 2, 1, None, _lib.P1_LINE, 1


class LinearTriangleElement(Element):
    N, SHAPE_DEGREE, SUB_TYPE, KIND, REF_DIM = 3, 1, LinearLineElement, _lib.P1_TRI, 2


class LinearTetrahedralElement(Element):
    N, SHAPE_DEGREE, SUB_TYPE, KIND, REF_DIM = 4, 1, LinearTriangleElement, _lib.P1_TET, 3


class QuadraticLineElement(Element):
    N, SHAPE_DEGREE, SUB_TYPE, KIND, REF_DIM = 3, 2, None, _lib.P2_LINE, 1


class QuadraticTriangleElement(Element):
    N, SHAPE_DEGREE, SUB_TYPE, KIND, REF_DIM = 6, 2, QuadraticLineElement, _lib.P2_TRI, 2


_SIMPLEX = {2: LinearLineElement, 3: LinearTriangleElement, 4: LinearTetrahedralElement}


def element_type_for(mesh):
    n = mesh.elements.shape[1]
    if n not in _SIMPLEX:
        raise NotImplementedError(f'no linear element for {n}-node elements')
    return _SIMPLEX[n]
