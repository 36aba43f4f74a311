"""Persistence for meshes and solutions in the reference's on-disk formats (ref fem/io.py:52-147): meshes as
JSON {vertices, elements, boundary}; solutions as one .npz holding the value arrays (`value.<field>`)
beside the mesh geometry, the component count (`__dim__`), the solution class name and the element type
name.  No pickle: `np.load` keeps `allow_pickle=False`, and a ragged field is refused at save time.
Archives written by the reference load here and vice versa.
"""
import dataclasses
import json

import numpy as np

_SOLUTION_HEADER = ('mesh', 'n_components', 'element_type')
_VALUE_PREFIX = 'value.'
_MESH_CLASS, _ELEMENT_TYPE = '__mesh_class__', '__element_type__'
_MESH_VERTICES, _MESH_ELEMENTS, _MESH_BOUNDARY = '__mesh_vertices__', '__mesh_elements__', '__mesh_boundary__'
_N_COMPONENTS, _SOLUTION_CLASS = '__dim__', '__solution_class__'


def save_mesh(mesh, path='test_mesh.json'):
    with open(path, 'w') as f:
        json.dump({'vertices': np.asarray(mesh.vertices).tolist(), 'elements': np.asarray(mesh.elements).tolist(),
                   'boundary': np.asarray(mesh.boundary).tolist()}, f)


def load_mesh(path='test_mesh.json'):
    from .mesh import Mesh
    with open(path, 'r') as f:
        data = json.load(f)
    return Mesh(data['vertices'], data['elements'], data['boundary'])


def _host(value):
    if isinstance(value, (list, tuple)):
        value = [v.detach().cpu().numpy() if hasattr(v, 'detach') else v for v in value]
    elif hasattr(value, 'detach'):
        value = value.detach().cpu().numpy()
    try:
        return np.asarray(value)
    except ValueError:                      # ragged: numpy >= 1.24 refuses instead of building an object array
        return np.empty(0, dtype=object)


def save_solution(solution, path='solution.npz'):
    mesh = solution.mesh
    arrays = {_MESH_VERTICES: np.asarray(mesh.vertices), _MESH_ELEMENTS: np.asarray(mesh.elements),
              _MESH_BOUNDARY: np.asarray(mesh.boundary), _MESH_CLASS: np.array('Mesh'),
              _N_COMPONENTS: np.asarray(solution.n_components), _SOLUTION_CLASS: np.array(type(solution).__name__)}
    et = getattr(solution, 'element_type', None)
    arrays[_ELEMENT_TYPE] = np.array(et.__name__ if et is not None else '')
    for f in dataclasses.fields(solution):
        if f.name in _SOLUTION_HEADER:
            continue
        value = _host(getattr(solution, f.name))
        if value.dtype == object:
            raise ValueError(f'solution field {f.name!r} is ragged and cannot be saved without pickle')
        arrays[_VALUE_PREFIX + f.name] = value
    with open(path, 'wb') as f:             # a handle, so numpy does not append its own .npz
        np.savez(f, **arrays)


def load_solution(path='solution.npz'):
    from . import elements as elements_module
    from . import solution as solution_module
    from .mesh import Mesh
    with np.load(path) as data:
        mesh_class = str(data[_MESH_CLASS])
        if mesh_class not in ('Mesh', 'FEMesh'):
            raise ValueError(f'Unknown mesh class in saved solution: {mesh_class}')
        mesh = Mesh(data[_MESH_VERTICES], data[_MESH_ELEMENTS], data[_MESH_BOUNDARY])
        n_components = int(data[_N_COMPONENTS])
        name = str(data[_SOLUTION_CLASS])
        cls = getattr(solution_module, name, None)
        if cls is None or not (isinstance(cls, type) and issubclass(cls, solution_module.Solution)):
            raise ValueError(f'Unknown solution class in saved solution: {name}')
        type_name = str(data[_ELEMENT_TYPE]) if _ELEMENT_TYPE in data else ''
        element_type = getattr(elements_module, type_name, None) if type_name else None
        if type_name and not (isinstance(element_type, type) and issubclass(element_type, elements_module.Element)):
            raise ValueError(f'Unknown element type in saved solution: {type_name}')
        fields = {f.name: data[_VALUE_PREFIX + f.name] for f in dataclasses.fields(cls)
                  if f.name not in _SOLUTION_HEADER}
        return cls(mesh, n_components, element_type=element_type, **fields)
