"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python oracle/gen_golden.py

Imports janetyq/Finite-Element-Solver from /root/reference through `oracle/ref_shim.py`
and records its outputs on small deterministic inputs: assembled CSR operators, element
matrices, constrained solves, the cone filter, the OC update, a 5-iteration SIMP
trajectory, two time-integrator histories (theta method, Newmark), adjoint gradients and a
DesignOptimizer trajectory.  The fixtures are committed; this script is how they were made.  Inputs are
stored with the outputs so the tests never need the reference tree.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from fem.boundary import BCType, BoundaryConditions  # noqa: E402
from fem.design import DesignOptimizer, SIMPModel, optimality_criteria_update  # noqa: E402
from fem.sensitivity import (Compliance, DensityField, MeanStress, ModulusField, PointValue,  # noqa: E402
                             SensitivityAnalysis, SoftMaxStress)
from fem.elements import QuadraticTriangleElement  # noqa: E402
from fem.equations import LinearElastic, Poisson  # noqa: E402
from fem.forms import (DiffusionForm, GeometricStiffnessForm, LaplacianForm, LinearElasticForm,  # noqa: E402
                       LinearForm, MaskedMassForm, MassForm, PrecomputedForm, ScaledForm)
from fem.materials import LinearElasticMaterial  # noqa: E402
from fem.mesh.structured import create_box_mesh, create_rect_mesh  # noqa: E402
from fem.integrators import NewmarkMethod, ThetaMethod, wave_energy  # noqa: E402
from fem.problem import LinearProblem, heat, linear_elastic, wave  # noqa: E402
from fem.regions import everywhere, in_box, intersect, on_plane  # noqa: E402
from fem.solve import LinearSolve  # noqa: E402
from fem.solver import Solver  # noqa: E402
from fem.space import FunctionSpace  # noqa: E402
from fem.topology import TopologyOptimizer, calculate_smoothing_matrix  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def csr_fields(prefix, A):
    return {f'{prefix}_indptr': A.indptr.astype(np.int64), f'{prefix}_indices': A.indices.astype(np.int64),
            f'{prefix}_data': A.data.astype(np.float64)}


def space_fields(V):
    return {'vertices': V.mesh.vertices, 'elements': V.mesh.elements, 'boundary': V.mesh.boundary,
            'element_nodes': V.element_nodes, 'node_coords': V.node_coords,
            'boundary_nodes': V.boundary_nodes}


def assembly_case(name, mesh, element_type=None):
    d = mesh.spatial_dim
    out = {}
    V1 = FunctionSpace(mesh, element_type, n_components=1)
    Vd = FunctionSpace(mesh, element_type, n_components=d)
    out.update(space_fields(V1))
    n_el = len(V1.element_nodes)
    E_var = np.linspace(50.0, 200.0, n_el)
    mask = (np.arange(len(V1.boundary_nodes)) % 3 != 0)
    out['E_var'], out['facet_mask'] = E_var, mask
    out['volumes'] = V1.element_volumes
    forms = {
        'lap': (V1, LaplacianForm(), False),
        'mass1': (V1, MassForm(1), False),
        'bmass1': (V1, MassForm(1), True),
        'massd': (Vd, MassForm(d), False),
        'bmassd': (Vd, MassForm(d), True),
        'bmaskd': (Vd, MaskedMassForm(d, mask), True),
        'elast': (Vd, LinearElasticForm(LinearElasticMaterial(200.0, 0.3)), False),
        'elastvar': (Vd, LinearElasticForm(LinearElasticMaterial(E_var, 0.3)), False),
        'lapscaled': (V1, ScaledForm(2.25, LaplacianForm()), False),
    }
    for key, (V, form, bnd) in forms.items():
        A = V.assemble(form, boundary=bnd)
        out.update(csr_fields(key, A))
        geo = V.boundary_geometry if bnd else V.geometry
        out[f'{key}_ke'] = form.element_matrices(geo)
    K0 = LinearElasticForm(LinearElasticMaterial(200.0, 0.3)).element_matrices(Vd.geometry)
    rho = np.linspace(0.2, 1.0, n_el)
    out['rho'] = rho
    out.update(csr_fields('precomp', Vd.assemble(PrecomputedForm(rho[:, None, None] ** 3 * K0))))
    np.savez_compressed(os.path.join(OUT, f'assembly_{name}.npz'), **out)
    print(name, 'n_el', n_el, 'nnz elast', out['elast_data'].size)


def solve_cases():
    out = {}
    # 2D cantilever (ref: tests/test_backends.py:73-84)
    mesh = create_rect_mesh([[0, 0], [2, 1]], (9, 5))
    bc = BoundaryConditions()
    bc.add(BCType.DIRICHLET, on_plane(0, 0.0), [0, 0])
    bc.add(BCType.NEUMANN, on_plane(0, 2.0), [50, -10])
    prob = linear_elastic(mesh, LinearElasticMaterial(200.0, 0.3), bc)
    free, fixed, vals = prob.constraints
    out.update({'c2d_vertices': mesh.vertices, 'c2d_elements': mesh.elements, 'c2d_boundary': mesh.boundary,
                'c2d_free': free, 'c2d_fixed': fixed, 'c2d_vals': vals, 'c2d_load': prob.load,
                'c2d_u': LinearSolve().solve(prob)})
    # 3D cantilever (ref: tests/test_backends.py:109-113)
    mesh = create_box_mesh([[0, 0, 0], [1, 1, 1]], (4, 4, 4))
    bc = BoundaryConditions()
    bc.add(BCType.DIRICHLET, on_plane(0, 0.0), [0, 0, 0])
    bc.add(BCType.NEUMANN, on_plane(0, 1.0), [0, -5, 0])
    prob = linear_elastic(mesh, LinearElasticMaterial(200.0, 0.3), bc)
    free, fixed, vals = prob.constraints
    out.update({'c3d_vertices': mesh.vertices, 'c3d_elements': mesh.elements, 'c3d_boundary': mesh.boundary,
                'c3d_free': free, 'c3d_fixed': fixed, 'c3d_vals': vals, 'c3d_load': prob.load,
                'c3d_u': LinearSolve().solve(prob)})
    # Poisson, nonzero Dirichlet data + source + Robin on one side
    mesh = create_rect_mesh([[0, 0], [1, 1]], (7, 6))
    bc = BoundaryConditions()
    bc.add(BCType.DIRICHLET, on_plane(0, 0.0), 1.5)
    bc.add(BCType.DIRICHLET, on_plane(0, 1.0), -0.5)
    bc.add_robin(on_plane(1, 1.0), 2.0, 0.75)
    solver = Solver(mesh, Poisson(source=lambda p: [1.0 + p[0] * p[1]]), bc)
    sol = solver.solve()
    V = FunctionSpace(mesh)
    prob = LinearProblem(V, LaplacianForm(), lambda p: [1.0 + p[0] * p[1]], bc)
    free, fixed, vals = prob.constraints
    A = prob.tangent(None)
    out.update({'poi_vertices': mesh.vertices, 'poi_elements': mesh.elements, 'poi_boundary': mesh.boundary,
                'poi_free': free, 'poi_fixed': fixed, 'poi_vals': vals, 'poi_load': prob.load,
                'poi_u': sol.u, 'poi_A_dense': A.toarray()})
    np.savez_compressed(os.path.join(OUT, 'solve.npz'), **out)
    print('solve cases written')


def simp_case():
    # MBB beam of examples/solver_demos.py:1271-1288 on a small mesh
    w, h = 3.0, 1.0
    mesh = create_rect_mesh([[0, 0], [w, h]], (25, 9))
    bc = BoundaryConditions()
    bottom, top = on_plane(1, 0.0), on_plane(1, h)
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
    bc.add(BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
    r = 2.5 * (w / 24)
    topo = TopologyOptimizer(mesh, LinearElastic(200.0, 0.4), bc, iters=5, volume_frac=0.5,
                             smoothing_radius=r, penalty=3.0)
    sens_series = []
    orig = topo._compliance_sensitivity

    def spy(solution):
        s = orig(solution)
        sens_series.append(s.copy())
        return s
    topo._compliance_sensitivity = spy
    hist = topo.solve()
    free, fixed, vals = topo._problem.constraints
    S = topo.smoothing_matrix
    out = {'vertices': mesh.vertices, 'elements': mesh.elements, 'boundary': mesh.boundary,
           'free': free, 'fixed': fixed, 'vals': vals, 'load': topo._problem.load,
           'radius': r, 'E0': 200.0, 'nu': 0.4, 'penalty': 3.0, 'volume_frac': 0.5,
           'rho': np.array(hist.rho), 'u': np.array(hist.u), 'compliance': np.array(hist.compliance),
           'von_mises': np.array(hist.von_mises), 'sens': np.array(sens_series), 'rho_final': topo.rho,
           'K0': topo._solid_stiffness, 'volumes': topo.space.element_volumes}
    out.update(csr_fields('filter', S))
    np.savez_compressed(os.path.join(OUT, 'simp.npz'), **out)
    print('simp written; compliance', [float(c.sum()) for c in hist.compliance])

    # the filter alone on a unit-square mesh (ref: tests/test_topology_optimization.py:129-174)
    mesh = create_rect_mesh([[0, 0], [1, 1]], (12, 12))
    S = calculate_smoothing_matrix(mesh, 0.15)
    box = create_box_mesh([[0, 0, 0], [1, 1, 1]], (4, 4, 4))
    S3 = calculate_smoothing_matrix(box, 0.3)
    f = {'sq_vertices': mesh.vertices, 'sq_elements': mesh.elements, 'sq_r': 0.15,
         'box_vertices': box.vertices, 'box_elements': box.elements, 'box_r': 0.3}
    f.update(csr_fields('sq', S))
    f.update(csr_fields('box', S3))
    # OC update on random data
    rng = np.random.default_rng(7)
    rho = rng.uniform(0.05, 1.0, 500)
    sens = rng.uniform(0.0, 3.0, 500)
    vol = rng.uniform(0.5, 1.5, 500)
    f.update({'oc_rho': rho, 'oc_sens': sens, 'oc_vol': vol,
              'oc_out_04': optimality_criteria_update(rho, sens, vol, 0.4, move=0.1),
              'oc_out_06': optimality_criteria_update(rho, sens, vol, 0.6, move=0.2)})
    np.savez_compressed(os.path.join(OUT, 'filter_oc.npz'), **f)
    print('filter/oc written')


def simp100_case():
    """Config 4 as BASELINE.json states it -- 100 design iterations -- on a reduced MBB mesh (61x21 nodes,
    2400 triangles, r = 2.5 cells): the reference's TopologyOptimizer.solve() trajectory (DirectBackend)."""
    w, h = 3.0, 1.0
    res = (61, 21)
    mesh = create_rect_mesh([[0, 0], [w, h]], res)
    bc = BoundaryConditions()
    bottom, top = on_plane(1, 0.0), on_plane(1, h)
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
    bc.add(BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
    bc.add(BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
    r = 2.5 * (w / (res[0] - 1))
    topo = TopologyOptimizer(mesh, LinearElastic(200.0, 0.4), bc, iters=100, volume_frac=0.5,
                             smoothing_radius=r, penalty=3.0)
    hist = topo.solve()
    rho = np.array(hist.rho + [topo.rho])                 # rho[k] = density going INTO iteration k; rho[100] = final
    comp = np.array([float(np.sum(c)) for c in hist.compliance])
    keep = np.array([0, 1, 2, 5, 10, 25, 50, 75, 98, 99])
    vol = topo.space.element_volumes
    free, fixed, vals = topo._problem.constraints
    out = {'free': free, 'fixed': fixed, 'load': topo._problem.load, 'resolution': np.array(res), 'width_height': np.array([w, h]), 'radius': r, 'E0': 200.0, 'nu': 0.4,
           'penalty': 3.0, 'volume_frac': 0.5, 'total_compliance': comp, 'kept_iterations': keep,
           'rho_kept': rho[keep], 'rho_next_kept': rho[keep + 1], 'compliance_kept': np.array(hist.compliance)[keep],
           'volume_fraction': np.array([float((vol * r_).sum() / vol.sum()) for r_ in rho])}
    np.savez_compressed(os.path.join(OUT, 'simp100.npz'), **out)
    print('simp100 written; compliance', comp[0], '->', comp[-1])


def transient_cases():
    """ThetaMethod on a heat problem (source + nonzero Dirichlet data + Neumann flux) and NewmarkMethod on
    a wave problem, both with the reference's DirectBackend (ref fem/integrators.py)."""
    out = {}
    mesh = create_rect_mesh([[0, 0], [2, 1]], (17, 13))
    bc = BoundaryConditions()
    bc.add(BCType.DIRICHLET, on_plane(0, 0.0), [1.5])
    bc.add(BCType.NEUMANN, on_plane(0, 2.0), [-0.75])
    prob = heat(mesh, source=lambda p: np.array([np.sin(3 * p[0]) + p[1]]), bc=bc)
    u0 = 1.5 * np.ones(prob.space.n_dofs)
    for tag, theta in (('cn', 0.5), ('be', 1.0)):
        sol = ThetaMethod(dt=0.02, steps=6, theta=theta).run(prob, u0)
        out[f'heat_{tag}_t'] = np.asarray(sol.t)
        out[f'heat_{tag}_u'] = np.asarray(sol.u_history if hasattr(sol, 'u_history') else sol.u)
    free, fixed, vals = prob.constraints
    out.update(heat_vertices=mesh.vertices, heat_elements=mesh.elements, heat_boundary=mesh.boundary, heat_u0=u0,
               heat_free=free, heat_fixed=fixed, heat_vals=vals, heat_load=prob.load)
    out.update(csr_fields('heat_K', prob.tangent(None)))
    out.update(csr_fields('heat_M', prob.space.mass_matrix))

    bcw = BoundaryConditions()
    bcw.add(BCType.DIRICHLET, on_plane(0, 0.0), [0.0])
    bcw.add(BCType.DIRICHLET, on_plane(0, 2.0), [0.0])
    probw = wave(mesh, 1.7, bcw)
    x, y = mesh.vertices[:, 0], mesh.vertices[:, 1]
    w0 = np.sin(np.pi * x / 2.0) * (1 + 0.3 * np.cos(np.pi * y))
    v0 = 0.2 * np.sin(np.pi * x)
    sol = NewmarkMethod(dt=0.01, steps=6).run(probw, w0, v0)
    hist_u = np.asarray(sol.u_history if hasattr(sol, 'u_history') else sol.u)
    hist_v = np.asarray(sol.dudt_history if hasattr(sol, 'dudt_history') else sol.dudt)
    freew, fixedw, valsw = probw.constraints
    out.update(wave_t=np.asarray(sol.t), wave_u=hist_u, wave_v=hist_v, wave_u0=w0, wave_v0=v0, wave_free=freew,
               wave_fixed=fixedw, wave_vals=valsw, wave_load=probw.load,
               wave_energy=np.array([wave_energy(probw, hist_u[i], hist_v[i]) for i in range(len(hist_u))]))
    out.update(csr_fields('wave_K', probw.tangent(None)))
    np.savez_compressed(os.path.join(OUT, 'transient.npz'), **out)


def adjoint_cases():
    """SensitivityAnalysis gradients (Compliance / PointValue x DensityField / ModulusField) and a
    DesignOptimizer trajectory on a small cantilever (ref fem/sensitivity.py, fem/design.py)."""
    out = {}
    mesh = create_rect_mesh([[0, 0], [1, 1]], (9, 9))
    bc = BoundaryConditions()
    bc.add(BCType.DIRICHLET, on_plane(0, 0.0), [0.0, 0.0])
    bc.add(BCType.NEUMANN, on_plane(0, 1.0), [0.0, -1.0])
    space = FunctionSpace(mesh, n_components=2)
    n_el = len(mesh.elements)
    rho = 0.3 + 0.6 * (np.arange(n_el) % 7) / 6.0
    E_el = 0.5 + (np.arange(n_el) % 5) / 4.0
    nu, penalty = 0.3, 3.0
    model = SIMPModel(space, base_E=1.0, nu=nu, bc=bc, penalty=penalty)
    prob = model.problem(rho)
    an = SensitivityAnalysis(prob)
    u = an.solve_forward()
    tip = int(2 * np.argmax(mesh.vertices[:, 0] + mesh.vertices[:, 1]) + 1)
    out.update(vertices=mesh.vertices, elements=mesh.elements, boundary=mesh.boundary, rho=rho, E_el=E_el, tip_dof=tip,
               u_rho=u, load=prob.load, free=prob.constraints[0], fixed=prob.constraints[1], vals=prob.constraints[2],
               grad_compliance_density=an.gradient(Compliance(), model.parameterization(rho), u),
               grad_point_density=an.gradient(PointValue(tip), model.parameterization(rho), u),
               lam_point=an.adjoint(PointValue(tip), u),
               compliance_value=Compliance().value(prob, u), point_value=PointValue(tip).value(prob, u))
    probE = LinearProblem(space, LinearElasticForm(LinearElasticMaterial(E_el, nu)), None, bc)
    anE = SensitivityAnalysis(probE)
    uE = anE.solve_forward()
    out.update(u_E=uE, grad_point_modulus=anE.gradient(PointValue(tip), ModulusField.create(space, E_el, nu), uE),
               grad_compliance_modulus=anE.gradient(Compliance(), ModulusField.create(space, E_el, nu), uE))
    filt = calculate_smoothing_matrix(mesh, 0.2)
    modelf = SIMPModel(space, base_E=1.0, nu=nu, bc=bc, penalty=penalty, sensitivity_filter=filt)
    design = DesignOptimizer(modelf, Compliance(), volume_frac=0.5, iters=5, move=0.1)
    hist = design.solve()
    out.update(design_rho=np.asarray(hist.rho), design_u=np.asarray(hist.u), design_objective=np.asarray(hist.objective),
               design_rho_final=design.rho, filter_radius=0.2)
    np.savez_compressed(os.path.join(OUT, 'adjoint.npz'), **out)


def stress_qoi_cases():
    """MeanStress / SoftMaxStress values, adjoint loads and density gradients on a jittered cantilever,
    P1 and P2 (ref fem/sensitivity.py:108-276)."""
    out = {}
    for name, etype in (('p1', None), ('p2', QuadraticTriangleElement)):
        mesh = _jitter(create_rect_mesh([[0, 0], [1, 1]], (9, 9) if etype is None else (6, 6)), 9, 0.02)
        bc = BoundaryConditions()
        bc.add(BCType.DIRICHLET, on_plane(0, 0.0), [0.0, 0.0])
        bc.add(BCType.NEUMANN, on_plane(0, 1.0), [0.0, -1.0])
        space = FunctionSpace(mesh, etype, n_components=2)
        n_el = len(mesh.elements)
        rho = 0.3 + 0.6 * (np.arange(n_el) % 7) / 6.0
        E_el = 0.5 + (np.arange(n_el) % 5) / 4.0
        nu, penalty = 0.3, 3.0
        model = SIMPModel(space, base_E=1.0, nu=nu, bc=bc, penalty=penalty)
        prob = model.problem(rho)
        an = SensitivityAnalysis(prob)
        u = an.solve_forward()
        region = (np.arange(n_el) % 3 != 1)
        mat_u, mat_v = LinearElasticMaterial(1.0, nu), LinearElasticMaterial(E_el, nu)
        out.update({f'{name}_vertices': mesh.vertices, f'{name}_elements': mesh.elements, f'{name}_boundary': mesh.boundary,
                    f'{name}_rho': rho, f'{name}_E_el': E_el, f'{name}_u': u, f'{name}_region': region})
        for tag, qoi in (('mean', MeanStress(space, mat_u)), ('mean_region_var', MeanStress(space, mat_v, region)),
                         ('soft', SoftMaxStress(space, mat_u, None, 8.0)),
                         ('soft_region_var', SoftMaxStress(space, mat_v, region, 6.0))):
            out[f'{name}_{tag}_value'] = qoi.value(prob, u)
            out[f'{name}_{tag}_dJ_du'] = qoi.dJ_du(prob, u)
            out[f'{name}_{tag}_grad'] = an.gradient(qoi, model.parameterization(rho), u)
        out[f'{name}_vm_u'] = MeanStress(space, mat_u)._stress().von_mises(u)
        out[f'{name}_vm_v'] = MeanStress(space, mat_v)._stress().von_mises(u)
    np.savez_compressed(os.path.join(OUT, 'stress_qoi.npz'), **out)
    print('stress qoi cases written')


def io_cases():
    """Files written by the reference's own fem.io (ref fem/io.py:52-147): a mesh JSON, an ElasticSolution
    archive (P2, so the element type is stored) and a WaveSolution archive."""
    from fem.io import save_mesh, save_solution
    from fem.solution import ElasticSolution
    mesh = _jitter(create_rect_mesh([[0, 0], [1.5, 1]], (4, 3)), 7, 0.03)
    save_mesh(mesh, os.path.join(OUT, 'io_mesh.json'))
    space = FunctionSpace(mesh, QuadraticTriangleElement, n_components=2)
    u = np.random.default_rng(3).normal(size=space.n_dofs) * 0.01
    sol = ElasticSolution.from_solve(space, u, LinearElasticForm(LinearElasticMaterial(200.0, 0.3)))
    save_solution(sol, os.path.join(OUT, 'io_elastic_p2.npz'))
    bc = BoundaryConditions()
    bc.add(BCType.DIRICHLET, on_plane(0, 0.0), 0.0)
    prob = wave(create_rect_mesh([[0, 0], [1, 1]], (4, 4)), 1.5, bc)
    x = prob.space.node_coords
    hist = NewmarkMethod(dt=0.05, steps=3).run(prob, np.sin(np.pi * x[:, 0]) * x[:, 1], np.zeros(len(x)))
    save_solution(hist, os.path.join(OUT, 'io_wave.npz'))
    from fem.solution import ModalSolution, ScalarFieldSolution
    V1 = FunctionSpace(mesh, n_components=1)
    u1 = np.sin(mesh.vertices[:, 0]) * mesh.vertices[:, 1]
    save_solution(ScalarFieldSolution.from_solve(V1, u1), os.path.join(OUT, 'io_scalar.npz'))
    rng = np.random.default_rng(4)
    save_solution(ModalSolution(mesh, 2, np.array([1.5, 2.5, 7.0]), rng.normal(size=(3, 2 * len(mesh.vertices)))),
                  os.path.join(OUT, 'io_modal.npz'))
    print('io cases written')


def _jitter(mesh, seed, amp):
    """Interior vertices moved by a seeded random offset, so no two elements are congruent."""
    rng = np.random.default_rng(seed)
    v = mesh.vertices.copy()
    interior = np.setdiff1d(np.arange(len(v)), mesh.boundary_idxs)
    v[interior] += rng.uniform(-amp, amp, (len(interior), v.shape[1]))
    return type(mesh)(v, mesh.elements, mesh.boundary)


def fields_cases():
    """LinearForm loads, DiffusionForm / GeometricStiffnessForm operators, nodal recovery and
    LinearElasticForm.derived_fields on jittered P1-tri, P2-tri and P1-tet meshes
    (ref fem/forms.py:273-441, fem/space.py:545-708)."""
    out = {}
    kappa = lambda p: [1.0 + p[0] ** 2 + 0.5 * p[1]]                                  # noqa: E731
    cases = {
        'tri': (_jitter(create_rect_mesh([[0, 0], [2, 1]], (6, 5)), 3, 0.04), None),
        'p2tri': (_jitter(create_rect_mesh([[0, 0], [1.5, 1]], (5, 4)), 4, 0.04), QuadraticTriangleElement),
        'tet': (_jitter(create_box_mesh([[0, 0, 0], [1, 2, 1.5]], (3, 4, 3)), 5, 0.05), None),
    }
    for name, (mesh, etype) in cases.items():
        d = mesh.spatial_dim
        rng = np.random.default_rng(11 + d + (etype is not None))
        V1 = FunctionSpace(mesh, etype, n_components=1)
        Vd = FunctionSpace(mesh, etype, n_components=d)
        n_el = len(V1.element_nodes)
        pre = {f'{name}_{k}': v for k, v in space_fields(V1).items()}
        out.update(pre)
        # loads: a scalar field (degree 2) and a vector field (the highest tabulated degree)
        f1 = (lambda p: [np.sin(p[0]) + p[1] ** 2]) if d == 2 else (lambda p: [np.sin(p[0]) + p[1] * p[2]])
        fd = (lambda p: [p[0] * p[1], 1.0 + p[0]]) if d == 2 else (lambda p: [p[0] * p[1], 1.0 + p[2], p[1] - p[0]])
        hi = 4 if d == 2 else 2
        out[f'{name}_load1'] = V1.assemble_load(LinearForm(f1, 1, 2))
        out[f'{name}_loadd'] = Vd.assemble_load(LinearForm(fd, d, hi))
        out[f'{name}_loadd_degree'] = hi
        out[f'{name}_points2'] = V1.geometry_at(2).points
        out[f'{name}_wdet2'] = V1.geometry_at(2).weight_detJ
        # operators
        out.update(csr_fields(f'{name}_diff', V1.assemble(DiffusionForm(kappa if d == 2 else (lambda p: [1.0 + p[0] ** 2 + 0.5 * p[2]])))))
        sym = rng.normal(size=(n_el, d, d))
        prestress = sym + np.transpose(sym, (0, 2, 1))
        out[f'{name}_prestress'] = prestress
        out.update(csr_fields(f'{name}_geom', Vd.assemble(GeometricStiffnessForm(prestress))))
        # nodal recovery
        vals = rng.normal(size=n_el)
        tens = rng.normal(size=(n_el, 2, 2))
        out[f'{name}_vals'], out[f'{name}_tens'] = vals, tens
        out[f'{name}_avg'] = V1.recover_nodal(vals)
        out[f'{name}_avg_tens'] = V1.recover_nodal(tens)
        out[f'{name}_l2'] = V1.recover_nodal(vals, method='l2')
        out[f'{name}_l2_tens'] = Vd.recover_nodal(tens, method='l2')
        # derived fields of a random displacement, uniform and per-element modulus
        u = rng.normal(size=Vd.n_dofs) * 0.01
        out[f'{name}_u'] = u
        ue = u.reshape(-1, d)[Vd.element_nodes]
        E_var = np.linspace(50.0, 200.0, n_el)
        out[f'{name}_E_var'] = E_var
        for tag, E in (('uni', 200.0), ('var', E_var)):
            fields = LinearElasticForm(LinearElasticMaterial(E, 0.3)).derived_fields(Vd.geometry, ue)
            out[f'{name}_strain_{tag}'], out[f'{name}_stress_{tag}'] = fields.strain, fields.stress
            out[f'{name}_compliance_{tag}'] = fields.compliance
            # stress at every quadrature point (ref fem/forms.py:376-398): the space's own rule, and a richer one
            form = LinearElasticForm(LinearElasticMaterial(E, 0.3))
            out[f'{name}_stressfield_{tag}'] = form.stress_field(Vd.geometry, ue)
            out[f'{name}_stressfield2_{tag}'] = form.stress_field(Vd.geometry_at(2), ue)
    np.savez_compressed(os.path.join(OUT, 'fields.npz'), **out)
    print('fields cases written')


if __name__ == '__main__':
    assembly_case('tri', create_rect_mesh([[0, 0], [2, 1]], (6, 5)))
    assembly_case('tet', create_box_mesh([[0, 0, 0], [1, 2, 1.5]], (3, 4, 3)))
    assembly_case('p2tri', create_rect_mesh([[0, 0], [1.5, 1]], (5, 4)), QuadraticTriangleElement)
    solve_cases()
    simp_case()
    simp100_case()
    transient_cases()
    adjoint_cases()
    fields_cases()
    stress_qoi_cases()
    io_cases()
