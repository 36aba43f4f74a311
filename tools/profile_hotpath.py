"""Small driver for ncu: a few fused assemblies (config 2 triangles, a 61^3 tet box) and a
short PCG run, so each kernel of interest appears a handful of times (`p2`: config 3, P2 elasticity)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import fes_b200 as F  # noqa: E402
from fes_b200.regions import on_plane  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
if what in ('all', 'tri'):
    mesh = F.create_rect_mesh([[0, 0], [4, 1]], (2829, 708))
    V = F.FunctionSpace(mesh, n_components=2)
    form = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3))
    for _ in range(4):
        A = V.assemble_device(form)
    torch.cuda.synchronize()
    if what == 'all':
        bc = F.BoundaryConditions()
        bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0, 0])
        bc.add(F.BCType.NEUMANN, on_plane(0, 4.0), [0, -1])
        prob = F.LinearProblem(V, form, None, bc)
        system = F.DiscreteSystem(prob.tangent(None), prob.constraints, F.JacobiPCGBackend(rtol=1e-10, maxiter=40))
        try:
            system.solve(prob.load_d)
        except RuntimeError as e:
            print('expected:', str(e)[:60])
if what == 'trimass':
    mesh = F.create_rect_mesh([[0, 0], [4, 1]], (2829, 708))
    V = F.FunctionSpace(mesh, n_components=2)
    for _ in range(4):
        A = V.assemble_device(F.MassForm(2))
    torch.cuda.synchronize()
if what == 'p2':
    mesh = F.create_rect_mesh([[0, 0], [1, 1]], (708, 708))
    V = F.FunctionSpace(mesh, F.QuadraticTriangleElement, n_components=2)
    form = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3))
    for _ in range(4):
        A = V.assemble_device(form)
    torch.cuda.synchronize()
if what in ('all', 'tet'):
    mesh = F.create_box_mesh([[0, 0, 0], [1, 1, 1]], (61, 61, 61))
    V = F.FunctionSpace(mesh, n_components=3)
    form = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3))
    for _ in range(4):
        A = V.assemble_device(form)
    torch.cuda.synchronize()
if what in ('simp',):
    from fes_b200.regions import in_box, intersect
    w, h = 3.0, 1.0
    m4 = F.create_rect_mesh([[0, 0], [w, h]], (1201, 401))
    bc = F.BoundaryConditions()
    bottom, top = on_plane(1, 0.0), on_plane(1, h)
    bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
    bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
    bc.add(F.BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
    topo = F.TopologyOptimizer(m4, F.LinearElastic(200.0, 0.4), bc, iters=1, volume_frac=0.5,
                               smoothing_radius=0.00625, penalty=3.0, history=None,
                               backend=F.JacobiPCGBackend(rtol=1e-10, maxiter=int(sys.argv[2]) if len(sys.argv) > 2 else 200))
    try:
        topo.step()
    except RuntimeError as e:
        print('expected:', str(e)[:60])
    torch.cuda.synchronize()
print('done')
