"""Small driver for ncu: a few fused assemblies (config 2 triangles, a 61^3 tet box) and a
short PCG run, so each kernel of interest appears a handful of times."""
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
if what in ('all', 'tet'):
    mesh = F.create_box_mesh([[0, 0, 0], [1, 1, 1]], (61, 61, 61))
    V = F.FunctionSpace(mesh, n_components=3)
    form = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3))
    for _ in range(4):
        A = V.assemble_device(form)
    torch.cuda.synchronize()
print('done')
