"""Time the Jacobi-PCG kernel: microseconds per iteration on the config-4 (SIMP) and config-2
operators with a capped iteration count."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import fes_b200 as F  # noqa: E402
from fes_b200.regions import on_plane  # noqa: E402

its = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for name, corners, res, c3 in [('config4 1201x401', [[0, 0], [3, 1]], (1201, 401), False),
                               ('config2 2829x708', [[0, 0], [4, 1]], (2829, 708), False),
                               ('tets 81^3', [[0, 0, 0], [1, 1, 1]], (81, 81, 81), True)]:
    if c3:
        mesh = F.create_box_mesh(corners, res)
        load = [0, -5, 0]
    else:
        mesh = F.create_rect_mesh(corners, res)
        load = [0, -1]
    d = mesh.spatial_dim
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0] * d)
    bc.add(F.BCType.NEUMANN, on_plane(0, corners[1][0]), load)
    prob = F.linear_elastic(mesh, F.LinearElasticMaterial(200.0, 0.3), bc)
    system = F.DiscreteSystem(prob.tangent(None), prob.constraints, F.JacobiPCGBackend(rtol=1e-30, maxiter=its))
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            system.solve(prob.load_d)
        except RuntimeError:
            pass
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    n, nnz = prob.space.n_dofs, prob.tangent(None).nnz
    by = (8 + 4.0 / d ** 2) * nnz + 11 * 8 * n
    print(f'{name}: n={n} nnz={nnz} {1e6 * dt / its:.1f} us/iter  ~{by / (dt / its) / 1e9:.0f} GB/s effective', flush=True)
    del system, prob
