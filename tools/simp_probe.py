"""Config 4 (1201x401 MBB beam), n design iterations: CG iterations per design iteration and iterations/s.
Usage: simp_probe.py [n_iterations]   (FES_SIMP_EXTRAPOLATE=0|1 switches the warm-start extrapolation)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import fes_b200 as F  # noqa: E402
from fes_b200.regions import in_box, intersect, on_plane  # noqa: E402

n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 12
w, h = 3.0, 1.0
mesh = F.create_rect_mesh([[0, 0], [w, h]], (1201, 401))
bc = F.BoundaryConditions()
bottom, top = on_plane(1, 0.0), on_plane(1, h)
bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([None, None], [0.04 * w, None])), [0, 0])
bc.add(F.BCType.DIRICHLET, intersect(bottom, in_box([0.96 * w, None], [None, None])), [None, 0])
bc.add(F.BCType.NEUMANN, intersect(top, in_box([0.4 * w, None], [0.6 * w, None])), [0, -0.5])
topo = F.TopologyOptimizer(mesh, F.LinearElastic(200.0, 0.4), bc, iters=n_it, volume_frac=0.5, smoothing_radius=0.00625,
                           penalty=3.0, history=None, backend=F.JacobiPCGBackend(rtol=1e-10))
torch.cuda.synchronize()
t0 = time.perf_counter()
comps = [topo.step().total_compliance for _ in range(n_it)]
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print({'extrapolate': topo.extrapolate, 'it_per_s': round(n_it / dt, 4), 'cg': topo.cg_iterations, 'compliance_last': comps[-1]})
