"""ncu driver: four fused assemblies of the config-5 tet box (151^3 nodes by default, P1 elasticity)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import fes_b200 as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 151
mesh = F.create_box_mesh([[0, 0, 0], [1, 1, 1]], (n, n, n))
V = F.FunctionSpace(mesh, n_components=3)
form = F.LinearElasticForm(F.LinearElasticMaterial(200.0, 0.3))
for _ in range(4):
    A = V.assemble_device(form)
torch.cuda.synchronize()
print('done')
