"""Warm fused-assembly timings (CUDA events, 20 replays) for every assembly configuration:
   c1  P1 Laplacian 225^2          c2  P1 elastic 2829x708      c3  P2 elastic + mass 708^2
   c4  P1 elastic 1201x401 (+rho)  tet P1 elastic tets n^3 (default 101)
Prints one JSON line per case with ms, elements/s and the algorithmic-byte HBM fraction."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import fes_b200 as F  # noqa: E402

PEAK = 6552.0
which = sys.argv[1].split(',') if len(sys.argv) > 1 else ['c1', 'c2', 'c3', 'c4', 'tet']
tet_n = int(sys.argv[2]) if len(sys.argv) > 2 else 101


def timeit(V, form, tag, coef_bytes=0, reps=20):
    t0 = time.perf_counter()
    plan = V.plan()
    torch.cuda.synchronize()
    plan_s = time.perf_counter() - t0
    A = V.assemble_device(form)
    for _ in range(3):
        V.assemble_device(form, out=A.data_d)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        V.assemble_device(form, out=A.data_d)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    n_el = V._n_elements()
    N = V._en_d.shape[1]
    alg = 4 * N * n_el + 8 * V.spatial_dim * V.n_nodes + coef_bytes + 8 * plan.nnz
    print(json.dumps({'case': tag, 'ms': round(ms, 4), 'elements': n_el, 'Gelem_s': round(n_el / ms / 1e6, 3),
                      'nnz': plan.nnz, 'alg_MB': round(alg / 1e6, 1), 'hbm_frac': round(alg / ms / 1e6 / PEAK, 4),
                      'plan_s': round(plan_s, 3), 'plan_read_MB': round(plan.read_bytes / 1e6, 1),
                      'checksum': float(A.data_d.sum().item())}), flush=True)


mat = F.LinearElasticMaterial(200.0, 0.3)
if 'c1' in which:
    V = F.FunctionSpace(F.create_rect_mesh([[0, 0], [1, 1]], (225, 225)))
    timeit(V, F.LaplacianForm(), 'c1 P1 tri Laplacian 225^2')
if 'c2' in which:
    V = F.FunctionSpace(F.create_rect_mesh([[0, 0], [4, 1]], (2829, 708)), n_components=2)
    timeit(V, F.LinearElasticForm(mat), 'c2 P1 tri elastic 2829x708')
    timeit(V, F.MassForm(2), 'c2 P1 tri vector mass 2829x708')
if 'c4' in which:
    V = F.FunctionSpace(F.create_rect_mesh([[0, 0], [3, 1]], (1201, 401)), n_components=2)
    n_el = V._n_elements()
    rho = np.linspace(0.2, 1.0, n_el) ** 3
    timeit(V, F.LinearElasticForm(F.LinearElasticMaterial(200.0 * rho, 0.4)), 'c4 P1 tri elastic per-element E 1201x401',
           coef_bytes=8 * n_el)
if 'c3' in which:
    V = F.FunctionSpace(F.create_rect_mesh([[0, 0], [1, 1]], (708, 708)), F.QuadraticTriangleElement, n_components=2)
    timeit(V, F.LinearElasticForm(mat), 'c3 P2 tri elastic 708^2')
    timeit(V, F.MassForm(2), 'c3 P2 tri vector mass 708^2')
if 'tet' in which:
    V = F.FunctionSpace(F.create_box_mesh([[0, 0, 0], [1, 1, 1]], (tet_n,) * 3), n_components=3)
    timeit(V, F.LinearElasticForm(mat), f'tet P1 elastic {tet_n}^3', reps=5)
    V1 = F.FunctionSpace(V.mesh, n_components=1)
    timeit(V1, F.LaplacianForm(), f'tet P1 Laplacian {tet_n}^3', reps=5)
