"""Device-only timing of the run kernel with per-element coefficients (the SIMP reassembly: rho^p as d_scale, an
E field as d_E) on config 4 (1201x401) and config 2 (2829x708): CUDA events around 20 fes_assemble calls."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import fes_b200 as F  # noqa: E402
from fes_b200 import _lib  # noqa: E402
from fes_b200.forms import Lowered  # noqa: E402

L = _lib.lib()
for name, corners, res in (('config4 1201x401', [[0, 0], [3, 1]], (1201, 401)), ('config2 2829x708', [[0, 0], [4, 1]], (2829, 708))):
    V = F.FunctionSpace(F.create_rect_mesh(corners, res), n_components=2)
    plan = V.plan()
    n_el = V._n_elements()
    values = torch.empty(plan.nnz, dtype=torch.float64, device=V.device)
    scale = torch.rand(n_el, dtype=torch.float64, device=V.device)
    Ef = 100.0 + 100.0 * torch.rand(n_el, dtype=torch.float64, device=V.device)
    for tag, low in (('uniform', Lowered(_lib.FORM_ELASTIC, E=200.0, nu=0.3)),
                     ('d_scale', Lowered(_lib.FORM_ELASTIC, E=200.0, nu=0.3, scale_elem=scale)),
                     ('d_E + d_scale', Lowered(_lib.FORM_ELASTIC, E=1.0, nu=0.3, E_elem=Ef, scale_elem=scale))):
        cf = low.coeffs()

        def step():
            _lib.check(L.fes_assemble(plan.handle, _lib.P1_TRI, _lib.FORM_ELASTIC, 2, _lib.ptr(V._en_d), _lib.ptr(V._coords_d),
                                      C.byref(cf), _lib.ptr(values), _lib.stream()))
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            step()
        e1.record()
        torch.cuda.synchronize()
        print(f'{name} {tag}: {e0.elapsed_time(e1) / 20:.4f} ms', flush=True)
