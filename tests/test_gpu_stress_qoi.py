"""GPU parity of the stress quantities of interest (fes_b200.MeanStress / SoftMaxStress: fes_von_mises_gradient +
fes_node_scatter) against values, adjoint loads and density gradients recorded from the unmodified reference
(tests/golden/stress_qoi.npz; ref fem/sensitivity.py:108-276)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
CASES = {'mean': (False, None), 'mean_region_var': (True, None), 'soft': (False, 8.0), 'soft_region_var': (True, 6.0)}


def close(a, b, rtol):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * float(np.abs(b).max()))


def setup(name):
    import fes_b200 as F
    from fes_b200.regions import on_plane
    g = load_golden('stress_qoi')
    mesh = F.Mesh(g[f'{name}_vertices'], g[f'{name}_elements'], g[f'{name}_boundary'])
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0.0, 0.0])
    bc.add(F.BCType.NEUMANN, on_plane(0, 1.0), [0.0, -1.0])
    etype = F.QuadraticTriangleElement if name == 'p2' else None
    return F, g, bc, F.FunctionSpace(mesh, etype, n_components=2)


@pytest.mark.parametrize('name', ['p1', 'p2'])
def test_stress_qoi_values_loads_and_gradients(name):
    F, g, bc, space = setup(name)
    model = F.SIMPModel(space, base_E=1.0, nu=0.3, bc=bc, penalty=3.0)
    rho = g[f'{name}_rho']
    prob = model.problem(rho)
    an = F.SensitivityAnalysis(prob, F.JacobiPCGBackend(rtol=1e-12))
    u = an.solve_forward()
    close(u, g[f'{name}_u'], 1e-8)
    u_ref = g[f'{name}_u']          # the reference's own state: isolates the quantity from the solve
    for tag, (var, p) in CASES.items():
        mat = F.LinearElasticMaterial(g[f'{name}_E_el'] if var else 1.0, 0.3)
        region = g[f'{name}_region'] if var else None
        qoi = F.MeanStress(space, mat, region) if p is None else F.SoftMaxStress(space, mat, region, p)
        assert not qoi.self_adjoint
        np.testing.assert_allclose(qoi.value(prob, u_ref), g[f'{name}_{tag}_value'], rtol=1e-12)
        dJ = qoi.dJ_du(prob, u_ref)
        assert isinstance(dJ, torch.Tensor) and dJ.shape == (space.n_dofs,)
        close(dJ.cpu().numpy(), g[f'{name}_{tag}_dJ_du'], 1e-10)
        close(an.gradient(qoi, model.parameterization(rho), u), g[f'{name}_{tag}_grad'], 1e-7)


def test_stress_qoi_is_plane_strain_2d_only():
    import fes_b200 as F
    mesh = F.create_box_mesh([[0, 0, 0], [1, 1, 1]], (3, 3, 3))
    V = F.FunctionSpace(mesh, n_components=3)
    with pytest.raises(NotImplementedError):
        F.MeanStress(V, F.LinearElasticMaterial(1.0, 0.3)).value(None, np.zeros(V.n_dofs))
    m2 = F.create_rect_mesh([[0, 0], [1, 1]], (4, 4))
    V2 = F.FunctionSpace(m2, n_components=2)
    assert F.SoftMaxStress(V2, F.LinearElasticMaterial(1.0, 0.3)).dJ_du(None, np.zeros(V2.n_dofs)).abs().max().item() == 0.0
