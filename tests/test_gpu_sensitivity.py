"""GPU parity of the adjoint-sensitivity and design-optimization mirrors (fes_b200.sensitivity,
fes_b200.design) against gradients and a DesignOptimizer trajectory recorded from the unmodified
reference (tests/golden/adjoint.npz; ref fem/sensitivity.py, fem/design.py).  Solves are Jacobi-PCG at
rtol 1e-12: gradients to 1e-8 of their maximum, teacher-forced densities to 1e-10."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def fes():
    import fes_b200
    return fes_b200


def close(a, b, rtol=1e-8):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * float(np.abs(b).max()))


def setup(F, g):
    from fes_b200.regions import on_plane
    mesh = F.Mesh(g['vertices'], g['elements'], g['boundary'])
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0.0, 0.0])
    bc.add(F.BCType.NEUMANN, on_plane(0, 1.0), [0.0, -1.0])
    return mesh, bc, F.FunctionSpace(mesh, n_components=2)


def test_adjoint_gradients_match_reference():
    F, g = fes(), load_golden('adjoint')
    mesh, bc, space = setup(F, g)
    backend = F.JacobiPCGBackend(rtol=1e-12)
    model = F.SIMPModel(space, base_E=1.0, nu=0.3, bc=bc, penalty=3.0)
    prob = model.problem(g['rho'])
    an = F.SensitivityAnalysis(prob, backend)
    u = an.solve_forward()
    close(u, g['u_rho'])
    tip = int(g['tip_dof'])
    assert abs(F.Compliance().value(prob, u) - g['compliance_value']) < 1e-8 * abs(g['compliance_value'])
    assert abs(F.PointValue(tip).value(prob, u) - g['point_value']) < 1e-8 * abs(g['point_value'])
    par = model.parameterization(g['rho'])
    assert par.size == len(g['rho'])
    close(an.gradient(F.Compliance(), par, u), g['grad_compliance_density'])
    close(an.adjoint(F.PointValue(tip), u), g['lam_point'])
    close(an.gradient(F.PointValue(tip), par, u), g['grad_point_density'])
    # device vectors in -> device gradient out, same numbers
    gd = an.gradient(F.PointValue(tip), par, torch.as_tensor(u, device=space.device))
    assert isinstance(gd, torch.Tensor)
    close(gd.cpu().numpy(), g['grad_point_density'])
    probE = F.LinearProblem(space, F.LinearElasticForm(F.LinearElasticMaterial(g['E_el'], 0.3)), None, bc)
    anE = F.SensitivityAnalysis(probE, backend)
    uE = anE.solve_forward()
    close(uE, g['u_E'])
    close(anE.gradient(F.PointValue(tip), F.ModulusField.create(space, g['E_el'], 0.3), uE), g['grad_point_modulus'])
    close(anE.gradient(F.Compliance(), F.ModulusField.create(space, g['E_el'], 0.3), uE), g['grad_compliance_modulus'])


def test_design_optimizer_matches_reference_trajectory():
    F, g = fes(), load_golden('adjoint')
    mesh, bc, space = setup(F, g)
    from fes_b200.topology import ConeFilter
    filt = ConeFilter(space, float(g['filter_radius']))
    model = F.SIMPModel(space, base_E=1.0, nu=0.3, bc=bc, penalty=3.0, sensitivity_filter=filt)
    backend = F.JacobiPCGBackend(rtol=1e-12)
    # teacher-forced single steps from the reference's densities
    for k in range(4):
        d = F.DesignOptimizer(model, F.Compliance(), volume_frac=0.5, iters=1, move=0.1, backend=backend)
        d._rho = torch.as_tensor(g['design_rho'][k], device=space.device)
        u, value = d.step()
        assert abs(value - g['design_objective'][k]) < 1e-8 * g['design_objective'][k]
        close(u, g['design_u'][k])
        np.testing.assert_allclose(d.rho, g['design_rho'][k + 1], rtol=1e-10, atol=1e-10)
    # free-running, the reference's own acceptance bar (tests/test_design.py:46-47 compares 5 iterations)
    seen = []
    design = F.DesignOptimizer(model, F.Compliance(), volume_frac=0.5, iters=5, move=0.1, backend=backend)
    hist = design.solve(on_iteration=lambda i, rho, j: seen.append(i))
    assert seen == list(range(5))
    np.testing.assert_allclose(np.asarray(hist.rho), g['design_rho'], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(hist.objective, g['design_objective'], rtol=1e-8)
    np.testing.assert_allclose(design.rho, g['design_rho_final'], rtol=1e-8, atol=1e-8)
    vol = space.element_volumes
    assert abs(float((vol * design.rho).sum() / vol.sum()) - 0.5) < 1e-6


def test_signed_sensitivity_is_rejected_like_the_reference():
    """A raw point-displacement objective gives a signed sensitivity: OC must fail loudly
    (ref tests/test_design.py:72-84)."""
    F, g = fes(), load_golden('adjoint')
    mesh, bc, space = setup(F, g)
    model = F.SIMPModel(space, base_E=1.0, nu=0.3, bc=bc, penalty=3.0)
    design = F.DesignOptimizer(model, F.PointValue(int(g['tip_dof'])), volume_frac=0.5, iters=1)
    with pytest.raises(ValueError, match='nonnegative'):
        design.step()
