"""GPU parity of the device-resident time integrators (ThetaMethod, NewmarkMethod) against histories
recorded from the unmodified reference (tests/golden/transient.npz, ref fem/integrators.py).  The
reference stepped with a direct solver; here every step is a Jacobi-PCG solve at rtol 1e-12, so the
histories agree to 1e-8 of the field's magnitude (SURVEY 8d solution bar)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def fes():
    import fes_b200
    return fes_b200


def close(a, b, rtol=1e-8):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * float(np.abs(b).max()))


def heat_problem(F, g):
    from fes_b200.regions import on_plane
    mesh = F.Mesh(g['heat_vertices'], g['heat_elements'], g['heat_boundary'])
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [1.5])
    bc.add(F.BCType.NEUMANN, on_plane(0, 2.0), [-0.75])
    return F.heat(mesh, source=lambda p: np.array([np.sin(3 * p[0]) + p[1]]), bc=bc)


def test_theta_method_matches_reference():
    F, g = fes(), load_golden('transient')
    prob = heat_problem(F, g)
    np.testing.assert_array_equal(prob.constraints[0], g['heat_free'])
    np.testing.assert_allclose(prob.load, g['heat_load'], rtol=1e-12, atol=1e-14)
    backend = F.JacobiPCGBackend(rtol=1e-12)
    for tag, theta in (('cn', 0.5), ('be', 1.0)):
        sol = F.ThetaMethod(dt=0.02, steps=6, theta=theta, backend=backend).run(prob, g['heat_u0'])
        np.testing.assert_allclose(sol.t, g[f'heat_{tag}_t'], rtol=1e-14)
        close(np.asarray(sol.u), g[f'heat_{tag}_u'])
    with pytest.raises(ValueError):
        F.ThetaMethod(dt=0.02, steps=1).run(prob, np.zeros(3))


def test_newmark_matches_reference_and_conserves_energy():
    F, g = fes(), load_golden('transient')
    from fes_b200.regions import on_plane
    mesh = F.Mesh(g['heat_vertices'], g['heat_elements'], g['heat_boundary'])
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [0.0])
    bc.add(F.BCType.DIRICHLET, on_plane(0, 2.0), [0.0])
    prob = F.wave(mesh, 1.7, bc)
    sol = F.NewmarkMethod(dt=0.01, steps=6, backend=F.JacobiPCGBackend(rtol=1e-12)).run(prob, g['wave_u0'], g['wave_v0'])
    close(np.asarray(sol.u), g['wave_u'])
    close(np.asarray(sol.dudt), g['wave_v'])
    e = np.array([F.wave_energy(prob, sol.u[i], sol.dudt[i]) for i in range(len(sol.u))])
    np.testing.assert_allclose(e, g['wave_energy'], rtol=1e-8)
    with pytest.raises(ValueError, match='v0 must be zero'):
        F.NewmarkMethod(dt=0.01, steps=1).run(prob, g['wave_u0'], np.ones(len(g['wave_u0'])))
    with pytest.raises(ValueError, match='u0 disagrees'):
        F.NewmarkMethod(dt=0.01, steps=1).run(prob, g['wave_u0'] + 1.0, g['wave_v0'])


def test_heat_steps_at_scale_stay_on_the_device():
    """1201 x 401 nodes (config-4 mesh): 20 Crank-Nicolson steps; the discrete maximum principle bounds the
    field by its data and every step's solve converges (no exception)."""
    F = fes()
    from fes_b200.regions import on_plane
    mesh = F.create_rect_mesh([[0, 0], [3, 1]], (1201, 401))
    bc = F.BoundaryConditions()
    bc.add(F.BCType.DIRICHLET, on_plane(0, 0.0), [1.0])
    bc.add(F.BCType.DIRICHLET, on_plane(0, 3.0), [0.0])
    prob = F.heat(mesh, bc=bc)
    u0 = np.clip(1.0 - mesh.vertices[:, 0] / 3.0, 0.0, 1.0)
    sol = F.ThetaMethod(dt=1e-3, steps=20, theta=1.0).run(prob, u0)
    u = sol.u[-1]
    assert u.min() > -1e-9 and u.max() < 1.0 + 1e-9
    np.testing.assert_allclose(u, u0, atol=1e-6)      # the linear profile is the steady state
