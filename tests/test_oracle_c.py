"""The C oracle (Riccati recursion) against the numpy oracle (dense KKT) and against the decoded acados run."""
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import nmpc_oracle as o

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
P = np.array([[o.MASS, o.GRAVITY_ACC]])


def _max_err(series, g, name):
    st, v = g[name + '_step'], g[name]
    return float(np.max(np.abs(series[st] - v)))


@pytest.mark.parametrize('model,tag,lo', [(co.MODEL_FORCE, 'force', 0), (co.MODEL_JERK, 'jerk', 500)])
def test_c_oracle_reproduces_acados_run(model, tag, lo):
    g = np.load(os.path.join(GOLD, f'acados_{tag}.npz'))
    eps = o.main_py_noise()[lo:lo + 500]
    r = co.closed_loop(co.default_opts(model), o.gen_circle_traj(), np.array([[1.0, 0, 0, 0.62]]), eps[:, None], P, P, 500)
    assert np.all(r['status'] == 0)
    assert _max_err(r['U_plant'][0, :, 0], g, 'theta') < 5e-7
    assert _max_err(r['U_plant'][0, :, 1], g, 'Fd') < 1e-7
    for name, col in (('px', 0), ('pz', 1), ('vx', 2), ('vz', 3)):
        assert _max_err(r['Xsim'][0, :, col], g, name) < 1e-7


@pytest.mark.parametrize('model', [co.MODEL_FORCE, co.MODEL_JERK])
def test_c_vs_numpy_single_solves(model):
    """Random x0 / reference phases, some with bounds active; compare whole solutions and integer outputs."""
    rng = np.random.default_rng(5)
    spec = o.force_ocp() if model == co.MODEL_FORCE else o.jerk_ocp()
    nx, nu, N = spec.nx, spec.nu, spec.N
    B = 6
    x0s, yrefs = [], []
    for i in range(B):
        ref = o.gen_circle_traj(radius=rng.uniform(0.5, 1.0), center=rng.uniform(-0.15, 0.15, 2), phase=rng.uniform(0, 2 * np.pi))
        st = rng.integers(0, 400)
        x0 = ref[st, :4] + rng.uniform(-0.08, 0.08, 4)
        if model == co.MODEL_JERK:
            x0 = np.hstack([x0, [rng.uniform(-1, 1), o.GRAVITY_ACC + rng.uniform(-1, 1)]])
            y = np.hstack([ref[st:st + N, :8].ravel(), ref[st + N, :6]])
        else:
            y = np.hstack([ref[st:st + N, :6].ravel(), ref[st + N, :4]])
        x0s.append(x0); yrefs.append(y)
    x0s, yrefs = np.array(x0s), np.array(yrefs)
    rc = co.solve_batch(co.default_opts(model), x0s, yrefs, np.repeat(P, B, 0))
    for i in range(B):
        sol = o.OracleOcpSolver(spec)
        ny = nx + nu
        for k in range(N):
            sol.set(k, 'yref', yrefs[i, k * ny:(k + 1) * ny])
        sol.set(N, 'yref', yrefs[i, N * ny:])
        sol.set(0, 'lbx', x0s[i])
        st = sol.solve()
        assert st == rc['status'][i]
        assert sol.sqp_iter == rc['sqp_iter'][i] and sol.qp_iter == rc['qp_iter'][i]
        np.testing.assert_allclose(rc['u'][i], sol.u, rtol=0, atol=1e-9)
        np.testing.assert_allclose(rc['x'][i], sol.x, rtol=0, atol=1e-9)
        np.testing.assert_allclose(rc['pi'][i], sol.pi, rtol=0, atol=1e-8)


def test_threads_do_not_change_results():
    rng = np.random.default_rng(1)
    B = 16
    ref = o.gen_circle_traj()
    x0 = ref[0, :4] + rng.uniform(-0.05, 0.05, (B, 4))
    noise = rng.normal(0, 0.01, (20, B))
    pp = np.repeat(P, B, 0)
    a = co.closed_loop(co.default_opts(co.MODEL_FORCE), ref, x0, noise, pp, pp, 20, nthreads=1)
    b = co.closed_loop(co.default_opts(co.MODEL_FORCE), ref, x0, noise, pp, pp, 20, nthreads=4)
    assert np.array_equal(a['Xsim'], b['Xsim']) and np.array_equal(a['qp_iter'], b['qp_iter'])


def test_c_vs_numpy_nonlinear_thrust_ocp():
    """The thrust OCP (plant model as controller model, not in the reference) exercises the general SQP path: several
    SQP iterations, sensitivities that change with the iterate.  The two oracles must agree there too."""
    from common import thrust_solve_inputs
    B = 4
    x0s, yrefs = thrust_solve_inputs(B, seed=3)
    rc = co.solve_batch(co.default_opts(co.MODEL_THRUST), x0s, yrefs, np.repeat(P, B, 0))
    assert rc['sqp_iter'].min() >= 2
    spec = o.thrust_ocp()
    for i in range(B):
        sol = o.OracleOcpSolver(spec)
        for k in range(30):
            sol.set(k, 'yref', yrefs[i, k * 6:(k + 1) * 6])
        sol.set(30, 'yref', yrefs[i, 180:])
        sol.set(0, 'lbx', x0s[i])
        assert sol.solve() == rc['status'][i]
        assert sol.sqp_iter == rc['sqp_iter'][i] and sol.qp_iter == rc['qp_iter'][i]
        np.testing.assert_allclose(rc['u'][i], sol.u, rtol=0, atol=1e-9)
        np.testing.assert_allclose(rc['x'][i], sol.x, rtol=0, atol=1e-9)


def test_irk_gl4_integrator():
    """erk_stages = 0 = acados IRK (Gauss-Legendre, 4 stages, 3 Newton iterations, IFT sensitivities), the integrator_type
    of the reference's force OCP (src/force_model/ocp.py:85).  (a) On the affine force model and on the plant (linear in
    the state for a constant input) it returns the exact discretisation = ERK4, state and sensitivities, to 1e-14.  (b) On
    a system that is nonlinear in the state it agrees with a tight-tolerance scipy solve_ivp to 1e-9 and its sensitivities
    with finite differences.  (c) The OCP solved with it (C oracle, numpy oracle) equals the ERK4 solution to 1e-12 with
    identical iteration counts."""
    from scipy.integrate import solve_ivp
    from common import P_NOM, random_solve_inputs, thrust_solve_inputs
    rng = np.random.default_rng(4)
    p = (o.MASS, o.GRAVITY_ACC)
    for f, jac, u in ((o.f_force, o.jac_force, np.array([0.1, 0.3])), (o.f_plant, o.jac_plant, np.array([0.4, 0.35]))):
        x = rng.normal(size=4)
        xe, Se = o.erk_step(f, jac, x, u, p, 0.02, 4)
        xi, Si = o.erk_step(f, jac, x, u, p, 0.02, 0)
        assert np.abs(xe - xi).max() < 1e-14 and np.abs(Se - Si).max() < 1e-14
        sol = solve_ivp(lambda t, y: f(y, u, p), [0, 0.02], x, rtol=1e-13, atol=1e-14)
        assert np.abs(sol.y[:, -1] - xi).max() < 1e-12
    f = lambda x, u, p: np.array([x[1], -9.0 * np.sin(x[0]) + u[0]])                     # pendulum: Newton has work to do
    jac = lambda x, u, p: (np.array([[0, 1.0], [-9.0 * np.cos(x[0]), 0]]), np.array([[0.0], [1.0]]))
    x, u = np.array([1.0, 0.5]), np.array([0.3])
    xi, Si = o.erk_step(f, jac, x, u, p, 0.1, 0)
    sol = solve_ivp(lambda t, y: f(y, u, p), [0, 0.1], x, rtol=1e-13, atol=1e-14)
    assert np.abs(sol.y[:, -1] - xi).max() < 1e-9
    eps, fd = 1e-6, np.zeros((2, 3))
    for j in range(3):
        dx, du = np.zeros(2), np.zeros(1)
        (dx if j < 2 else du)[j if j < 2 else 0] = eps
        fd[:, j] = (o.erk_step(f, jac, x + dx, u + du, p, 0.1, 0, sens=False) - o.erk_step(f, jac, x - dx, u - du, p, 0.1, 0, sens=False)) / (2 * eps)
    assert np.abs(fd - Si).max() < 1e-8
    for model, gen in ((co.MODEL_FORCE, lambda: random_solve_inputs(0, 5, seed=3)), (co.MODEL_THRUST, lambda: thrust_solve_inputs(5, seed=4))):
        x0, yref = gen()
        pb = np.repeat(P_NOM[None], 5, 0)
        a = co.solve_batch(co.default_opts(model), x0, yref, pb)
        b = co.solve_batch(co.default_opts(model, erk_stages=0), x0, yref, pb)
        assert np.array_equal(a['qp_iter'], b['qp_iter']) and np.array_equal(a['sqp_iter'], b['sqp_iter']) and np.array_equal(a['status'], b['status'])
        np.testing.assert_allclose(a['u'], b['u'], rtol=0, atol=1e-12)
    spec = o.force_ocp(); spec.erk_stages = 0
    s_irk, s_erk = o.OracleOcpSolver(spec), o.OracleOcpSolver(o.force_ocp())
    x0, yref = random_solve_inputs(0, 1, seed=5)
    for s in (s_irk, s_erk):
        for k in range(30):
            s.set(k, 'yref', yref[0, k * 6:(k + 1) * 6])
        s.set(30, 'yref', yref[0, 180:])
        s.set(0, 'lbx', x0[0]); s.set(0, 'ubx', x0[0])
        assert s.solve() == 0
    np.testing.assert_allclose(s_irk.get(0, 'u'), s_erk.get(0, 'u'), rtol=0, atol=1e-12)
