"""Pin the numpy oracle to the reference's own run: the series decoded from experiment_data/img/*.pdf
(tools/extract_golden.py -> tests/golden/acados_{force,jerk}.npz) were produced by acados through main.py as
committed (seed 42, noise on; force run then jerk run sharing one noise stream)."""
import os

import numpy as np
import pytest

from oracle import nmpc_oracle as o

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
X0 = np.array([1.0, 0.0, 0.0, 0.62])          # reference src/main.py:45


def _max_err(series, g, name):
    st, v = g[name + '_step'], g[name]
    m = st < len(series)
    return float(np.max(np.abs(series[st[m]] - v[m])))


def test_constants_match_survey_appendix_b():
    assert o.GRAVITY == pytest.approx(0.32147370000000003, abs=0)
    assert o.MAX_F == pytest.approx(0.41791581000000005, abs=0)
    assert o.MIN_F == pytest.approx(-0.06429474, abs=1e-17)
    eps = o.main_py_noise()
    assert eps[0] == 0.004967141530112327 and eps[1] == -0.0013826430117118466
    ref = o.gen_circle_traj()
    assert ref.shape == (530, 8)
    np.testing.assert_allclose(ref[0], [1, 0, 0, 0.62831853, -0.39478418, 9.81, 0, 0], atol=5e-9)
    np.testing.assert_allclose(ref[529], [0.9340684, 0.35709413, -0.22436886, 0.58689249, -0.36875542, 9.66902489, 0, 0],
                               atol=5e-9)


def test_step0_known_answers():
    """First solve of each closed loop against the decoded acados values."""
    ref = o.gen_circle_traj()
    gf = np.load(os.path.join(GOLD, 'acados_force.npz'))
    r = o.follow_trajectory_force(ref[:, :4], ref[:, 4:6], X0, np.zeros(500), n_steps=1)
    assert abs(r['U_plant'][0, 0] - gf['theta'][0]) < 5e-8
    assert abs(r['U_plant'][0, 1] - gf['Fd'][0]) < 5e-8
    gj = np.load(os.path.join(GOLD, 'acados_jerk.npz'))
    r = o.follow_trajectory_jerk(ref[:, :6], ref[:, 6:], X0, np.zeros(500), n_steps=1)
    assert abs(r['U_plant'][0, 0] - gj['theta'][0]) < 5e-8
    assert abs(r['U_plant'][0, 1] - gj['Fd'][0]) < 5e-8
    assert abs(r['a'][0, 0] - gj['ax'][0]) < 5e-7 and abs(r['a'][0, 1] - gj['az'][0]) < 5e-7


@pytest.mark.slow
def test_force_closed_loop_matches_acados_run():
    """All 500 steps, 381 of them with an input bound active: the restated HPIPM iterate sequence reproduces
    acados to plot resolution (theta axis: 4.8e-8 rad per 1e-6 pt; observed max 1.5e-7)."""
    g = np.load(os.path.join(GOLD, 'acados_force.npz'))
    ref = o.gen_circle_traj()
    eps = o.main_py_noise()
    r = o.follow_trajectory_force(ref[:, :4], ref[:, 4:6], X0, eps[:500])
    assert np.all(r['status'] == 0)
    assert _max_err(r['U_plant'][:, 0], g, 'theta') < 5e-7
    assert _max_err(r['U_plant'][:, 1], g, 'Fd') < 1e-7
    for name, col in (('px', 0), ('pz', 1), ('vx', 2), ('vz', 3)):
        assert _max_err(r['Xsim'][:, col], g, name) < 1e-7
    assert r['cost'] == pytest.approx(66.3063, abs=1e-3)
    assert o.calc_aed(ref[:500, :2], r['Xsim'][:500, :2]) == pytest.approx(0.016563, abs=1e-5)


@pytest.mark.slow
def test_jerk_closed_loop_matches_acados_run():
    g = np.load(os.path.join(GOLD, 'acados_jerk.npz'))
    ref = o.gen_circle_traj()
    eps = o.main_py_noise()
    r = o.follow_trajectory_jerk(ref[:, :6], ref[:, 6:], X0, eps[500:])
    assert np.all(r['status'] == 0)
    assert _max_err(r['U_plant'][:, 0], g, 'theta') < 5e-8
    assert _max_err(r['U_plant'][:, 1], g, 'Fd') < 2e-8
    for name, col in (('px', 0), ('pz', 1), ('vx', 2), ('vz', 3)):
        assert _max_err(r['Xsim'][:, col], g, name) < 1e-7
    assert _max_err(r['a'][:, 0], g, 'ax') < 1e-6 and _max_err(r['a'][:, 1], g, 'az') < 1e-6
    assert r['cost'] == pytest.approx(456.636, abs=1e-2)


def test_erk_sensitivities_against_finite_differences():
    rng = np.random.default_rng(0)
    p = (o.MASS, o.GRAVITY_ACC)
    for f, jac, nx in ((o.f_force, o.jac_force, 4), (o.f_jerk, o.jac_jerk, 6), (o.f_plant, o.jac_plant, 4)):
        for stages in (1, 2, 3, 4):
            x = rng.normal(size=nx); u = rng.normal(size=2) * 0.3
            xn, S = o.erk_step(f, jac, x, u, p, 0.02, stages)
            h = 1e-6
            for j in range(nx + 2):
                d = np.zeros(nx + 2); d[j] = h
                xp = o.erk_step(f, jac, x + d[:nx], u + d[nx:], p, 0.02, stages, sens=False)
                xm = o.erk_step(f, jac, x - d[:nx], u - d[nx:], p, 0.02, stages, sens=False)
                np.testing.assert_allclose((xp - xm) / (2 * h), S[:, j], atol=1e-8)


def test_force_erk4_is_exact_discretisation():
    """for the affine force model ERK4 equals the exact map (so it stands in for acados' IRK)."""
    h, m, g = o.DT, o.MASS, o.GRAVITY_ACC
    x = np.array([0.3, -0.2, 0.5, -0.4]); u = np.array([0.1, 0.35])
    xn, S = o.erk_step(o.f_force, o.jac_force, x, u, (m, g), h, 4)
    acc = np.array([u[0] / m, u[1] / m - g])
    np.testing.assert_allclose(xn, np.hstack([x[:2] + h * x[2:] + 0.5 * h * h * acc, x[2:] + h * acc]), rtol=0, atol=1e-15)
    np.testing.assert_allclose(S[:, 4:], np.vstack([np.eye(2) * h * h / (2 * m), np.eye(2) * h / m]), atol=1e-15)
