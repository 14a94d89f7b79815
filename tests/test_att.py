"""3-D attitude-and-total-thrust model (BNMPC_MODEL_ATT; north-star extension, SURVEY 8f rank 2 - not in the reference).

CPU part: the model's Jacobians against finite differences, the two independently written oracles (dense KKT in numpy,
Riccati in C) against each other, the product's solver templates compiled for the host (tests/hostsim) against the C oracle,
and a closed-loop sanity check of the OCP formulation.  GPU part (-m gpu): the CUDA path through the C-ABI against the C
oracle - single solves, the device-resident closed loop (bnmpc_step_for_x0 per control step), the plant integrator, the
reference-style host loop, FP32.  Tolerances: status / SQP / QP iteration counts bit-exact, x / u / pi within 1e-9."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import nmpc_oracle as o

M, G = o.MASS, o.GRAVITY_ACC


def helix_ref(rows, radius=0.8, center=(0.0, 0.0, 0.0), phase=0.0, y_amp=0.3, dt=0.02):
    """[rows, 14] = [xref (10) | uref (4)]: circle in the x-z plane with a lateral sway, identity attitude, hover input"""
    t = np.arange(rows) * dt
    om = 2 * np.pi / 10
    a = om * t + phase
    r = np.zeros((rows, 14))
    r[:, 0] = center[0] + radius * np.cos(a); r[:, 1] = center[1] + y_amp * np.sin(2 * a); r[:, 2] = center[2] + radius * np.sin(a)
    r[:, 3] = -radius * om * np.sin(a); r[:, 4] = 2 * om * y_amp * np.cos(2 * a); r[:, 5] = radius * om * np.cos(a)
    r[:, 6] = 1.0
    r[:, 10] = M * G
    return r


def att_inputs(B, seed, rows=80, mass_sigma=0.05, spread=0.05):
    rng = np.random.default_rng(seed)
    refs = np.stack([helix_ref(rows, radius=rng.uniform(0.5, 0.9), center=rng.uniform(-0.1, 0.1, 3), phase=rng.uniform(0, 2 * np.pi),
                               y_amp=rng.uniform(0.1, 0.3)) for _ in range(B)])
    x0 = refs[:, 0, :10].copy()
    x0[:, :6] += rng.uniform(-spread, spread, (B, 6))
    # a tilted start attitude for some instances (unit quaternion about a random axis, up to ~0.3 rad)
    ang = rng.uniform(0, 0.3, B) * (rng.uniform(size=B) < 0.5)
    ax = rng.normal(size=(B, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    x0[:, 6] = np.cos(ang / 2); x0[:, 7:10] = ax * np.sin(ang / 2)[:, None]
    p_ctrl = np.tile([M, G], (B, 1))
    p_plant = p_ctrl.copy()
    p_plant[:, 0] *= 1 + np.clip(rng.normal(0, mass_sigma, B), -0.15, 0.15)
    return refs, x0, p_ctrl, p_plant


def window(refs, i, N=30):
    B = refs.shape[0]
    return np.hstack([refs[:, i:i + N, :].reshape(B, -1), refs[:, i + N, :10]])


def start_iterate(B, N=30, p=None):
    xg = np.zeros((B, N + 1, 10)); xg[:, :, 6] = 1.0
    ug = np.zeros((B, N, 4)); ug[:, :, 0] = (M * G) if p is None else (p[:, 0] * p[:, 1])[:, None]
    return xg, ug


# ---------------------------------------------------------------------------------------------------------------------
# CPU
# ---------------------------------------------------------------------------------------------------------------------
def test_att_jacobians_against_finite_differences():
    rng = np.random.default_rng(1)
    for _ in range(5):
        x = rng.normal(size=10); u = rng.normal(size=4); p = (M * rng.uniform(0.8, 1.2), G)
        fx, fu = o.jac_att(x, u, p)
        e = 1e-6
        fxn = np.array([(o.f_att(x + e * np.eye(10)[i], u, p) - o.f_att(x - e * np.eye(10)[i], u, p)) / (2 * e) for i in range(10)]).T
        fun = np.array([(o.f_att(x, u + e * np.eye(4)[i], p) - o.f_att(x, u - e * np.eye(4)[i], p)) / (2 * e) for i in range(4)]).T
        assert np.abs(fx - fxn).max() < 1e-7 and np.abs(fu - fun).max() < 1e-7


def test_att_restricted_to_the_plane_is_the_reference_plant():
    """pitch theta about y, no roll / yaw rate, no lateral motion: (px, pz, vx, vz) follow reference src/plant.py:27-33 with
    u = (theta, Fd) = (pitch, T)"""
    th, fd = 0.3, 0.4
    x = np.zeros(10); x[[0, 2, 3, 5]] = [0.1, -0.2, 0.3, 0.4]
    x[6] = np.cos(th / 2); x[8] = np.sin(th / 2)
    xd = o.f_att(x, np.array([fd, 0, 0, 0]), (M, G))
    want = o.f_plant(np.array([0.1, -0.2, 0.3, 0.4]), (th, fd), (M, G))
    np.testing.assert_allclose(xd[[0, 2, 3, 5]], want, rtol=0, atol=1e-14)
    assert abs(xd[1]) + abs(xd[4]) < 1e-15


def test_att_oracles_agree():
    """dense-KKT numpy oracle vs Riccati C oracle on one att OCP (several SQP iterations, active input bounds)"""
    N = 30
    refs, x0, pc, _ = att_inputs(1, seed=5, mass_sigma=0.0, spread=0.08)
    y = window(refs, 0)[0]
    s = o.OracleOcpSolver(o.att_ocp(), (M, G))
    for k in range(N):
        s.set(k, 'yref', y[k * 14:(k + 1) * 14])
    s.set(N, 'yref', y[N * 14:])
    xg, ug = start_iterate(1)
    for k in range(N + 1):
        s.set(k, 'x', xg[0, k])
    for k in range(N):
        s.set(k, 'u', ug[0, k])
    s.set(0, 'lbx', x0[0]); s.set(0, 'ubx', x0[0])
    st = s.solve()
    r = co.solve_batch(co.default_opts(co.MODEL_ATT), x0, y[None], pc, xg, ug)
    assert st == r['status'][0] == 0 and s.sqp_iter == r['sqp_iter'][0] and s.qp_iter == r['qp_iter'][0]
    assert s.sqp_iter >= 3
    assert np.abs(r['x'][0] - s.x).max() < 1e-10 and np.abs(r['u'][0] - s.u).max() < 1e-10 and np.abs(r['pi'][0] - s.pi).max() < 1e-10


def test_att_device_templates_on_the_host_match_the_oracle():
    """the product's solver templates instantiated for Model_att (tests/hostsim) against the C oracle"""
    import hostsim
    B = 8
    oc = co.default_opts(co.MODEL_ATT)
    refs, x0, pc, pp = att_inputs(B, seed=9)
    y = window(refs, 3)
    xg, ug = start_iterate(B)
    want = co.solve_batch(oc, x0, y, pp, xg, ug)
    got = hostsim.solve_batch(hostsim.MODEL_ATT, 0, hostsim.opts_from_oracle(oc), x0, y, pp, xg, ug)
    for k in ('status', 'sqp_iter', 'qp_iter'):
        assert np.array_equal(want[k], got[k]), k
    assert (want['status'] == 0).all() and want['qp_iter'].max() > want['qp_iter'].min()
    for k in ('x', 'u', 'pi'):
        assert np.abs(want[k] - got[k]).max() < 1e-10, k


def test_att_closed_loop_tracks_the_reference():
    """formulation check on the oracle: from an offset start, with a 5 % mass mismatch and noise, the drone converges to the
    helix and the quaternion stays a unit quaternion to integration accuracy"""
    B, S = 4, 120
    refs, x0, pc, pp = att_inputs(B, seed=2, rows=S + 30)
    noise = np.random.default_rng(0).normal(0, 0.002, (S, B))
    r = co.closed_loop_att(co.default_opts(co.MODEL_ATT), refs, x0, noise, pc, pp, S)
    assert (r['status'] == 0).all()
    err = np.abs(r['Xsim'][:, -20:, :3] - refs[:, S - 19:S + 1, :3]).max()
    assert err < 0.05, err
    qn = np.linalg.norm(r['Xsim'][:, :, 6:10], axis=2)
    assert np.abs(qn - 1).max() < 1e-3


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
def _solver(B, **kw):
    import torch
    import drone_attitude_control_b200 as pkg
    s = pkg.BatchedAcadosOcpSolver('att', batch=B, device=0, **kw)
    assert (s.nx, s.nu, s.ny, s.N) == (10, 4, 14, 30)
    return s, torch


@pytest.mark.gpu
def test_att_single_solves_match_oracle():
    B = 48
    refs, x0, pc, pp = att_inputs(B, seed=11, spread=0.08)
    y = window(refs, 5)
    xg, ug = start_iterate(B)
    want = co.solve_batch(co.default_opts(co.MODEL_ATT), x0, y, pp, xg, ug)
    s, torch = _solver(B)
    dev = lambda a: torch.tensor(np.ascontiguousarray(a), device='cuda')
    s.set(0, 'p', dev(pp))
    for k in range(31):
        s.set(k, 'x', dev(xg[:, k]))
    for k in range(30):
        s.set(k, 'u', dev(ug[:, k]))
    s.set_yref_all(dev(y))
    s.set(0, 'lbx', dev(x0)); s.set(0, 'ubx', dev(x0))
    st = s.solve()
    assert np.array_equal(st.cpu().numpy(), want["status"]) and (want["status"] == 0).mean() > 0.9
    assert np.array_equal(s.get_stats('sqp_iter').cpu().numpy(), want['sqp_iter'])
    assert np.array_equal(s.get_stats('qp_iter').cpu().numpy(), want['qp_iter'])
    assert want['sqp_iter'].max() >= 3 and want['qp_iter'].max() > want['qp_iter'].min()
    for k in range(30):
        np.testing.assert_allclose(s.get(k, 'u').cpu().numpy(), want['u'][:, k], rtol=0, atol=1e-9)
        np.testing.assert_allclose(s.get(k, 'pi').cpu().numpy(), want['pi'][:, k], rtol=0, atol=1e-9)
    for k in range(31):
        np.testing.assert_allclose(s.get(k, 'x').cpu().numpy(), want['x'][:, k], rtol=0, atol=1e-9)
    u0 = s.get(0, 'u').cpu().numpy()
    assert (np.abs(u0 - want['u'][:, 0]) <= 1e-6 * np.maximum(np.abs(want['u'][:, 0]), 1e-3)).all()      # north-star: u0 within 1e-6 relative


@pytest.mark.gpu
def test_att_closed_loop_matches_oracle():
    from drone_attitude_control_b200.attitude_model import follow_trajectory_batched
    B, S = 24, 25
    refs, x0, pc, pp = att_inputs(B, seed=13, rows=S + 30)
    noise = np.random.default_rng(4).normal(0, 0.002, (S, B))
    want = co.closed_loop_att(co.default_opts(co.MODEL_ATT), refs, x0, noise, pc, pp, S)
    got = follow_trajectory_batched(refs, x0, S, noise=noise, p_ctrl=pc, p_plant=pp)
    assert np.array_equal(got['status'].cpu().numpy(), want['status']) and (want['status'] == 0).all()
    assert np.array_equal(got['sqp_iter'].cpu().numpy(), want['sqp_iter'])
    assert np.array_equal(got['qp_iter'].cpu().numpy(), want['qp_iter'])
    np.testing.assert_allclose(got['Xsim'].cpu().numpy(), want['Xsim'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(got['U_ctrl'].cpu().numpy(), want['U_ctrl'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(got['cost'].cpu().numpy(), want['cost'], rtol=1e-9)


@pytest.mark.gpu
def test_att_closed_loop_sqp_rti_matches_oracle():
    """the same closed loop with ONE QP per control step (acados SQP_RTI, the north-star's wording): status, iteration counts
    and trajectory against the oracle run with rti"""
    from drone_attitude_control_b200.attitude_model import follow_trajectory_batched
    B, S = 24, 25
    refs, x0, pc, pp = att_inputs(B, seed=13, rows=S + 30)
    noise = np.random.default_rng(4).normal(0, 0.002, (S, B))
    want = co.closed_loop_att(co.default_opts(co.MODEL_ATT, rti=True), refs, x0, noise, pc, pp, S)
    got = follow_trajectory_batched(refs, x0, S, noise=noise, p_ctrl=pc, p_plant=pp, rti=True)
    assert np.array_equal(got['status'].cpu().numpy(), want['status']) and (want['status'] == 0).all()
    assert np.array_equal(got['sqp_iter'].cpu().numpy(), want['sqp_iter']) and want['sqp_iter'].max() == 1
    assert np.array_equal(got['qp_iter'].cpu().numpy(), want['qp_iter'])
    np.testing.assert_allclose(got['Xsim'].cpu().numpy(), want['Xsim'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(got['U_ctrl'].cpu().numpy(), want['U_ctrl'], rtol=0, atol=1e-9)
    assert np.abs(want['Xsim'][:, -1, :3] - refs[:, S, :3]).max() < 0.05          # and it tracks


@pytest.mark.gpu
def test_att_sim_solver_matches_oracle():
    import drone_attitude_control_b200 as pkg
    import torch
    B = 33
    rng = np.random.default_rng(6)
    x = rng.normal(size=(B, 10)) * 0.3; x[:, 6] += 1.0
    u = np.hstack([rng.uniform(0.1, 0.6, (B, 1)), rng.uniform(-3, 3, (B, 3))])
    p = np.tile([M, G], (B, 1)); p[:, 0] *= rng.uniform(0.9, 1.1, B)
    sim = pkg.BatchedAcadosSimSolver(T=0.02, num_stages=4, batch=B, device=0, model='att', numpy_io=True)
    got = sim.simulate(x=x, u=u, p=p)
    want = co.sim_batch_model(co.MODEL_ATT, x, u[:, None, :], p, 4, 1, 0.02)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-14)


@pytest.mark.gpu
def test_att_reference_style_loop_single_drone():
    """the reference's follow_trajectory call sequence (set_up_ocp, lbx / ubx, solve, get, simulate_next_x) for one drone"""
    from drone_attitude_control_b200.attitude_model import follow_trajectory, gen_helix_traj
    S = 12
    xref, uref = gen_helix_traj(radius=0.8)
    x0 = xref[0].copy(); x0[:3] += [0.04, -0.03, 0.05]
    cost, Xsim, U = follow_trajectory(xref, uref, x0, noise=False, n_steps=S)
    ref = np.hstack([xref, uref])
    want = co.closed_loop_att(co.default_opts(co.MODEL_ATT), ref, x0[None], None, np.array([[M, G]]), np.array([[M, G]]), S)
    np.testing.assert_allclose(Xsim, want['Xsim'][0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(U, want['U_ctrl'][0], rtol=0, atol=1e-9)
    assert abs(cost - want['cost'][0]) <= 1e-9 * want['cost'][0]


@pytest.mark.gpu
def test_att_fp32_tracks_fp64():
    from drone_attitude_control_b200.attitude_model import follow_trajectory_batched
    B, S = 16, 40
    refs, x0, pc, pp = att_inputs(B, seed=17, rows=S + 30)
    a = follow_trajectory_batched(refs, x0, S, p_ctrl=pc, p_plant=pp)
    b = follow_trajectory_batched(refs, x0, S, p_ctrl=pc, p_plant=pp, precision='fp32')
    assert (a['status'] == 0).all() and (b['status'] == 0).all()
    dp = (a['Xsim'][:, :, :3] - b['Xsim'][:, :, :3]).abs().max().item()
    assert dp < 1e-3, dp          # stated FP32 tolerance: positions within 1e-3 m of the FP64 closed loop


@pytest.mark.gpu
def test_att_has_no_fused_loop_and_reports_it():
    import ctypes as C
    import drone_attitude_control_b200 as pkg
    from drone_attitude_control_b200._lib import ClosedLoopArgs, lib
    s, torch = _solver(2)
    a = ClosedLoopArgs()
    ref = torch.zeros((100, 8), dtype=torch.float64, device='cuda')
    a.n_steps, a.ref_rows, a.ref_shared, a.ref = 1, 100, 1, ref.data_ptr()
    assert lib().bnmpc_closed_loop_run(s.handle, C.byref(a)) == -5          # BNMPC_E_UNSUPPORTED
    assert b'step_for_x0' in lib().bnmpc_last_error()
    assert isinstance(pkg.BnmpcError('x'), RuntimeError)
