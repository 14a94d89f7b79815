"""CPU-only check of the PRODUCT's solver templates (csrc/bnmpc_core.cuh, bnmpc_loop.cuh compiled for the host by
tests/hostsim, a test harness that is not part of the library) against the C oracle.  The GPU parity tests
(test_gpu_parity.py) repeat this through the C-ABI on the device."""
import numpy as np
import pytest

import hostsim as hs
from common import random_loop_inputs, random_solve_inputs
from oracle import c_oracle as co


@pytest.mark.parametrize('model', [0, 1, 2, 3])
def test_single_solves_match_oracle(model):
    om = model % 2
    oo = co.default_opts(om)
    x0, yref = random_solve_inputs(om, 5, seed=11 + model)
    p = np.repeat(np.array([[0.03277, 9.81]]), 5, 0)
    want = co.solve_batch(oo, x0, yref, p)
    got = hs.solve_batch(model, hs.FP64, hs.opts_from_oracle(oo), x0, yref, p)
    assert np.array_equal(got['status'], want['status'])
    assert np.array_equal(got['sqp_iter'], want['sqp_iter']) and np.array_equal(got['qp_iter'], want['qp_iter'])
    np.testing.assert_allclose(got['u'], want['u'], rtol=0, atol=1e-10)
    np.testing.assert_allclose(got['x'], want['x'], rtol=0, atol=1e-10)
    np.testing.assert_allclose(got['pi'], want['pi'], rtol=0, atol=1e-9)


@pytest.mark.parametrize('model', [0, 1])
def test_closed_loop_matches_oracle(model):
    S, B = 25, 3
    refs, x0, noise, pc, pp = random_loop_inputs(B, S, seed=3 + model, mass_sigma=0.05)
    oo = co.default_opts(model)
    want = co.closed_loop(oo, refs, x0, noise, pc, pp, S)
    got = hs.closed_loop(model, hs.FP64, hs.opts_from_oracle(oo), refs, x0, noise, pc, pp, S, instance_major=(model == 0))
    assert np.array_equal(got['status'], want['status']) and np.array_equal(got['qp_iter'], want['qp_iter'])
    for k in ('Xsim', 'U_ctrl', 'U_plant', 'a'):
        np.testing.assert_allclose(got[k], want[k], rtol=0, atol=1e-10, err_msg=k)
    np.testing.assert_allclose(got['cost'], want['cost'], rtol=1e-12)


def test_on_the_fly_circle_reference():
    """ref_shared = 3: the trajectory rows are computed in the kernel from (radius, centre, phase); same result as the
    materialised table, which in turn is gen_circle_traj of the reference (src/generate_trajectory.py:7-28)."""
    from oracle import nmpc_oracle as o
    rng = np.random.default_rng(8)
    B, S = 3, 12
    prm = np.stack([rng.uniform(0.5, 1.0, B), rng.uniform(-0.15, 0.15, B), rng.uniform(-0.15, 0.15, B), rng.uniform(0, 6.28, B)], 1)
    tab = hs.circle_table(prm, 530, 500)
    for i in range(B):
        want = o.gen_circle_traj(center=prm[i, 1:3], radius=prm[i, 0], phase=prm[i, 3])
        np.testing.assert_allclose(tab[i], want, rtol=0, atol=5e-15)
    np.testing.assert_array_equal(hs.circle_table(np.array([[1.0, 0, 0, 0]]), 530, 500)[0, 0], [1, 0, -0.0, 2 * np.pi / 10, -(2 * np.pi / 10) ** 2, 9.81, 0, 0])
    x0 = tab[:, 0, :4] + rng.uniform(-0.05, 0.05, (B, 4)); noise = rng.normal(0, 0.01, (S, B))
    p = np.repeat(np.array([[0.03277, 9.81]]), B, 0)
    oo = hs.opts_from_oracle(co.default_opts(1))
    a = hs.closed_loop(1, hs.FP64, oo, tab, x0, noise, p, p, S, instance_major=True)
    b = hs.closed_loop(1, hs.FP64, oo, prm, x0, noise, p, p, S, circle_rows=530)
    for k in ('Xsim', 'U_ctrl', 'cost', 'qp_iter', 'status'):
        assert np.array_equal(a[k], b[k]), k


def test_nonlinear_thrust_ocp_matches_oracle():
    """General path of the product templates (one block, sensitivities stored per stage, several SQP iterations)."""
    from common import thrust_refs, thrust_solve_inputs
    oo = co.default_opts(co.MODEL_THRUST)
    x0, yref = thrust_solve_inputs(3, seed=5)
    p = np.repeat(np.array([[0.03277, 9.81]]), 3, 0)
    want = co.solve_batch(oo, x0, yref, p)
    got = hs.solve_batch(hs.MODEL_THRUST, hs.FP64, hs.opts_from_oracle(oo), x0, yref, p)
    assert np.array_equal(got['status'], want['status']) and np.array_equal(got['sqp_iter'], want['sqp_iter'])
    assert np.array_equal(got['qp_iter'], want['qp_iter']) and want['sqp_iter'].min() >= 2
    np.testing.assert_allclose(got['u'], want['u'], rtol=0, atol=1e-10)
    np.testing.assert_allclose(got['x'], want['x'], rtol=0, atol=1e-10)
    S, B = 10, 2
    refs, x0, noise, pc, pp = random_loop_inputs(B, S, seed=9, mass_sigma=0.05)
    refs = thrust_refs(refs)
    want = co.closed_loop(oo, refs, x0, noise, pc, pp, S)
    got = hs.closed_loop(hs.MODEL_THRUST, hs.FP64, hs.opts_from_oracle(oo), refs, x0, noise, pc, pp, S, instance_major=True)
    assert np.array_equal(got['status'], want['status']) and np.array_equal(got['qp_iter'], want['qp_iter'])
    for k in ('Xsim', 'U_ctrl', 'U_plant', 'a'):
        np.testing.assert_allclose(got[k], want[k], rtol=0, atol=1e-10, err_msg=k)


@pytest.mark.parametrize('model,chunk,W', [(0, 1, 3), (0, 4, 5), (1, 25, 2), (0, 7, 16), (hs.MODEL_THRUST, 3, 4), (0, 6, 0), (1, 1, 0), (hs.MODEL_THRUST, 5, 0)])
def test_lockstep_schedule_is_bit_identical(model, chunk, W):
    """Multi-step launches - queue tickets of `chunk` control steps with the working set kept on chip inside a chunk, run by
    free warps (W = 0: closed_loop_chunk, the path of k_loop_step) or by the slotted lockstep schedule (bnmpc_lockstep.cuh: W
    warps of a CTA side by side, sweeps in sweep slots) - only re-time the work of an instance: every output must equal the
    one-launch-per-step path bit for bit, for any number of warps and any chunk length."""
    from common import thrust_refs
    S, B = 25, 7
    om = co.MODEL_THRUST if model == hs.MODEL_THRUST else model
    refs, x0, noise, pc, pp = random_loop_inputs(B, S, seed=13 + model, mass_sigma=0.05)
    if model == hs.MODEL_THRUST:
        refs, S = thrust_refs(refs), 8
        noise = noise[:S]
    oo = hs.opts_from_oracle(co.default_opts(om))
    a = hs.closed_loop(model, hs.FP64, oo, refs, x0, noise, pc, pp, S, instance_major=True)
    b = hs.closed_loop(model, hs.FP64, oo, refs, x0, noise, pc, pp, S, instance_major=True, lockstep=(chunk, W))
    for k in ('Xsim', 'U_ctrl', 'U_plant', 'a', 'cost', 'abs_err', 'qp_iter', 'status'):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(b['failures'], (a['status'] != 0).sum(1))


def test_lockstep_rti_and_failures():
    """SQP_RTI through the lockstep schedule, and the failure path: a NaN in the reference gives status 1 for the steps
    whose window sees it, the loop keeps going, the instance's failure count says how often, and the solver state an
    instance carries across a failed step is what the per-step path carries (multipliers of the last good solve)."""
    S, B = 12, 4
    refs, x0, noise, pc, pp = random_loop_inputs(B, S, seed=31)
    oo = co.default_opts(0, rti=True)
    want = co.closed_loop(oo, refs, x0, noise, pc, pp, S)
    got = hs.closed_loop(0, hs.FP64, hs.opts_from_oracle(oo), refs, x0, noise, pc, pp, S, instance_major=True, lockstep=(5, 3))
    assert np.array_equal(got['status'], want['status']) and np.array_equal(got['qp_iter'], want['qp_iter'])
    np.testing.assert_allclose(got['Xsim'], want['Xsim'], rtol=0, atol=1e-10)
    refs = refs.copy()
    refs[2, 30 + 3, 1] = np.nan                    # enters the window of instance 2 at step 3 (terminal stage) and stays until step 33
    oo = hs.opts_from_oracle(co.default_opts(0))
    a = hs.closed_loop(0, hs.FP64, oo, refs, x0, noise, pc, pp, S, instance_major=True)
    b = hs.closed_loop(0, hs.FP64, oo, refs, x0, noise, pc, pp, S, instance_major=True, lockstep=(4, 3))
    assert (a['status'][2, 3:] == 1).all() and (a['status'][2, :3] == 0).all() and (a['status'][[0, 1, 3]] == 0).all()
    for k in ('U_ctrl', 'qp_iter', 'status'):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(np.nan_to_num(a['Xsim'], nan=-7.0), np.nan_to_num(b['Xsim'], nan=-7.0))
    assert b['failures'].tolist() == [0, 0, S - 3, 0]
    # QP failure (status 4): a start state far outside the position box makes the first QPs infeasible
    refs, x0, noise, pc, pp = random_loop_inputs(B, S, seed=32)
    x0[1, 0] += 4.0
    a = hs.closed_loop(0, hs.FP64, oo, refs, x0, noise, pc, pp, S, instance_major=True)
    assert (a['status'][1] == 4).any() and (a['status'][[0, 2, 3]] == 0).all()
    for sched in ((6, 3), (5, 0)):
        b = hs.closed_loop(0, hs.FP64, oo, refs, x0, noise, pc, pp, S, instance_major=True, lockstep=sched)
        for k in ('Xsim', 'U_ctrl', 'qp_iter', 'status', 'cost'):
            assert np.array_equal(a[k], b[k]), (sched, k)
        assert b['failures'][1] == (a['status'][1] != 0).sum()


def test_philox_noise_generator():
    """Known-answer test of Philox4x32-10 (Random123 kat_vectors: counter 0 / key 0 and the all-ones vector) and the
    statistics / sharding-independence of the normal draws built on it."""
    assert hs.philox4x32([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert hs.philox4x32([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    z = hs.philox_noise(4096, 50, seed=2026, std=1.0)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01 and np.abs(z).max() < 6.5
    part = hs.philox_noise(100, 50, seed=2026, std=1.0, first_instance=1000)
    assert np.array_equal(part, z[:, 1000:1100])
    later = hs.philox_noise(4096, 10, seed=2026, std=1.0, first_step=40)
    assert np.array_equal(later, z[40:])
    S, B = 10, 5
    refs, x0, _, pc, pp = random_loop_inputs(B, S, seed=2)
    oo = hs.opts_from_oracle(co.default_opts(0))
    nz = hs.philox_noise(B, S, seed=99, std=0.01, first_instance=12)
    a = hs.closed_loop(0, hs.FP64, oo, refs, x0, nz, pc, pp, S, instance_major=True, lockstep=(3, 2))
    b = hs.closed_loop(0, hs.FP64, oo, refs, x0, None, pc, pp, S, instance_major=True, lockstep=(3, 2), philox=(99, 0.01, 12))
    assert np.array_equal(a['Xsim'], b['Xsim']) and np.array_equal(a['U_ctrl'], b['U_ctrl'])


@pytest.mark.parametrize('model', [0, hs.MODEL_THRUST])
def test_irk_integrator_in_the_product_templates(model):
    """erk_stages = 0: the product's irk_gl4_step (Gauss-Legendre collocation, Newton, IFT sensitivities; acados IRK of
    reference src/force_model/ocp.py:85) inside linearise / the constant-Jacobian binding, against the C oracle's."""
    from common import thrust_solve_inputs
    om = co.MODEL_THRUST if model == hs.MODEL_THRUST else 0
    x0, yref = thrust_solve_inputs(3, seed=15) if om else random_solve_inputs(0, 3, seed=14)
    p = np.repeat(np.array([[0.03277, 9.81]]), 3, 0)
    oo = co.default_opts(om, erk_stages=0)
    want = co.solve_batch(oo, x0, yref, p)
    erk = co.solve_batch(co.default_opts(om), x0, yref, p)
    got = hs.solve_batch(model, hs.FP64, hs.opts_from_oracle(oo), x0, yref, p)
    for ref in (want, erk):
        assert np.array_equal(got['status'], ref['status']) and np.array_equal(got['sqp_iter'], ref['sqp_iter'])
        assert np.array_equal(got['qp_iter'], ref['qp_iter'])
        np.testing.assert_allclose(got['u'], ref['u'], rtol=0, atol=1e-10)
        np.testing.assert_allclose(got['x'], ref['x'], rtol=0, atol=1e-10)


def _stage_bound_case(seed=21, N=30):
    """a force-model solve whose input box is tightened on stages 3..9 and whose vx box is tightened on stages 5..12"""
    from oracle import nmpc_oracle as o
    x0, yref = random_solve_inputs(0, 2, seed=seed)
    spec = o.force_ocp()
    bnd = np.zeros((2, N, 2, 6))
    bnd[:, :, 0, :2], bnd[:, :, 1, :2] = spec.lbu, spec.ubu
    bnd[:, :, 0, 2:], bnd[:, :, 1, 2:] = spec.lbx, spec.ubx
    bnd[:, 3:10, 1, :2] = 0.30                 # ubu
    bnd[:, 3:10, 0, 1] = 0.25                  # lbu of F_z
    bnd[0, 5:13, 1, 4] = 0.2; bnd[0, 5:13, 0, 4] = -0.2      # |vx| <= 0.2 for instance 0
    want = []
    for i in range(2):
        s = o.OracleOcpSolver(spec)
        for k in range(N):
            s.set(k, 'yref', yref[i, k * 6:(k + 1) * 6])
            s.set(k, 'lbu', bnd[i, k, 0, :2]); s.set(k, 'ubu', bnd[i, k, 1, :2])
            if k >= 1:
                s.set(k, 'lbx', bnd[i, k, 0, 2:]); s.set(k, 'ubx', bnd[i, k, 1, 2:])
        s.set(N, 'yref', yref[i, 180:])
        s.set(0, 'lbx', x0[i]); s.set(0, 'ubx', x0[i])
        st = s.solve()
        want.append(dict(status=st, u=s.u.copy(), x=s.x.copy(), qp_iter=s.qp_iter, sqp_iter=s.sqp_iter))
    return x0, yref, bnd, want


def test_per_stage_bounds_match_numpy_oracle():
    """'lbu' / 'ubu' / 'lbx' / 'ubx' set per stage (acados' ocp_solver.set at any stage): the product templates with the
    per-stage lookup compiled in against the dense-KKT numpy oracle."""
    x0, yref, bnd, want = _stage_bound_case()
    p = np.repeat(np.array([[0.03277, 9.81]]), 2, 0)
    oo = hs.opts_from_oracle(co.default_opts(0))
    got = hs.solve_batch(0, hs.FP64, oo, x0, yref, p, bnd=bnd)
    free = hs.solve_batch(0, hs.FP64, oo, x0, yref, p)
    for i in range(2):
        assert got['status'][i] == want[i]['status'] == 0
        assert got['qp_iter'][i] == want[i]['qp_iter'] and got['sqp_iter'][i] == want[i]['sqp_iter']
        np.testing.assert_allclose(got['u'][i], want[i]['u'], rtol=0, atol=1e-9)
        np.testing.assert_allclose(got['x'][i], want[i]['x'], rtol=0, atol=1e-9)
        assert got['u'][i, 3:10].max() <= 0.30 + 1e-9 and got['u'][i, 3:10, 1].min() >= 0.25 - 1e-9
    assert np.abs(got['x'][0, 5:13, 2]).max() <= 0.2 + 1e-9
    assert np.abs(got['u'] - free['u']).max() > 1e-3          # the per-stage boxes changed the solution


def thrust_iterate_chain(solve_one_iteration, S=3, B=24, seed=19):
    """Walk the SQP iterates of the nonlinear thrust OCP along a closed loop as the ORACLE visits them and hand every iterate
    to `solve_one_iteration(xs, yref, p, x, u) -> dict(x, u, pi, qp_iter)` (one SQP iteration = SQP_RTI from that iterate):
    returns the worst deviation of a single iteration from identical iterates and the number of iterations compared.  Used by
    the host-emulation test below and by the GPU parity test: it separates the parity of ONE iteration (linearisation with
    per-stage sensitivities, QP, full step - must agree to 1e-9) from the amplification of round-off over the 20-100
    full-step iterations a cold solve of this OCP takes (which is why accumulated closed-loop states carry a looser bound)."""
    from common import random_loop_inputs, thrust_refs
    refs, x0, noise, pc, pp = random_loop_inputs(B, S, seed=seed, mass_sigma=0.05)
    refs = thrust_refs(refs)
    oo, orti = co.default_opts(co.MODEL_THRUST), co.default_opts(co.MODEL_THRUST, rti=True)
    N = 30
    x, u, xs = np.zeros((B, N + 1, 4)), np.zeros((B, N, 2)), x0.copy()
    worst, nit = 0.0, 0
    for i in range(S):
        yref = np.concatenate([refs[:, i:i + N, :6].reshape(B, -1), refs[:, i + N, :4]], 1)
        full = co.solve_batch(oo, xs, yref, pc, x=x, u=u)
        xi, ui = x.copy(), u.copy()
        for _ in range(min(int(full['sqp_iter'].max()), 40)):
            a = co.solve_batch(orti, xs, yref, pc, x=xi, u=ui)
            b = solve_one_iteration(xs, yref, pc, xi, ui)
            assert np.array_equal(a['qp_iter'], b['qp_iter'])
            worst = max(worst, np.abs(a['u'] - b['u']).max(), np.abs(a['x'] - b['x']).max(), np.abs(a['pi'] - b['pi']).max())
            nit += 1
            xi, ui = a['x'], a['u']
        x, u = full['x'], full['u']
        xs = co.sim_batch(xs, u[:, 0][:, None, :], pp, 4, 1, 0.02) + noise[i][:, None]
    return worst, nit


def test_nonlinear_ocp_one_iteration_from_identical_iterates():
    ho = hs.opts_from_oracle(co.default_opts(co.MODEL_THRUST, rti=True))
    worst, nit = thrust_iterate_chain(lambda xs, yref, p, x, u: hs.solve_batch(hs.MODEL_THRUST, hs.FP64, ho, xs, yref, p, x=x, u=u), S=2, B=8)
    assert nit >= 20 and worst < 1e-9, (worst, nit)
