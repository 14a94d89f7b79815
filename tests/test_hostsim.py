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
