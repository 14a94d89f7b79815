# examples/reference_style - NOT part of the bnmpc package.
#
# This file transcribes the reference's own host loop (BroilerCompiler/drone-attitude-control, GPL-3.0, src/main.py:10-46) with
# the acados constructors swapped for the bnmpc shims, to show that the shims are a drop-in for that loop (same names, same
# call order, batch = 1, numpy in / out) and to reproduce the reference's committed run step by step
# (tests/test_gpu_parity.py::test_reference_style_main_reproduces_reference_run).  The product's own entry points for this
# path are drone_attitude_control_b200.force_model / jerk_model.follow_trajectory (fused, device-resident) and
# BatchedClosedLoop.
"""main.py of the reference (src/main.py:10-46) on libbnmpc: circle reference, force then jerk model, seed 42 noise."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import force_controller
import jerk_controller
from drone_attitude_control_b200.generate_trajectory import gen_circle_traj
from drone_attitude_control_b200.params import ExperimentParameters
from drone_attitude_control_b200.store_results import calc_aed


def main(x0, force=True, jerk=True, noise=True, verbose=False, device=0):
    p = ExperimentParameters()
    ref = gen_circle_traj(p.N, p.N_horizon, nx=6, nu=2, center=[0, 0], radius=1)
    out = {}
    if force:
        print('fly circle with force model')
        cost, xsim, a, uopt = force_controller.follow_trajectory(ref[:, :4], ref[:, 4:6], x0, noise, verbose, device=device)
        aed = calc_aed(ref[:p.N, :2], xsim[:p.N, :2])
        print(f'FORCE: Total cost: {np.round(cost, 2)}, AvgEucDist: {aed}')
        out['force'] = dict(cost=cost, Xsim=xsim, a=a, U_opt_plant=uopt, aed=aed)
    if jerk:
        print('fly circle with jerk model')
        cost, xsim, a, uopt = jerk_controller.follow_trajectory(ref[:, :6], ref[:, 6:], x0, noise, verbose, device=device)
        aed = calc_aed(ref[:p.N, :2], xsim[:p.N, :2])
        print(f'JERK: Total cost: {np.round(cost, 2)}, AvgEucDist: {aed}')
        out['jerk'] = dict(cost=cost, Xsim=xsim, a=a, U_opt_plant=uopt, aed=aed)
    return out


if __name__ == '__main__':
    np.random.seed(42)
    main(np.array([1.0, 0, 0, 0.62]))
