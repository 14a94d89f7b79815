# examples/reference_style - NOT part of the bnmpc package.
#
# This file transcribes the reference's own host loop (BroilerCompiler/drone-attitude-control, GPL-3.0, src/jerk_model/controller.py:8-58) with
# the acados constructors swapped for the bnmpc shims, to show that the shims are a drop-in for that loop (same names, same
# call order, batch = 1, numpy in / out) and to reproduce the reference's committed run step by step
# (tests/test_gpu_parity.py::test_reference_style_main_reproduces_reference_run).  The product's own entry points for this
# path are drone_attitude_control_b200.force_model / jerk_model.follow_trajectory (fused, device-resident) and
# BatchedClosedLoop.
"""follow_trajectory of the jerk model, written like reference src/jerk_model/controller.py:8-58."""
import numpy as np

from drone_attitude_control_b200.params import DroneData, ExperimentParameters
from drone_attitude_control_b200.jerk_model.ocp import OCP, Converter


def follow_trajectory(xref, uref, x0, noise, verbose=True, device=0):
    p = ExperimentParameters()
    dd = DroneData()
    converter = Converter()
    ocp = OCP(device=device)
    ocp.create_ocp()
    ocp.create_ocp_solver()
    ocp.create_simulator()
    Xsim = np.zeros((p.N + 1, 4))
    U_opt_plant = np.zeros((p.N, 2))
    a = np.zeros((p.N, 2))

    closedLoopCost = 0
    a_i = [0, dd.GRAVITY_ACC]       # current acceleration (hover initially), controller.py:23
    Xsim[0] = x0

    for iteration in range(p.N):
        ocp.set_up_ocp(iteration, xref, uref)

        x0_bar = np.hstack((Xsim[iteration], a_i))
        ocp.ocp_solver.set(0, 'lbx', x0_bar)
        ocp.ocp_solver.set(0, 'ubx', x0_bar)
        status = ocp.ocp_solver.solve()
        if status != 0:
            ocp.ocp_solver.print_statistics()
            raise Exception(f'Failed in iteration {iteration}\nbnmpc ocp_solver returned status {status}')
        U_opt_ctrl = ocp.ocp_solver.get(0, 'u')
        X_opt = ocp.ocp_solver.get(1, 'x')
        d = X_opt[:4] - xref[iteration, :4]
        cost = d @ np.diag([1e2, 1e2, 1e0, 1e0]) @ d

        u_tmp, a_i = converter.convert(U_opt_ctrl, a_i)
        a[iteration] = a_i
        U_opt_plant[iteration] = u_tmp[-1]
        Xsim[iteration + 1] = ocp.simulate_next_x(Xsim[iteration], u_tmp, noise)

        if verbose:
            print(f'{iteration}: U_opt [h_x h_z]: {np.round(U_opt_ctrl, 2)} '
                  f'X: {np.round(np.hstack((Xsim[iteration], a_i)), 2)} C: {np.round(cost, 5)}')
        closedLoopCost += cost

    return closedLoopCost, Xsim, a, U_opt_plant
