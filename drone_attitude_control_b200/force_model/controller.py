"""follow_trajectory of the force model, written like reference src/force_model/controller.py:8-56 (same loop, same
calls) against the libbnmpc shims.  This is the step-by-step host path (one set/solve/get round trip per call); the
fused device-resident loop for many drones is drone_attitude_control_b200.closed_loop.BatchedClosedLoop."""
import numpy as np

from ..params import DroneData, ExperimentParameters
from .ocp import OCP, Converter


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
    Xsim[0] = x0

    for iteration in range(p.N):
        ocp.set_up_ocp(iteration, xref, uref)

        x0_bar = Xsim[iteration]
        ocp.ocp_solver.set(0, 'lbx', x0_bar)
        ocp.ocp_solver.set(0, 'ubx', x0_bar)
        status = ocp.ocp_solver.solve()
        if status != 0:
            ocp.ocp_solver.print_statistics()
            raise Exception(f'Failed in iteration {iteration}\nbnmpc ocp_solver returned status {status}')
        U_opt_ctrl = ocp.ocp_solver.get(0, 'u')
        a[iteration] = U_opt_ctrl / dd.MASS
        X_opt = ocp.ocp_solver.get(0, 'x')
        d = X_opt[:4] - xref[iteration, :4]
        cost = d @ np.diag([1e2, 1e2, 1e0, 1e0]) @ d

        U_opt_plant[iteration] = converter.convert(U_opt_ctrl)
        Xsim[iteration + 1, :] = ocp.simulate_next_x(Xsim[iteration, :], U_opt_plant[iteration, :], noise)

        if verbose:
            print(f'{iteration}: U_opt [theta F_d]: {np.round(U_opt_plant[iteration, :], 2)} '
                  f'X: {np.round(Xsim[iteration, :], 2)} C: {cost}')
        closedLoopCost += cost

    return closedLoopCost, Xsim, a, U_opt_plant
