"""`OCP` and `Converter` of the jerk model with the reference's method names (reference src/jerk_model/ocp.py:13-123,
src/jerk_model/dynamics.py:55-83)."""
import numpy as np

from ..acados_shim import BatchedAcadosOcpSolver, BatchedAcadosSimSolver
from ..params import DroneData, ExperimentParameters

p = ExperimentParameters()
dd = DroneData()


class Converter:
    """jerk h -> 10 plant inputs (theta, Fd) while integrating a_i; reference src/jerk_model/dynamics.py:59-83.
    Like the reference, a_i is advanced in place and returned."""

    def convert(self, h, a_i):
        u = np.zeros((p.ctrls_per_sample, 2))
        for j in range(p.ctrls_per_sample):
            a_i += h * p.dt_conv
            F_x = dd.MASS * a_i[0]
            F_z = dd.MASS * a_i[1]
            u[j, 0] = np.arctan2(F_x, F_z)
            u[j, 1] = np.sqrt(F_x * F_x + F_z * F_z)
        return u, a_i


class OCP:
    MODEL = 'jerk'

    def __init__(self, ocp_name='acados_ocp', batch=1, device=0, precision='fp64', **solver_overrides):
        self.ocp_name = ocp_name
        self.ocp = None
        self.ocp_solver = None
        self.sim = None
        self.integrator = None
        self._batch, self._device, self._precision, self._overrides = batch, device, precision, solver_overrides

    def create_ocp(self, model=None, **ocp_overrides):
        """The OCP of the reference's create_ocp; `ocp_overrides` replace its numbers (fields of bnmpc_config: W, W_e, lbx, ubx,
        lbu, ubu, horizon, dt, tol, ...), e.g. create_ocp(ubu=[0.3, 0.4]).  `model` is accepted for signature compatibility:
        the dynamics are the generated device code of MODEL."""
        self._overrides.update(ocp_overrides)
        # LINEAR_LS, w_x = [100,100,1,1,0,0], w_u = .1, boxes on jerk and on (p, v, a) (jerk_model/ocp.py:27-79)
        self.ocp = dict(model=self.MODEL)

    def create_ocp_solver(self):
        # PARTIAL_CONDENSING_HPIPM, GAUSS_NEWTON, ERK with sim_method_num_stages = 1 (explicit Euler), SQP (:84-92)
        self.ocp_solver = BatchedAcadosOcpSolver(self.MODEL, batch=self._batch, device=self._device, precision=self._precision,
                                                 N_horizon=p.N_horizon, dt=p.dt, **self._overrides)

    def create_simulator(self, model=None):
        # AcadosSim: T = dt_conv, ERK, num_stages = 1 (:97-104)
        self.integrator = BatchedAcadosSimSolver(T=p.dt_conv, num_stages=1, batch=self._batch, device=self._device)

    def simulate_next_x(self, x0, u, noise):
        x_i = x0
        for i in range(p.ctrls_per_sample):
            self.integrator.set('u', u[i])
            self.integrator.set('x', x_i)
            self.integrator.solve()
            x_i = self.integrator.get('x')
        eps = np.random.normal(0, p.noise) if noise else 0
        return x_i + eps

    def set_up_ocp(self, iteration, xref, uref):
        for k in range(p.N_horizon):
            self.ocp_solver.set(k, 'yref', np.hstack((xref[iteration + k], uref[iteration + k])))
        self.ocp_solver.set(p.N_horizon, 'yref', xref[iteration + p.N_horizon])


def follow_trajectory(xref, uref, x0, noise, verbose=False, device=0):
    """follow_trajectory of the reference (src/jerk_model/controller.py:8-58) - same arguments, same return value (closedLoopCost, Xsim
    [N+1, 4], a [N, 2], U_opt_plant [N, 2]) - run as ONE drone of the fused device-resident loop (BatchedClosedLoop): per
    control step one kernel does set_up_ocp, the x0 embedding, solve(), Converter.convert, simulate_next_x and the logged cost.
    `noise` draws np.random.normal(0, p.noise) once per control step from numpy's global stream, in the reference's order.
    A non-zero solver status raises, as the reference does."""
    import torch
    from ..closed_loop import BatchedClosedLoop
    ref = np.zeros((xref.shape[0], 8))
    ref[:, :6] = xref[:, :6]
    ref[:, 6:6 + uref.shape[1]] = uref
    eps = np.array([np.random.normal(0, p.noise) if noise else 0.0 for _ in range(p.N)])
    loop = BatchedClosedLoop(OCP.MODEL, batch=1, device=device)
    loop.init(torch.tensor(np.asarray(x0, float)[:, None]), torch.tensor(ref), noise=torch.tensor(eps[:, None]), n_steps=p.N)
    loop.run(steps_per_launch=p.N)
    r = loop.results()
    status = r['status'][0].cpu().numpy()
    if status.any():
        k = int(np.flatnonzero(status)[0])
        raise Exception(f'Failed in iteration {k}\nbnmpc ocp_solver returned status {status[k]}')
    if verbose:
        loop.solver.print_statistics()
    return float(r['cost'][0]), r['Xsim'][0].cpu().numpy(), r['a'][0].cpu().numpy(), r['U_plant'][0].cpu().numpy()
