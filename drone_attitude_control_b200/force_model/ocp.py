"""`OCP` and `Converter` of the force model with the reference's method names (reference src/force_model/ocp.py:13-122,
src/force_model/dynamics.py:50-79).  The acados objects are replaced by the libbnmpc shims; the OCP formulation
(create_ocp :21-78) and solver options (create_ocp_solver :80-96) are the library's default configuration of the
'force' model (bnmpc_config_default), so create_ocp only records the choice."""
import numpy as np

from ..acados_shim import BatchedAcadosOcpSolver, BatchedAcadosSimSolver
from ..params import DroneData, ExperimentParameters

p = ExperimentParameters()
dd = DroneData()


class Converter:
    """controller input (Fx, Fz) -> plant input (theta, Fd); reference src/force_model/dynamics.py:54-79"""

    def convert(self, F):
        F = np.asarray(F)
        if F.ndim == 1:                                   # single input, dynamics.py:66-70
            return np.array([np.arctan2(F[0], F[1]), np.sqrt(F[0] * F[0] + F[1] * F[1])])
        u = np.zeros_like(F)                              # trajectory of inputs, dynamics.py:72-78
        u[:, 0] = np.arctan2(F[:, 0], F[:, 1])
        u[:, 1] = np.sqrt(F[:, 0] ** 2 + F[:, 1] ** 2)
        return u


class OCP:
    MODEL = 'force'

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
        # LINEAR_LS cost, W = blkdiag(diag(100,100,1,1), diag(.1,.1)), W_e, Vx/Vu selection, BGH boxes on u and x, x0 = 0
        self.ocp = dict(model=self.MODEL)

    def create_ocp_solver(self):
        # PARTIAL_CONDENSING_HPIPM, GAUSS_NEWTON, IRK (served by ERK4: exact for this affine model), SQP, N_horizon, tf
        self.ocp_solver = BatchedAcadosOcpSolver(self.MODEL, batch=self._batch, device=self._device, precision=self._precision,
                                                 N_horizon=p.N_horizon, dt=p.dt, **self._overrides)

    def create_simulator(self, model=None):
        # AcadosSim: T = dt, ERK, num_stages = 4 (ocp.py:98-104)
        self.integrator = BatchedAcadosSimSolver(T=p.dt, num_stages=4, batch=self._batch, device=self._device)

    def simulate_next_x(self, x0, u, noise):
        self.integrator.set('u', u)
        self.integrator.set('x', x0)
        self.integrator.solve()
        x_next = self.integrator.get('x')
        eps = np.random.normal(0, p.noise) if noise else 0          # one scalar for all states (ocp.py:114-115)
        return x_next + eps

    def set_up_ocp(self, iter, xref, uref):
        for k in range(p.N_horizon):
            self.ocp_solver.set(k, 'yref', np.hstack((xref[iter + k], uref[iter + k])))
        self.ocp_solver.set(p.N_horizon, 'yref', xref[iter + p.N_horizon])


def follow_trajectory(xref, uref, x0, noise, verbose=False, device=0):
    """follow_trajectory of the reference (src/force_model/controller.py:8-56) - same arguments, same return value (closedLoopCost, Xsim
    [N+1, 4], a [N, 2], U_opt_plant [N, 2]) - run as ONE drone of the fused device-resident loop (BatchedClosedLoop): per
    control step one kernel does set_up_ocp, the x0 embedding, solve(), Converter.convert, simulate_next_x and the logged cost.
    `noise` draws np.random.normal(0, p.noise) once per control step from numpy's global stream, in the reference's order.
    A non-zero solver status raises, as the reference does."""
    import torch
    from ..closed_loop import BatchedClosedLoop
    ref = np.zeros((xref.shape[0], 8))
    ref[:, :4] = xref[:, :4]
    ref[:, 4:4 + uref.shape[1]] = uref
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
