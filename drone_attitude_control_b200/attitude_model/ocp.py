"""`OCP`, `Converter` and `follow_trajectory` for the 3-D attitude model, with the reference's method names and call order
(the shape of reference src/force_model/ocp.py:13-122 and src/force_model/controller.py:8-56).

Model (BNMPC_MODEL_ATT, generated device code - codegen/gen_models.py): x = (p [3], v [3], q [4] = attitude quaternion
body->world, w first), u = (T, wx, wy, wz) total thrust and body rates, p = (mass, g):
    pdot = v,  vdot = (T / m) R(q) e3 - g e3,  qdot = 1/2 q (x) (0, w)
OCP: LINEAR_LS on [x; u] (position 100, velocity 1, quaternion vector part 10, inputs 0.1), boxes on thrust and body rates at
every stage and on position / velocity at stages 1..N-1, ERK4, SQP + the HPIPM-style interior point of the library.  The plant
is the same model integrated with the PLANT parameters (mass perturbation = model mismatch), one ERK4 step per control step;
the noise draw is added to position and velocity.  The planar plant of reference src/plant.py:27-33 is this model restricted to
the x-z plane (theta = pitch, Fd = T)."""
import numpy as np

from ..acados_shim import BatchedAcadosOcpSolver, BatchedAcadosSimSolver
from ..params import DroneData, ExperimentParameters

p = ExperimentParameters()
dd = DroneData()
NX, NU = 10, 4


def hover_input(mass=None, g=None):
    return np.array([(dd.MASS if mass is None else mass) * (dd.GRAVITY_ACC if g is None else g), 0.0, 0.0, 0.0])


def gen_helix_traj(n_steps=None, n_horizon=None, center=(0.0, 0.0, 0.0), radius=1.0, phase=0.0, y_amp=0.3):
    """Reference trajectory for the 3-D model: the circle of reference src/generate_trajectory.py:7-28 in the x-z plane (same
    linspace(0, T, n) sampling and wrap-around tail of n_horizon rows) plus a lateral sway y = y_amp sin(2 w t), identity
    attitude, hover input.  Returns xref [n_steps + n_horizon, 10], uref [n_steps + n_horizon, 4]."""
    n = p.N if n_steps is None else int(n_steps)
    nh = p.N_horizon if n_horizon is None else int(n_horizon)
    t = np.linspace(0, p.T, n)
    om = 2 * np.pi / p.T
    a = om * t + phase
    x = np.zeros((n, NX))
    x[:, 0] = center[0] + radius * np.cos(a); x[:, 1] = center[1] + y_amp * np.sin(2 * a); x[:, 2] = center[2] + radius * np.sin(a)
    x[:, 3] = -radius * om * np.sin(a); x[:, 4] = 2 * om * y_amp * np.cos(2 * a); x[:, 5] = radius * om * np.cos(a)
    x[:, 6] = 1.0
    u = np.tile(hover_input(), (n, 1))
    return np.vstack((x, x[:nh])), np.vstack((u, u[:nh]))


def helix_table(radius, center, phase, y_amp, rows, device=None, dt=None, mass=None, g=None):
    """Batched reference tables on the device (torch): [B, rows, 14] = [xref (10) | uref (4)] per row, the helix of
    gen_helix_traj sampled at the control rate (row i at t = i dt) for per-instance radius [B], centre [B, 3], phase [B],
    y_amp [B]."""
    import torch
    dt = p.dt if dt is None else dt
    f64 = dict(dtype=torch.float64, device=device)
    radius, center, phase, y_amp = (torch.as_tensor(a, **f64) for a in (radius, center, phase, y_amp))
    t = torch.arange(rows, **f64) * dt
    om = 2 * np.pi / p.T
    a = om * t[None, :] + phase[:, None]
    B = radius.shape[0]
    out = torch.zeros((B, rows, NX + NU), **f64)
    R, Y = radius[:, None], y_amp[:, None]
    out[:, :, 0] = center[:, 0:1] + R * torch.cos(a); out[:, :, 1] = center[:, 1:2] + Y * torch.sin(2 * a); out[:, :, 2] = center[:, 2:3] + R * torch.sin(a)
    out[:, :, 3] = -R * om * torch.sin(a); out[:, :, 4] = 2 * om * Y * torch.cos(2 * a); out[:, :, 5] = R * om * torch.cos(a)
    out[:, :, 6] = 1.0
    out[:, :, 10] = float(hover_input(mass, g)[0])
    return out


class Converter:
    """The OCP input already is the plant input (total thrust, body rates): convert() is the identity (the counterpart of
    reference src/force_model/dynamics.py:54-79, which maps (Fx, Fz) to (theta, Fd))."""

    def convert(self, u):
        return np.asarray(u)


class OCP:
    MODEL = 'att'

    def __init__(self, ocp_name='acados_ocp', batch=1, device=0, precision='fp64', **solver_overrides):
        self.ocp_name = ocp_name
        self.ocp = None
        self.ocp_solver = None
        self.integrator = None
        self._batch, self._device, self._precision, self._overrides = batch, device, precision, solver_overrides

    def create_ocp(self, model=None, **ocp_overrides):
        """`ocp_overrides` replace the numbers of the default OCP (fields of bnmpc_config: W, W_e, lbx, ubx, lbu, ubu, ...)."""
        self._overrides.update(ocp_overrides)
        self.ocp = dict(model=self.MODEL)

    def create_ocp_solver(self):
        self.ocp_solver = BatchedAcadosOcpSolver(self.MODEL, batch=self._batch, device=self._device, precision=self._precision,
                                                 N_horizon=p.N_horizon, dt=p.dt, **self._overrides)
        self.initial_guess()

    def initial_guess(self, mass=None, g=None):
        """identity attitude and hover thrust at every stage (a zero quaternion is not an attitude)"""
        s = self.ocp_solver
        x = np.zeros((s.batch, NX)); x[:, 6] = 1.0
        u = np.tile(hover_input(mass, g), (s.batch, 1))
        io, s.numpy_io = s.numpy_io, True
        try:
            for k in range(s.N + 1):
                s.set(k, 'x', x if s.batch > 1 else x[0])
            for k in range(s.N):
                s.set(k, 'u', u if s.batch > 1 else u[0])
        finally:
            s.numpy_io = io

    def create_simulator(self, model=None):
        self.integrator = BatchedAcadosSimSolver(T=p.dt, num_stages=4, batch=self._batch, device=self._device, model='att')

    def simulate_next_x(self, x0, u, noise):
        self.integrator.set('u', u)
        self.integrator.set('x', x0)
        self.integrator.solve()
        x_next = np.array(self.integrator.get('x'))
        eps = np.random.normal(0, p.noise) if noise else 0          # one scalar per step, on position and velocity
        x_next[..., :6] += eps
        return x_next

    def set_up_ocp(self, iter, xref, uref):
        for k in range(p.N_horizon):
            self.ocp_solver.set(k, 'yref', np.hstack((xref[iter + k], uref[iter + k])))
        self.ocp_solver.set(p.N_horizon, 'yref', xref[iter + p.N_horizon])


def follow_trajectory(xref, uref, x0, noise, verbose=False, device=0, n_steps=None):
    """The reference's follow_trajectory loop (src/force_model/controller.py:18-56) for ONE drone of the 3-D model, call for
    call: set_up_ocp, x0 embedding through set(0, 'lbx' / 'ubx'), solve(), get(0, 'u'), convert, simulate_next_x.
    Returns (closedLoopCost, Xsim [n+1, 10], U_opt [n, 4])."""
    n = p.N if n_steps is None else int(n_steps)
    ocp = OCP(device=device)
    ocp.create_ocp()
    ocp.create_ocp_solver()
    ocp.create_simulator()
    conv = Converter()
    Xsim = np.zeros((n + 1, NX)); U = np.zeros((n, NU))
    Xsim[0] = x0
    cost = 0.0
    wc = np.array([1e2, 1e2, 1e2, 1.0, 1.0, 1.0])
    for i in range(n):
        ocp.set_up_ocp(i, xref, uref)
        ocp.ocp_solver.set(0, 'lbx', Xsim[i])
        ocp.ocp_solver.set(0, 'ubx', Xsim[i])
        status = ocp.ocp_solver.solve()
        if status != 0:
            raise Exception(f'Failed in iteration {i}\nbnmpc ocp_solver returned status {status}')
        U[i] = ocp.ocp_solver.get(0, 'u')
        X_opt = ocp.ocp_solver.get(0, 'x')
        d = X_opt[:6] - xref[i, :6]
        cost += float(d @ (wc * d))
        Xsim[i + 1] = ocp.simulate_next_x(Xsim[i], conv.convert(U[i]), noise)
    if verbose:
        ocp.ocp_solver.print_statistics()
    return cost, Xsim, U


def follow_trajectory_batched(ref, x0, n_steps, noise=None, p_ctrl=None, p_plant=None, device=0, precision='fp64', log=True,
                              solver=None, **overrides):
    """Closed loop of `batch` drones, device-resident: per control step the yref windows are gathered on the device and ONE
    bnmpc_step_for_x0 call does the x0 embedding, solve(), get(0, 'u') and the plant step with the noise draw.
    ref [B, rows, 14] or [rows, 14] (shared) = [xref (10) | uref (4)] per row, rows >= n_steps + N; x0 [B, 10];
    noise [n_steps, B] or None; p_ctrl / p_plant [B, 2] = (mass, g) or None (nominal).  torch tensors or numpy arrays.
    Returns dict(Xsim [B, n+1, 10], U_ctrl [B, n, 4], status / qp_iter / sqp_iter [B, n], cost [B], solver)."""
    import torch
    dev = torch.device('cuda', device)
    T = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64).to(dev).contiguous()
    ref, x0, noise, p_ctrl, p_plant = T(ref), T(x0), T(noise), T(p_ctrl), T(p_plant)
    B = x0.shape[0]
    fresh = solver is None            # (a solver passed in continues from its iterate: a second leg of the same loop)
    s = solver if solver is not None else BatchedAcadosOcpSolver('att', batch=B, device=device, precision=precision,
                                                                 N_horizon=overrides.pop('N_horizon', p.N_horizon), numpy_io=False, **overrides)
    N, ny = s.N, NX + NU
    if ref.dim() == 2:
        ref = ref[None].expand(B, -1, -1)
    assert ref.shape[1] >= n_steps + N and ref.shape[2] == ny
    if p_ctrl is not None and fresh:
        s.set(0, 'p', p_ctrl)
    mass = p_ctrl[:, 0] if p_ctrl is not None else torch.full((B,), dd.MASS, dtype=torch.float64, device=dev)
    grav = p_ctrl[:, 1] if p_ctrl is not None else torch.full((B,), dd.GRAVITY_ACC, dtype=torch.float64, device=dev)
    xg = torch.zeros((B, NX), dtype=torch.float64, device=dev); xg[:, 6] = 1.0
    ug = torch.zeros((B, NU), dtype=torch.float64, device=dev); ug[:, 0] = mass * grav
    for k in range(N + 1 if fresh else 0):
        s.set(k, 'x', xg)
    for k in range(N if fresh else 0):
        s.set(k, 'u', ug)
    x = x0.clone()
    xn = torch.empty_like(x)
    u0 = torch.empty((B, NU), dtype=torch.float64, device=dev); up = torch.empty((B, 2), dtype=torch.float64, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    out = dict(cost=torch.zeros(B, dtype=torch.float64, device=dev))
    if log:
        out.update(Xsim=torch.zeros((B, n_steps + 1, NX), dtype=torch.float64, device=dev), U_ctrl=torch.zeros((B, n_steps, NU), dtype=torch.float64, device=dev),
                   status=torch.zeros((B, n_steps), dtype=torch.int32, device=dev), qp_iter=torch.zeros((B, n_steps), dtype=torch.int32, device=dev),
                   sqp_iter=torch.zeros((B, n_steps), dtype=torch.int32, device=dev))
        out['Xsim'][:, 0] = x
    wc = torch.tensor([1e2, 1e2, 1e2, 1.0, 1.0, 1.0], dtype=torch.float64, device=dev)
    for i in range(n_steps):
        yref = torch.cat((ref[:, i:i + N, :].reshape(B, N * ny), ref[:, i + N, :NX]), 1).contiguous()       # OCP.set_up_ocp
        s.set_yref_all(yref)
        s.step_device(x, None if noise is None else noise[i].contiguous(), u0, up, st, xn, p_plant)
        d = s.get(0, 'x')[:, :6] - ref[:, i, :6]
        out['cost'] += (d * d * wc).sum(1)
        if log:
            out['U_ctrl'][:, i] = u0; out['status'][:, i] = st
            out['qp_iter'][:, i] = s.get_stats('qp_iter'); out['sqp_iter'][:, i] = s.get_stats('sqp_iter')
            out['Xsim'][:, i + 1] = xn
        x, xn = xn, x
    out['solver'] = s
    return out
