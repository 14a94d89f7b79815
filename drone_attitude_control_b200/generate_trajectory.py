"""Reference trajectories.  gen_circle_traj mirrors reference src/generate_trajectory.py:7-28 (same signature, same
quirks: linspace(0, T, N) spacing, wrap-around tail, columns [px pz vx vz ax az+g 0 0]); gen_circle_traj_batched
produces one table per instance (radius, centre, phase per instance) in the batch-minor layout [rows][8][B] the fused
closed loop reads."""
import numpy as np
import torch

from .params import DroneData, ExperimentParameters

p = ExperimentParameters()
dd = DroneData()


def gen_circle_traj(N, N_horizon, nx, nu, center, radius):
    assert nx == 6 and nu == 2, 'the reference fills 8 columns (src/generate_trajectory.py:13-24)'
    ref = np.zeros((N + N_horizon, nx + nu))
    omega = 2 * np.pi / p.T
    i = np.linspace(0, p.T, N)
    ref[:N, 0] = center[0] + radius * np.cos(omega * i)
    ref[:N, 1] = center[1] + radius * np.sin(omega * i)
    ref[:N, 2] = -radius * omega * np.sin(omega * i)
    ref[:N, 3] = radius * omega * np.cos(omega * i)
    ref[:N, 4] = -radius * omega ** 2 * np.cos(omega * i)
    ref[:N, 5] = -radius * omega ** 2 * np.sin(omega * i) + dd.GRAVITY_ACC
    ref[N:] = ref[:N_horizon]
    return ref


def gen_circle_traj_batched(N, N_horizon, radius, center, phase, device='cpu'):
    """radius [B], center [B,2], phase [B] (torch, float64) -> ref [N+N_horizon, 8, B] on `device`."""
    radius = torch.as_tensor(radius, dtype=torch.float64, device=device)
    center = torch.as_tensor(center, dtype=torch.float64, device=device)
    phase = torch.as_tensor(phase, dtype=torch.float64, device=device)
    B = radius.shape[0]
    omega = 2 * np.pi / p.T
    t = torch.linspace(0, p.T, N, dtype=torch.float64, device=device)
    a = omega * t[:, None] + phase[None, :]                       # [N, B]
    ref = torch.zeros((N + N_horizon, 8, B), dtype=torch.float64, device=device)
    ref[:N, 0] = center[None, :, 0] + radius * torch.cos(a)
    ref[:N, 1] = center[None, :, 1] + radius * torch.sin(a)
    ref[:N, 2] = -radius * omega * torch.sin(a)
    ref[:N, 3] = radius * omega * torch.cos(a)
    ref[:N, 4] = -radius * omega ** 2 * torch.cos(a)
    ref[:N, 5] = -radius * omega ** 2 * torch.sin(a) + dd.GRAVITY_ACC
    ref[N:] = ref[:N_horizon]
    return ref.contiguous()


# ---- the other trajectory families of the reference (src/jerk_model/gen_trajectory.py), batched --------------------------
# The reference samples them at the converter rate and none of them is wired into its main.py; here they are sampled
# at the MPC rate, one table per instance, in the 8-column layout the closed loop reads ([px pz vx vz ax az+g hx hz],
# same as gen_circle_traj) and the instance-major layout [B, rows, 8] (`BatchedClosedLoop.init(ref=...)`).

def _table(B, rows, device):
    ref = torch.zeros((B, rows, 8), dtype=torch.float64, device=device)
    ref[:, :, 5] = dd.GRAVITY_ACC                                  # hover: the z acceleration column carries +g
    return ref


def gen_static_point_traj_batched(N, N_horizon, initial, device='cpu'):
    """Hold a point (reference src/jerk_model/gen_trajectory.py:43-51).  initial [B, 2] -> ref [B, N + N_horizon, 8]."""
    initial = torch.as_tensor(initial, dtype=torch.float64, device=device)
    ref = _table(initial.shape[0], N + N_horizon, device)
    ref[:, :, 0] = initial[:, None, 0]
    ref[:, :, 1] = initial[:, None, 1]
    return ref


def gen_straight_traj_batched(N, N_horizon, initial, length, fill_acc=True, device='cpu'):
    """Constant-jerk diagonal line (reference src/jerk_model/gen_trajectory.py:54-71, inputs :94-100): jerk =
    6 length / T^3 on both axes, p = initial + jerk t^3 / 6, v = jerk t^2 / 2.  initial [B, 2], length [B].
    The reference leaves the acceleration columns zero; fill_acc=True (default) writes the consistent a = jerk t and the
    jerk itself into the input columns, fill_acc=False reproduces the reference's table."""
    initial = torch.as_tensor(initial, dtype=torch.float64, device=device)
    length = torch.as_tensor(length, dtype=torch.float64, device=device)
    rows = N + N_horizon
    ref = _table(initial.shape[0], rows, device)
    t = torch.arange(rows, dtype=torch.float64, device=device) * p.dt
    jerk = 6 * length / p.T ** 3                                   # [B]
    cub = jerk[:, None] * t[None, :] ** 3 / 6
    ref[:, :, 0] = initial[:, None, 0] + cub
    ref[:, :, 1] = initial[:, None, 1] + cub
    ref[:, :, 2] = 0.5 * jerk[:, None] * t[None, :] ** 2
    ref[:, :, 3] = ref[:, :, 2]
    if fill_acc:
        ref[:, :, 4] = jerk[:, None] * t[None, :]
        ref[:, :, 5] = jerk[:, None] * t[None, :] + dd.GRAVITY_ACC
        ref[:, :, 6] = jerk[:, None]
        ref[:, :, 7] = jerk[:, None]
    return ref


def gen_square_traj_batched(N, N_horizon, initial, length, device='cpu'):
    """Square of side `length` in the x-z plane, N // 4 samples per side, positions only (reference
    src/jerk_model/gen_trajectory.py:8-40 walks the same corners in its first two columns); the tail wraps around like
    gen_circle_traj.  initial [B, 2], length [B].  The corners are velocity discontinuities: this family drives the
    input and state bounds active on purpose."""
    initial = torch.as_tensor(initial, dtype=torch.float64, device=device)
    length = torch.as_tensor(length, dtype=torch.float64, device=device)
    B = initial.shape[0]
    side = N // 4
    ref = _table(B, N + N_horizon, device)
    f = torch.arange(side, dtype=torch.float64, device=device) / side          # i / side
    x0, z0, L = initial[:, None, 0], initial[:, None, 1], length[:, None]
    one = torch.ones_like(f)[None, :]
    xs = torch.cat([x0 + L * f, (x0 + L) * one, x0 + L - L * f, x0 * one], dim=1)
    zs = torch.cat([z0 * one, z0 + L * f, (z0 + L) * one, z0 + L - L * f], dim=1)
    ref[:, :4 * side, 0] = xs
    ref[:, :4 * side, 1] = zs
    if 4 * side < N:                                                          # N not a multiple of 4: stay on the start corner
        ref[:, 4 * side:N, 0] = x0
        ref[:, 4 * side:N, 1] = z0
    ref[:, N:] = ref[:, :N_horizon]
    return ref
