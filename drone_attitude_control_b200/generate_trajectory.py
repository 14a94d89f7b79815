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
