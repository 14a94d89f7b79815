"""Multi-GPU sharding of the batch: one process per GPU, each rank owns a contiguous slice of instances.

The OCP instances are independent, so there is NO collective on the data path: inputs are generated or
sliced per rank from per-instance seeds keyed by the GLOBAL instance id (results do not depend on the number of ranks),
the closed loop runs with zero inter-GPU traffic, and only the final metrics - a few scalars - are combined with one
all_reduce (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """Contiguous slice [lo, hi) of `batch` instances owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def instance_inputs(lo, hi, n_steps, seed=2026, x0_spread=0.05, noise_std=0.01, mass_sigma=0.0):
    """BASELINE config 2 / 4 inputs for global instances lo..hi-1 (float64, CPU tensors, batch-minor):
    radius [b], center [b,2], phase [b] of the circle reference, dx0 [4,b] offset of the start state from the
    reference, noise [n_steps,b], mass_scale [b] (plant mass = 0.03277 * mass_scale).
    Every instance draws from its own generator seeded by (seed, global id), so any sharding gives the same numbers."""
    b = hi - lo
    radius = torch.empty(b, dtype=torch.float64); center = torch.empty(b, 2, dtype=torch.float64)
    phase = torch.empty(b, dtype=torch.float64); dx0 = torch.empty(4, b, dtype=torch.float64)
    noise = torch.empty(n_steps, b, dtype=torch.float64); mass = torch.ones(b, dtype=torch.float64)
    g = torch.Generator()
    for j, gid in enumerate(range(lo, hi)):
        g.manual_seed(int(seed) * 1000003 + gid)
        u = torch.rand(8, generator=g, dtype=torch.float64)
        radius[j] = 0.5 + 0.5 * u[0]
        center[j] = -0.15 + 0.3 * u[1:3]
        phase[j] = 2 * np.pi * u[3]
        dx0[:, j] = x0_spread * (2 * u[4:8] - 1)
        z = torch.randn(n_steps + 1, generator=g, dtype=torch.float64)
        noise[:, j] = noise_std * z[:n_steps]
        if mass_sigma > 0:
            mass[j] = 1 + float(torch.clamp(mass_sigma * z[n_steps], -0.15, 0.15))
    return dict(radius=radius, center=center, phase=phase, dx0=dx0, noise=noise, mass_scale=mass)


def reduce_metrics(values, op='sum', group=None):
    """all_reduce of a small float64 vector of per-rank metrics (the only collective of the whole job)."""
    t = torch.as_tensor(values, dtype=torch.float64).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == 'nccl':
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 'max' else dist.ReduceOp.SUM, group=group)
    return t.cpu()
