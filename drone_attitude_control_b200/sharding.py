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


_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11) on arrays of counters; the host mirror of philox4x32_10 in csrc/bnmpc_loop.cuh.
    Inputs are integer arrays / scalars (32-bit values), outputs four uint64 arrays holding 32-bit words."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _M32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = np.uint64(int(k0) & 0xFFFFFFFF), np.uint64(int(k1) & 0xFFFFFFFF)
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0, p1 = m0 * c0, m1 * c2
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ k0, p1 & _M32, (p0 >> np.uint64(32)) ^ c3 ^ k1, p0 & _M32
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & _M32, (k1 + np.uint64(0xBB67AE85)) & _M32
    return c0, c1, c2, c3


def _uniform53(hi, lo, open_left=False):
    """53-bit uniform from two 32-bit words: [0, 1), or (0, 1] with open_left (the argument of a logarithm)"""
    v = ((hi << np.uint64(32)) | lo) >> np.uint64(11)
    return (v + np.uint64(1 if open_left else 0)).astype(np.float64) * (1.0 / 9007199254740992.0)


def philox_uniforms(seed, gid, block, stream):
    """two uniforms in [0, 1) per element of `gid` from the Philox block with counter (gid, block, stream)"""
    gid = np.asarray(gid, dtype=np.uint64)
    r = philox4x32_10(gid & _M32, gid >> np.uint64(32), block, stream, int(seed) & 0xFFFFFFFF, int(seed) >> 32)
    return _uniform53(r[0], r[1]), _uniform53(r[2], r[3])


def philox_normal(seed, gid, step, stream=0):
    """N(0,1) of (seed, global instance, step): the host mirror of philox_normal in csrc/bnmpc_loop.cuh (Box-Muller on the two
    53-bit uniforms of the block with counter (instance, step, stream)); equal to the device draw up to the rounding of the
    device's log / cos."""
    gid = np.asarray(gid, dtype=np.uint64)
    r = philox4x32_10(gid & _M32, gid >> np.uint64(32), step, stream, int(seed) & 0xFFFFFFFF, int(seed) >> 32)
    u1, u2 = _uniform53(r[0], r[1], open_left=True), _uniform53(r[2], r[3])
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586476925 * u2)


def instance_inputs(lo, hi, n_steps, seed=2026, x0_spread=0.05, noise_std=0.01, mass_sigma=0.0, with_noise=True):
    """BASELINE config 2 / 4 inputs for global instances lo..hi-1 (float64, CPU tensors, batch-minor):
    radius [b] in [0.5, 1], center [b,2] in [-0.15, 0.15]^2, phase [b] in [0, 2 pi) of the circle reference, dx0 [4,b] offset
    of the start state from the reference (+-x0_spread), noise [n_steps,b] (None unless with_noise), mass_scale [b] (plant
    mass = 0.03277 * mass_scale, 1 + N(0, mass_sigma) clipped to +-15 %).
    Every number is a counter-based draw keyed by (seed, GLOBAL instance id[, step]) - Philox4x32-10, vectorised over the
    instances - so any sharding of the batch gives the same numbers, and the noise equals what the kernel draws itself when
    the loop is given a PhiloxNoise(seed, noise_std, first_instance=lo) instead of the array."""
    gid = np.arange(lo, hi, dtype=np.uint64)
    u_r, u_ph = philox_uniforms(seed, gid, 0, 1)
    u_cx, u_cz = philox_uniforms(seed, gid, 1, 1)
    u_a, u_b = philox_uniforms(seed, gid, 2, 1)
    u_c, u_d = philox_uniforms(seed, gid, 3, 1)
    radius = 0.5 + 0.5 * u_r
    center = np.stack([-0.15 + 0.3 * u_cx, -0.15 + 0.3 * u_cz], 1)
    phase = 2 * np.pi * u_ph
    dx0 = x0_spread * (2 * np.stack([u_a, u_b, u_c, u_d], 0) - 1)
    mass = np.ones(hi - lo)
    if mass_sigma > 0:
        mass = 1 + np.clip(mass_sigma * philox_normal(seed, gid, 4, stream=1), -0.15, 0.15)
    noise = None
    if with_noise:
        noise = noise_std * philox_normal(seed, gid[None, :], np.arange(n_steps, dtype=np.uint64)[:, None])
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    return dict(radius=t(radius), center=t(center), phase=t(phase), dx0=t(dx0), noise=t(noise), mass_scale=t(mass))


def reduce_metrics(values, op='sum', group=None):
    """all_reduce of a small float64 vector of per-rank metrics (the only collective of the whole job)."""
    t = torch.as_tensor(values, dtype=torch.float64).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == 'nccl':
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 'max' else dist.ReduceOp.SUM, group=group)
    return t.cpu()
