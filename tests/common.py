"""Shared test helpers: seeded problem generators in the oracle's (AoS) layout."""
import numpy as np

from oracle import nmpc_oracle as o

P_NOM = np.array([o.MASS, o.GRAVITY_ACC])


def random_solve_inputs(model, B, seed, spread=0.08, N=30):
    """Random circle references / start states for single solves; returns x0 [B,nx], yref [B,N*ny+nx]."""
    rng = np.random.default_rng(seed)
    jerk = model in (1, 3, 'jerk')
    x0s, yrefs = [], []
    for _ in range(B):
        ref = o.gen_circle_traj(n_horizon=max(N, 30), radius=rng.uniform(0.5, 1.0), center=rng.uniform(-0.15, 0.15, 2),
                                phase=rng.uniform(0, 2 * np.pi))
        st = int(rng.integers(0, 400))
        x0 = ref[st, :4] + rng.uniform(-spread, spread, 4)
        if jerk:
            x0 = np.hstack([x0, [rng.uniform(-1, 1), o.GRAVITY_ACC + rng.uniform(-1, 1)]])
            y = np.hstack([ref[st:st + N, :8].ravel(), ref[st + N, :6]])
        else:
            y = np.hstack([ref[st:st + N, :6].ravel(), ref[st + N, :4]])
        x0s.append(x0); yrefs.append(y)
    return np.array(x0s), np.array(yrefs)


def random_loop_inputs(B, S, seed, N=30, mass_sigma=0.0):
    """Config-2 style closed-loop inputs: per-instance circle (radius, centre, phase), x0 near the reference, noise.
    Returns ref [B,rows,8], x0 [B,4], noise [S,B], p_ctrl [B,2], p_plant [B,2]."""
    rng = np.random.default_rng(seed)
    refs = np.stack([o.gen_circle_traj(n_horizon=max(N, 30), radius=rng.uniform(0.5, 1.0), center=rng.uniform(-0.15, 0.15, 2),
                                       phase=rng.uniform(0, 2 * np.pi)) for _ in range(B)])
    x0 = refs[:, 0, :4] + rng.uniform(-0.05, 0.05, (B, 4))
    noise = rng.normal(0, o.NOISE_STD, (S, B))
    p_ctrl = np.repeat(P_NOM[None], B, 0)
    p_plant = p_ctrl.copy()
    if mass_sigma > 0:
        p_plant[:, 0] *= 1 + np.clip(rng.normal(0, mass_sigma, B), -0.15, 0.15)
    return refs, x0, noise, p_ctrl, p_plant


def thrust_refs(refs):
    """Reference tables for the thrust OCP (u = (theta, Fd)): columns 4, 5 hold the input reference (0, m g)."""
    r = np.array(refs, float)
    r[..., 4] = 0.0
    r[..., 5] = o.GRAVITY
    return r


def thrust_solve_inputs(B, seed, spread=0.08, N=30):
    rng = np.random.default_rng(seed)
    x0s, yrefs = [], []
    for _ in range(B):
        ref = thrust_refs(o.gen_circle_traj(n_horizon=max(N, 30), radius=rng.uniform(0.5, 1.0), center=rng.uniform(-0.15, 0.15, 2),
                                            phase=rng.uniform(0, 2 * np.pi)))
        st = int(rng.integers(0, 400))
        x0s.append(ref[st, :4] + rng.uniform(-spread, spread, 4))
        yrefs.append(np.hstack([ref[st:st + N, :6].ravel(), ref[st + N, :4]]))
    return np.array(x0s), np.array(yrefs)
