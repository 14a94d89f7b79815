"""main.py of the reference (src/main.py:10-46) on libbnmpc: circle reference, force then jerk model, seed 42 noise."""
import numpy as np

from . import force_model, jerk_model
from .generate_trajectory import gen_circle_traj
from .params import ExperimentParameters
from .store_results import calc_aed


def main(x0, force=True, jerk=True, noise=True, verbose=False, device=0):
    p = ExperimentParameters()
    ref = gen_circle_traj(p.N, p.N_horizon, nx=6, nu=2, center=[0, 0], radius=1)
    out = {}
    if force:
        print('fly circle with force model')
        cost, xsim, a, uopt = force_model.follow_trajectory(ref[:, :4], ref[:, 4:6], x0, noise, verbose, device=device)
        aed = calc_aed(ref[:p.N, :2], xsim[:p.N, :2])
        print(f'FORCE: Total cost: {np.round(cost, 2)}, AvgEucDist: {aed}')
        out['force'] = dict(cost=cost, Xsim=xsim, a=a, U_opt_plant=uopt, aed=aed)
    if jerk:
        print('fly circle with jerk model')
        cost, xsim, a, uopt = jerk_model.follow_trajectory(ref[:, :6], ref[:, 6:], x0, noise, verbose, device=device)
        aed = calc_aed(ref[:p.N, :2], xsim[:p.N, :2])
        print(f'JERK: Total cost: {np.round(cost, 2)}, AvgEucDist: {aed}')
        out['jerk'] = dict(cost=cost, Xsim=xsim, a=a, U_opt_plant=uopt, aed=aed)
    return out


if __name__ == '__main__':
    np.random.seed(42)
    main(np.array([1.0, 0, 0, 0.62]))
