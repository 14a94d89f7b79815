"""Metrics of reference src/store_results.py that main.py prints (plotting is out of scope)."""
import numpy as np
import torch


def calc_aed(pref, psim):
    """reference src/store_results.py:233-236 - despite its name, the mean absolute per-coordinate error"""
    if isinstance(pref, torch.Tensor):
        return torch.mean(torch.sqrt((pref - psim) ** 2))
    euclidean_distances = np.sqrt((pref - psim) ** 2)
    return np.mean(euclidean_distances)


def store_data(path_prefix, Xsim, a, U_opt_plant):
    """reference src/store_results.py:11-18: three .npy dumps"""
    np.save(path_prefix + '_xsim.npy', np.asarray(Xsim))
    np.save(path_prefix + '_a.npy', np.asarray(a))
    np.save(path_prefix + '_uopt.npy', np.asarray(U_opt_plant))


def export_instance(results, ref, instance=0, dt=1 / 50):
    """One drone of a batched run in the layout reference src/store_results.py:215-230 (`create_plots(dt, XRef, XSim, a,
    UOpt)`) and src/main.py:21-22 use: (dt, XRef [S, >=4], XSim [S, 4], a [S, 2], UOpt [S, 2]) as numpy arrays.
    `results` is BatchedClosedLoop.results(); `ref` the trajectory table of that drone ([rows, 8]) or of all ([B, rows, 8])."""
    to_np = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    xsim = to_np(results['Xsim'][instance])
    S = xsim.shape[0] - 1
    r = to_np(ref)
    r = r[instance] if r.ndim == 3 else r
    return dt, r[:S], xsim[:S], to_np(results['a'][instance]), to_np(results['U_plant'][instance])


def store_instance(path_prefix, results, ref, instance=0):
    """`store_data` of the reference (src/store_results.py:11-18: XRef, XSim, UOpt as .npy) for one drone of a batched
    run; returns the three paths like the reference does."""
    _, xref, xsim, _, uopt = export_instance(results, ref, instance)
    paths = tuple(f'{path_prefix}_{name}.npy' for name in ('XRef', 'XSim', 'UOpt'))
    for path, arr in zip(paths, (xref, xsim, uopt)):
        np.save(path, arr)
    return paths


def batch_statistics(results, percentiles=(5, 50, 95)):
    """Monte-Carlo summary of a batched run: mean and percentiles of the per-drone closed-loop cost
    (controller.py:40-41,54) and of calc_aed per drone (store_results.py:233-236), solver status histogram (acados
    codes 0-4 over all drones and steps) and the mean / max interior-point iterations per control step."""
    to_np = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    cost, aed, status, qp = (to_np(results[k]) for k in ('cost', 'aed', 'status', 'qp_iter'))
    out = {'n': int(cost.shape[0])}
    for name, v in (('cost', cost), ('aed', aed)):
        out[name] = {'mean': float(v.mean()), **{f'p{q}': float(np.percentile(v, q)) for q in percentiles}, 'max': float(v.max())}
    out['status_hist'] = {int(c): int((status == c).sum()) for c in range(5)}
    out['qp_iter'] = {'mean': float(qp.mean()), 'max': int(qp.max())}
    return out
