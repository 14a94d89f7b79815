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
