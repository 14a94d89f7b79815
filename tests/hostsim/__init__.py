"""ctypes binding of tests/hostsim/hostsim.cpp (TEST HARNESS ONLY): the product's solver templates compiled for the
host so that the CPU-only test run can check the device code's logic against the oracle.  Not part of the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libhostsim.so')
_SRC = [os.path.join(_HERE, 'hostsim.cpp')] + [
    os.path.join(_HERE, '..', '..', 'drone_attitude_control_b200', 'csrc', f)
    for f in ('bnmpc_core.cuh', 'bnmpc_loop.cuh', 'bnmpc_lockstep.cuh', 'generated/models_gen.cuh')]

MODEL_FORCE, MODEL_JERK, MODEL_FORCE_DENSE, MODEL_JERK_DENSE, MODEL_THRUST, MODEL_ATT = 0, 1, 2, 3, 4, 5
FP64, FP32 = 0, 1


class Opts(C.Structure):
    _fields_ = [('N', C.c_int), ('erk_stages', C.c_int), ('sqp_max_iter', C.c_int), ('qp_max_iter', C.c_int), ('rti', C.c_int),
                ('sim_erk_stages', C.c_int), ('sim_substeps', C.c_int), ('smem_stride', C.c_int), ('order', C.c_void_p),
                ('dt', C.c_double), ('sim_dt', C.c_double),
                ('W', C.c_double * 16), ('W_e', C.c_double * 12), ('lbx', C.c_double * 12), ('ubx', C.c_double * 12),
                ('lbu', C.c_double * 4), ('ubu', C.c_double * 4), ('tol', C.c_double * 4), ('qp_tol', C.c_double * 4),
                ('mu0', C.c_double), ('thr0', C.c_double), ('alpha_min', C.c_double), ('lam_min', C.c_double), ('t_min', C.c_double)]


def build(force=False):
    newest = max(os.path.getmtime(f) for f in _SRC)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        extra = ['-DHS_ONLY_ATT'] if os.environ.get('BNMPC_HOSTSIM_ATT_ONLY') else []
        subprocess.check_call(['g++', '-O1', '-std=c++17', '-fPIC', '-shared', '-pthread', '-mfma', '-o', _SO, _SRC[0]] + extra)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.hs_sizeof_opts() == C.sizeof(Opts)
    return _lib


def opts_from_oracle(oo):
    """oracle.c_oracle.Opts -> product Opts (same numbers the C-ABI's bnmpc_config_default produces)."""
    o = Opts()
    jerk = oo.model == 1        # (oracle model 2 = thrust: force-like integrator settings)
    o.N, o.erk_stages, o.sqp_max_iter, o.qp_max_iter, o.rti = oo.N, oo.erk_stages, oo.sqp_max_iter, oo.qp_max_iter, oo.rti
    o.sim_erk_stages, o.sim_substeps = (1, 10) if jerk else (4, 1)
    o.dt, o.sim_dt = oo.dt, (1.0 / 500 if jerk else oo.dt)
    for i in range(16):
        o.W[i] = oo.w[i]
    for i in range(12):
        o.W_e[i], o.lbx[i], o.ubx[i] = oo.w_e[i], oo.lbx[i], oo.ubx[i]
    for i in range(4):
        o.lbu[i], o.ubu[i], o.tol[i], o.qp_tol[i] = oo.lbu[i], oo.ubu[i], oo.tol[i], oo.qp_tol[i]
    o.mu0, o.thr0, o.alpha_min, o.lam_min, o.t_min = oo.mu0, oo.thr0, oo.alpha_min, oo.lam_min, oo.t_min
    return o


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def solve_batch(model, prec, o, x0, yref, p, x=None, u=None, bnd=None):
    """bnd: per-stage bounds [B, N, 2, nu + nx] (lower / upper, per stage [u; x]) or None = the boxes of `o`"""
    nx, nu = (10, 4) if model == MODEL_ATT else ((6, 2) if model in (1, 3) else (4, 2))
    B, N = x0.shape[0], o.N
    x0 = np.ascontiguousarray(x0, float); yref = np.ascontiguousarray(yref, float); p = np.ascontiguousarray(p, float)
    x = np.zeros((B, N + 1, nx)) if x is None else np.array(x, float, order='C')
    u = np.zeros((B, N, nu)) if u is None else np.array(u, float, order='C')
    pi = np.zeros((B, N, nx))
    st = np.zeros(B, np.int32); si = np.zeros(B, np.int32); qi = np.zeros(B, np.int32)
    bnd = None if bnd is None else np.ascontiguousarray(bnd, float)
    rc = lib().hs_solve_batch(model, prec, C.byref(o), B, _dp(x0), _dp(yref), _dp(p), _dp(x), _dp(u), _dp(pi), _ip(st), _ip(si), _ip(qi),
                              _dp(bnd))
    assert rc == 0, rc
    return dict(x=x, u=u, pi=pi, status=st, sqp_iter=si, qp_iter=qi)


def circle_table(params, rows, n):
    params = np.ascontiguousarray(params, float)
    out = np.zeros((params.shape[0], rows, 8))
    lib().hs_circle_table(params.shape[0], rows, n, _dp(params), _dp(out))
    return out


def closed_loop(model, prec, o, ref, x0, noise, p_ctrl, p_plant, n_steps, instance_major=False, circle_rows=None, lockstep=None,
                philox=None):
    """lockstep = (chunk, W): run through the slotted lockstep schedule (bnmpc_lockstep.cuh) with one emulated CTA of W warps
    and queue tickets of `chunk` control steps; philox = (seed, std, first_instance): noise drawn by the device-side generator."""
    return _closed_loop(model, prec, o, ref, x0, noise, p_ctrl, p_plant, n_steps, instance_major, circle_rows, lockstep, philox)


def philox_noise(B, n_steps, seed, std, first_instance=0, first_step=0):
    out = np.zeros((n_steps, B))
    lib().hs_philox_noise(B, n_steps, first_step, C.c_ulonglong(seed), C.c_double(std), C.c_longlong(first_instance), _dp(out))
    return out


def philox4x32(ctr, key):
    c = (C.c_uint * 4)(*ctr); k = (C.c_uint * 2)(*key); o = (C.c_uint * 4)()
    lib().hs_philox4x32(c, k, o)
    return list(o)


def _closed_loop(model, prec, o, ref, x0, noise, p_ctrl, p_plant, n_steps, instance_major, circle_rows, lockstep, philox):
    """ref [rows,8] shared or [B,rows,8]; x0 [B,4]; noise [n_steps,B]; p_* [B,2] (AoS like the oracle); converted to the
    layouts of the C-ABI (per-instance ref tables batch-minor [rows,8,B], or left instance-major).  Returns oracle-shaped
    arrays."""
    B = x0.shape[0]
    if circle_rows is not None:                       # ref = circle parameters [B, 4], rows of the virtual table
        shared, rows, refd = 3, int(circle_rows), np.ascontiguousarray(ref, float)
    else:
        shared = 1 if ref.ndim == 2 else (2 if instance_major else 0)
        rows = ref.shape[-2]
        refd = np.ascontiguousarray(ref if shared else np.transpose(ref, (1, 2, 0)), float)
    x0t = np.ascontiguousarray(x0.T, float)
    nz = None if noise is None else np.ascontiguousarray(noise, float)
    pc = np.ascontiguousarray(p_ctrl.T, float); pp = np.ascontiguousarray(p_plant.T, float)
    Xsim = np.zeros((n_steps + 1, 4, B)); Up = np.zeros((n_steps, 2, B)); Uc = np.zeros((n_steps, 2, B)); al = np.zeros((n_steps, 2, B))
    cost = np.zeros(B); ae = np.zeros(B); st = np.zeros((n_steps, B), np.int32); qi = np.zeros((n_steps, B), np.int32)
    fails = np.zeros(B, np.int32)
    if lockstep is None:
        assert philox is None
        rc = lib().hs_closed_loop(model, prec, C.byref(o), B, n_steps, rows, _dp(refd), int(shared), _dp(x0t), _dp(nz), _dp(pc), _dp(pp),
                                  _dp(Xsim), _dp(Up), _dp(Uc), _dp(al), _dp(cost), _dp(ae), _ip(st), _ip(qi))
    else:
        chunk, W = lockstep
        seed, std, i0 = philox if philox is not None else (0, 0.0, 0)
        rc = lib().hs_closed_loop_ls(model, prec, C.byref(o), B, n_steps, int(chunk), int(W), rows, _dp(refd), int(shared), _dp(x0t),
                                     _dp(nz), _dp(pc), _dp(pp), _dp(Xsim), _dp(Up), _dp(Uc), _dp(al), _dp(cost), _dp(ae), _ip(st),
                                     _ip(qi), _ip(fails), C.c_ulonglong(seed), C.c_double(std), C.c_longlong(i0))
    assert rc == 0, rc
    return dict(failures=fails, cost=cost, abs_err=ae, Xsim=np.transpose(Xsim, (2, 0, 1)), U_plant=np.transpose(Up, (2, 0, 1)),
                U_ctrl=np.transpose(Uc, (2, 0, 1)), a=np.transpose(al, (2, 0, 1)), status=st.T.copy(), qp_iter=qi.T.copy())
