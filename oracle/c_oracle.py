"""ctypes binding of oracle/nmpc_oracle.c (TEST INFRASTRUCTURE ONLY - see the header of that file)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libnmpc_oracle.so')
MODEL_FORCE, MODEL_JERK, MODEL_THRUST = 0, 1, 2      # THRUST: the plant model as controller model (our nonlinear extension)
MODEL_ATT = 3                                        # 3-D attitude-and-total-thrust model (north-star extension)
NXM, NUM, NSM = 12, 4, 16


class Opts(C.Structure):
    _fields_ = [('model', C.c_int), ('N', C.c_int), ('erk_stages', C.c_int), ('sqp_max_iter', C.c_int),
                ('qp_max_iter', C.c_int), ('rti', C.c_int), ('dt', C.c_double),
                ('w', C.c_double * NSM), ('w_e', C.c_double * NXM),
                ('lbx', C.c_double * NXM), ('ubx', C.c_double * NXM), ('lbu', C.c_double * NUM), ('ubu', C.c_double * NUM),
                ('tol', C.c_double * 4), ('qp_tol', C.c_double * 4),
                ('mu0', C.c_double), ('thr0', C.c_double), ('alpha_min', C.c_double), ('lam_min', C.c_double),
                ('t_min', C.c_double)]


def build(force=False):
    src = os.path.join(_HERE, 'nmpc_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE] + (['-B'] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        assert _lib.orc_sizeof_opts() == C.sizeof(Opts)
    return _lib


def load_variant(path):
    """another build of the same source (bench.py: -march=native with compile-time dimensions, for timing)"""
    L = C.CDLL(path)
    assert L.orc_sizeof_opts() == C.sizeof(Opts)
    return L


def default_opts(model, N=30, rti=False, **kw):
    o = Opts()
    lib().orc_default_opts(C.c_int(model), C.byref(o))
    o.N = N
    o.rti = int(rti)
    for k, v in kw.items():
        if k in ('tol', 'qp_tol'):
            for i in range(4):
                getattr(o, k)[i] = v
        else:
            setattr(o, k, v)
    return o


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def dims(model):
    return (10, 4) if model == MODEL_ATT else ((6, 2) if model == MODEL_JERK else (4, 2))


def solve_batch(o, x0, yref, p, x=None, u=None, nthreads=0):
    """One SQP solve per instance.  x0 [B,nx], yref [B,N*ny+nx], p [B,2]; x [B,N+1,nx], u [B,N,nu] = start iterate."""
    nx, nu = dims(o.model)
    B, N = x0.shape[0], o.N
    x0 = np.ascontiguousarray(x0, float); yref = np.ascontiguousarray(yref, float); p = np.ascontiguousarray(p, float)
    x = np.zeros((B, N + 1, nx)) if x is None else np.array(x, float, order='C')
    u = np.zeros((B, N, nu)) if u is None else np.array(u, float, order='C')
    pi = np.zeros((B, N, nx)); lam = np.zeros((B, 2 * (N * nu + (N + 1) * nx)))
    st = np.zeros(B, np.int32); si = np.zeros(B, np.int32); qi = np.zeros(B, np.int32)
    lib().orc_solve_batch(C.byref(o), B, _dp(x0), _dp(yref), _dp(p), _dp(x), _dp(u), _dp(pi), _dp(lam), _ip(st), _ip(si),
                          _ip(qi), nthreads)
    return dict(x=x, u=u, pi=pi, lam=lam, status=st, sqp_iter=si, qp_iter=qi)


def sim_batch(x, u, p, ns, nsub, T):
    B = x.shape[0]
    x = np.ascontiguousarray(x, float); u = np.ascontiguousarray(u, float); p = np.ascontiguousarray(p, float)
    xn = np.zeros((B, 4))
    lib().orc_sim_batch(B, ns, nsub, C.c_double(T), _dp(x), _dp(u), _dp(p), _dp(xn))
    return xn


def sim_batch_model(model, x, u, p, ns, nsub, T):
    """plant step of any model as its own plant (MODEL_ATT: x [B,10], u [B,nsub,4])"""
    nx, _ = dims(model)
    B = x.shape[0]
    x = np.ascontiguousarray(x, float); u = np.ascontiguousarray(u, float); p = np.ascontiguousarray(p, float)
    xn = np.zeros((B, nx))
    lib().orc_sim_batch_model(model, B, ns, nsub, C.c_double(T), _dp(x), _dp(u), _dp(p), _dp(xn))
    return xn


def closed_loop_att(o, ref, x0, noise, p_ctrl, p_plant, n_steps, nthreads=0):
    """3-D attitude model: ref [B,rows,14] or [rows,14] (shared); x0 [B,10]; noise [n_steps,B] or None; p_* [B,2]."""
    B = x0.shape[0]
    ref = np.ascontiguousarray(ref, float)
    shared = ref.ndim == 2
    rows = ref.shape[-2]
    x0 = np.ascontiguousarray(x0, float)
    noise = None if noise is None else np.ascontiguousarray(noise, float)
    p_ctrl = np.ascontiguousarray(p_ctrl, float); p_plant = np.ascontiguousarray(p_plant, float)
    out = dict(cost=np.zeros(B), Xsim=np.zeros((B, n_steps + 1, 10)), U_ctrl=np.zeros((B, n_steps, 4)),
               status=np.zeros((B, n_steps), np.int32), qp_iter=np.zeros((B, n_steps), np.int32), sqp_iter=np.zeros((B, n_steps), np.int32))
    rc = lib().orc_closed_loop_att(C.byref(o), B, n_steps, rows, _dp(ref), int(shared), _dp(x0), _dp(noise), _dp(p_ctrl), _dp(p_plant),
                                   _dp(out['Xsim']), _dp(out['U_ctrl']), _dp(out['cost']), _ip(out['status']), _ip(out['qp_iter']),
                                   _ip(out['sqp_iter']), nthreads)
    assert rc == 0, rc
    return out


def closed_loop(o, ref, x0, noise, p_ctrl, p_plant, n_steps, nthreads=0, outputs=True, L=None):
    """ref [B,rows,8] or [rows,8] (shared); x0 [B,4]; noise [n_steps,B] or None; p_* [B,2]."""
    B = x0.shape[0]
    ref = np.ascontiguousarray(ref, float)
    shared = ref.ndim == 2
    rows = ref.shape[-2]
    x0 = np.ascontiguousarray(x0, float)
    noise = None if noise is None else np.ascontiguousarray(noise, float)
    p_ctrl = np.ascontiguousarray(p_ctrl, float); p_plant = np.ascontiguousarray(p_plant, float)
    out = dict(cost=np.zeros(B))
    if outputs:
        out.update(Xsim=np.zeros((B, n_steps + 1, 4)), U_plant=np.zeros((B, n_steps, 2)), U_ctrl=np.zeros((B, n_steps, 2)),
                   a=np.zeros((B, n_steps, 2)), status=np.zeros((B, n_steps), np.int32), qp_iter=np.zeros((B, n_steps), np.int32))
    rc = (L or lib()).orc_closed_loop(C.byref(o), B, n_steps, rows, _dp(ref), int(shared), _dp(x0), _dp(noise), _dp(p_ctrl),
                               _dp(p_plant), _dp(out.get('Xsim')), _dp(out.get('U_plant')), _dp(out.get('U_ctrl')),
                               _dp(out.get('a')), _dp(out['cost']), _ip(out.get('status')), _ip(out.get('qp_iter')), nthreads)
    assert rc == 0, rc
    return out
