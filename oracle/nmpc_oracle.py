"""CPU oracle (numpy, FP64) for the NMPC hot path of BroilerCompiler/drone-attitude-control.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as the
checker.  It restates, on the CPU and with dense linear algebra, what the reference computes per control step:

    follow_trajectory (reference src/force_model/controller.py:8-56, src/jerk_model/controller.py:8-58)
      -> OCP.set_up_ocp               (src/force_model/ocp.py:117-122, src/jerk_model/ocp.py:118-123)
      -> AcadosOcpSolver.solve()      (OCP built in src/force_model/ocp.py:21-96, src/jerk_model/ocp.py:20-95)
      -> Converter.convert            (src/force_model/dynamics.py:66-70, src/jerk_model/dynamics.py:76-83)
      -> OCP.simulate_next_x          (src/force_model/ocp.py:106-115, src/jerk_model/ocp.py:106-116)

The arithmetic of AcadosOcpSolver / AcadosSimSolver lives in third-party code that is NOT in /root/reference
and not installable here: acados (+ HPIPM, BLASFEO) driven through acados_template==0.1, and casadi==3.6.7 /
3.7.0 (reference requirements.txt:1-4, src/requirements.txt:1-4; the acados version itself is unpinned, API
evidence says >= v0.4.0).  This file therefore restates their *published algorithms*:

  * acados SQP (ocp_nlp_sqp): linearise, Gauss-Newton LINEAR_LS Hessian/gradient with the stage cost scaled by
    the interval length and the terminal cost unscaled, x0 equality eliminated from stage 0, full step, residual
    test against tol 1e-6, QP cold-started every call, primal iterate kept between calls.
  * HPIPM d_ocp_qp_ipm_solve (mode BALANCE with acados' overrides mu0=1, iter_max=50, alpha_min=1e-8, tolerances
    1e-6): infeasible-start primal-dual IPM with Mehrotra predictor-corrector, conditional centering step,
    single step length for primal and dual, the step-shortening rule alpha*((1-alpha)*0.99+alpha*0.9999999).
    The Newton systems are solved here by a DENSE symmetric-indefinite solve of the whole KKT matrix - on
    purpose: the C oracle (oracle/nmpc_oracle.c) and the CUDA product use a Riccati recursion, so agreement
    between the three is agreement between independent linear-algebra paths.
  * acados sim_erk: explicit Runge-Kutta (1, 2, 3 or 4 stages) with forward sensitivities.

Parity pin: the reference has no tests.  The only reference-produced numbers are the plots in
experiment_data/img/; tools/extract_golden.py decodes them into tests/golden/acados_{force,jerk}.npz and
tests/test_oracle_golden.py checks this oracle against them (jerk: whole 500-step closed loop; force: interior
steps tightly, active-bound steps at closed-loop level - see DESIGN.md "Parity contract").
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------------------------------------------------
# Constants: reference src/params.py:37-61 (DroneData) and :113-122 (ExperimentParameters)
# ----------------------------------------------------------------------------------------------------------------------
MASS = 0.03277                      # params.py:42
GRAVITY_ACC = 9.81                  # params.py:37
GRAVITY = GRAVITY_ACC * MASS        # params.py:45
MAX_F = 1.3 * GRAVITY               # params.py:46
MIN_F = -0.2 * GRAVITY              # params.py:47
P_LIM = 1.2                         # params.py:48-51
V_LIM = 1.0                         # params.py:52-55
A_LIM = 5.0                         # params.py:56-59  (a_z bounds are shifted by +g)
JERK_LIM = 5.0                      # params.py:60-61
T_END = 10                          # params.py:115
DT = 1 / 50                         # params.py:116
DT_CONV = 1 / 500                   # params.py:117
CTRLS_PER_SAMPLE = int(DT / DT_CONV)  # params.py:118  (== 10)
N_STEPS = int(T_END / DT)           # params.py:119  (== 500)
N_HORIZON = 30                      # params.py:121
NOISE_STD = 0.01                    # params.py:122


# ----------------------------------------------------------------------------------------------------------------------
# Models.  p = (mass, g) is the per-instance parameter vector (the reference bakes both in as constants).
# ----------------------------------------------------------------------------------------------------------------------
def f_force(x, u, p):
    """reference src/force_model/dynamics.py:32-37"""
    m, g = p
    return np.array([x[2], x[3], u[0] / m, u[1] / m - g])


def jac_force(x, u, p):
    m, g = p
    fx = np.zeros((4, 4)); fx[0, 2] = 1.0; fx[1, 3] = 1.0
    fu = np.zeros((4, 2)); fu[2, 0] = 1.0 / m; fu[3, 1] = 1.0 / m
    return fx, fu


def f_jerk(x, u, p):
    """reference src/jerk_model/dynamics.py:35-42"""
    m, g = p
    return np.array([x[2], x[3], x[4], x[5] - g, u[0], u[1]])


def jac_jerk(x, u, p):
    fx = np.zeros((6, 6)); fx[0, 2] = fx[1, 3] = fx[2, 4] = fx[3, 5] = 1.0
    fu = np.zeros((6, 2)); fu[4, 0] = fu[5, 1] = 1.0
    return fx, fu


def f_plant(x, u, p):
    """reference src/plant.py:27-33; u = (theta, Fd)"""
    m, g = p
    th, fd = u
    return np.array([x[2], x[3], fd * np.sin(th) / m, fd * np.cos(th) / m - g])


def jac_plant(x, u, p):
    m, g = p
    th, fd = u
    fx = np.zeros((4, 4)); fx[0, 2] = 1.0; fx[1, 3] = 1.0
    fu = np.zeros((4, 2))
    fu[2, 0] = fd * np.cos(th) / m; fu[2, 1] = np.sin(th) / m
    fu[3, 0] = -fd * np.sin(th) / m; fu[3, 1] = np.cos(th) / m
    return fx, fu


def f_att(x, u, p):
    """NOT in the reference: the 3-D attitude-and-total-thrust model of the north-star.  x = (p, v, q) with
    q = (w, x, y, z) the attitude quaternion body->world, u = (T, wx, wy, wz):  pdot = v, vdot = (T/m) R(q) e3 - g e3,
    qdot = 1/2 q (x) (0, w).  The planar plant (src/plant.py:27-33) is its restriction to the x-z plane."""
    m, g = p
    qw, qx, qy, qz = x[6:10]
    a = u[0] / m
    wx, wy, wz = u[1:4]
    return np.array([x[3], x[4], x[5],
                     2.0 * (qx * qz + qw * qy) * a, 2.0 * (qy * qz - qw * qx) * a, (1.0 - 2.0 * (qx * qx + qy * qy)) * a - g,
                     0.5 * (-qx * wx - qy * wy - qz * wz), 0.5 * (qw * wx + qy * wz - qz * wy),
                     0.5 * (qw * wy - qx * wz + qz * wx), 0.5 * (qw * wz + qx * wy - qy * wx)])


def jac_att(x, u, p):
    m, g = p
    qw, qx, qy, qz = x[6:10]
    a = u[0] / m
    wx, wy, wz = u[1:4]
    fx = np.zeros((10, 10)); fu = np.zeros((10, 4))
    fx[0, 3] = fx[1, 4] = fx[2, 5] = 1.0
    fx[3, 6:10] = [2 * qy * a, 2 * qz * a, 2 * qw * a, 2 * qx * a]
    fx[4, 6:10] = [-2 * qx * a, -2 * qw * a, 2 * qz * a, 2 * qy * a]
    fx[5, 7] = -4 * qx * a; fx[5, 8] = -4 * qy * a
    fx[6, 7:10] = [-0.5 * wx, -0.5 * wy, -0.5 * wz]
    fx[7, 6] = 0.5 * wx; fx[7, 8] = 0.5 * wz; fx[7, 9] = -0.5 * wy
    fx[8, 6] = 0.5 * wy; fx[8, 7] = -0.5 * wz; fx[8, 9] = 0.5 * wx
    fx[9, 6] = 0.5 * wz; fx[9, 7] = 0.5 * wy; fx[9, 8] = -0.5 * wx
    fu[3, 0] = 2.0 * (qx * qz + qw * qy) / m; fu[4, 0] = 2.0 * (qy * qz - qw * qx) / m; fu[5, 0] = (1.0 - 2.0 * (qx * qx + qy * qy)) / m
    fu[6, 1:4] = [-0.5 * qx, -0.5 * qy, -0.5 * qz]
    fu[7, 1:4] = [0.5 * qw, -0.5 * qz, 0.5 * qy]
    fu[8, 1:4] = [0.5 * qz, 0.5 * qw, -0.5 * qx]
    fu[9, 1:4] = [-0.5 * qy, 0.5 * qx, 0.5 * qw]
    return fx, fu


# Butcher tableaus of acados sim_erk for num_stages = 1, 2, 3, 4 (explicit Euler, midpoint, Kutta-3, classic RK4)
_ERK = {
    1: (np.array([[0.0]]), np.array([1.0])),
    2: (np.array([[0.0, 0.0], [0.5, 0.0]]), np.array([0.0, 1.0])),
    3: (np.array([[0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [-1.0, 2.0, 0.0]]), np.array([1 / 6, 2 / 3, 1 / 6])),
    4: (np.array([[0, 0, 0, 0], [0.5, 0, 0, 0], [0, 0.5, 0, 0], [0, 0, 1.0, 0]]), np.array([1 / 6, 1 / 3, 1 / 3, 1 / 6])),
}


# Gauss-Legendre collocation with 4 stages (order 8): the tableau of acados sim_irk with its default
# sim_method_num_stages = 4 (integrator_type 'IRK', reference src/force_model/ocp.py:85).  Nodes = roots of the shifted
# Legendre polynomial P4, a_ij = int_0^{c_i} l_j, b_j = int_0^1 l_j; computed with 50 digits (mpmath), rounded to double.
GL4_A = np.array([
    [0.086963711284363464343, -0.026604180084998793313, 0.012627462689404724515, -0.0035551496857956831569],
    [0.18811811749986807165, 0.16303628871563653566, -0.027880428602470895224, 0.0067355005945381555154],
    [0.16719192197418877317, 0.35395300603374396654, 0.16303628871563653566, -0.014190694931141142964],
    [0.17748257225452261184, 0.3134451147418683468, 0.35267675751627186463, 0.086963711284363464343]])
GL4_B = np.array([0.17392742256872692869, 0.32607257743127307131, 0.32607257743127307131, 0.17392742256872692869])
IRK_NEWTON_ITER = 3                 # acados sim_method_newton_iter default


def irk_gl4_step(f, jac, x, u, p, T, num_steps=1, sens=True):
    """acados sim_irk restated (parity unpinned beyond the affine case: acados is not installable here and the reference
    only integrates the affine force model with it, for which every Newton variant returns the exact discretisation):
    per step, the stage derivatives K_i solve K_i = f(x + h sum_j a_ij K_j, u); Newton from K = 0 with the exact Jacobian
    I - h (A (x) f_x) re-evaluated in each of the IRK_NEWTON_ITER iterations, dense LU; the forward sensitivities follow
    from the implicit function theorem at the final iterate: (I - h A (x) f_x) dK/dw = [f_x S_x, f_x S_u + f_u]."""
    nx, nu, ns = len(x), len(u), 4
    h = T / num_steps
    x = np.array(x, float)
    S = np.hstack([np.eye(nx), np.zeros((nx, nu))])
    for _ in range(num_steps):
        K = np.zeros((ns, nx))
        for it in range(IRK_NEWTON_ITER + (1 if sens else 0)):
            last = it == IRK_NEWTON_ITER
            G = np.eye(ns * nx)
            R = np.zeros((ns * nx, nx + nu if last else 1))
            for i in range(ns):
                xi = x + h * (GL4_A[i] @ K)
                fx, fu = jac(xi, u, p)
                for j in range(ns):
                    G[i * nx:(i + 1) * nx, j * nx:(j + 1) * nx] -= h * GL4_A[i, j] * fx
                if last:
                    R[i * nx:(i + 1) * nx] = fx @ S
                    R[i * nx:(i + 1) * nx, nx:] += fu
                else:
                    R[i * nx:(i + 1) * nx, 0] = -(K[i] - f(xi, u, p))
            sol = np.linalg.solve(G, R)
            if last:
                S = S + h * np.tensordot(GL4_B, sol.reshape(ns, nx, nx + nu), axes=1)
            else:
                K = K + sol[:, 0].reshape(ns, nx)
        x = x + h * (GL4_B @ K)
    return (x, S) if sens else x


def erk_step(f, jac, x, u, p, T, num_stages, num_steps=1, sens=True):
    """acados sim_erk: num_steps steps of an explicit RK scheme over [0, T] with forward sensitivities.
    num_stages = 0 selects the implicit Gauss-Legendre scheme irk_gl4_step instead (the `erk_stages = 0` of the C-ABI).

    Returns x_next and, if sens, S = [d x_next/d x, d x_next/d u]  (nx x (nx+nu)).
    """
    if num_stages == 0:
        return irk_gl4_step(f, jac, x, u, p, T, num_steps, sens)
    A, b = _ERK[num_stages]
    nx, nu = len(x), len(u)
    h = T / num_steps
    x = np.array(x, float)
    S = np.hstack([np.eye(nx), np.zeros((nx, nu))])
    for _ in range(num_steps):
        K = np.zeros((num_stages, nx))
        SK = np.zeros((num_stages, nx, nx + nu))
        for i in range(num_stages):
            xi = x + h * (A[i, :i] @ K[:i]) if i else x
            K[i] = f(xi, u, p)
            if sens:
                Si = S + h * np.tensordot(A[i, :i], SK[:i], axes=1) if i else S
                fx, fu = jac(xi, u, p)
                SK[i] = fx @ Si
                SK[i][:, nx:] += fu
        x = x + h * (b @ K)
        if sens:
            S = S + h * np.tensordot(b, SK, axes=1)
    return (x, S) if sens else x


# ----------------------------------------------------------------------------------------------------------------------
# OCP description (what create_ocp / create_ocp_solver fix)
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class OcpSpec:
    name: str
    nx: int
    nu: int
    f: callable
    jac: callable
    erk_stages: int                 # integrator of the OCP dynamics, one step per interval
    w: np.ndarray                   # diag of W  (ny = nx+nu; Vx=[I;0], Vu=[0;I])
    w_e: np.ndarray                 # diag of W_e (ny_e = nx; Vx_e = I)
    lbx: np.ndarray
    ubx: np.ndarray
    lbu: np.ndarray
    ubu: np.ndarray
    N: int = N_HORIZON
    dt: float = DT
    # solver options (acados defaults, see module docstring)
    tol: float = 1e-6               # nlp_solver_tol_{stat,eq,ineq,comp}
    qp_tol_stat: float = None       # QP tolerances; None -> same as tol (acados hands the NLP tolerances to HPIPM)
    qp_tol_eq: float = None
    qp_tol_ineq: float = None
    qp_tol_comp: float = None
    mu_aff_shrink: float = 1.0      # COMPUTE_MU_AFF_QP alpha factor (HPIPM has an "alpha *= 0.99" there, disabled; 1.0 reproduces the golden run)
    sqp_max_iter: int = 100         # nlp_solver_max_iter
    qp_max_iter: int = 50           # qp_solver_iter_max
    mu0: float = 1.0
    thr0: float = 0.1
    alpha_min: float = 1e-8
    lam_min: float = 1e-16
    t_min: float = 1e-16


def force_ocp(N=N_HORIZON, **kw):
    """reference src/force_model/ocp.py:21-96.  integrator_type IRK (Gauss-Legendre) is exact for this affine
    model and so is ERK4 (the solution is quadratic in t); we use ERK4."""
    return OcpSpec('force', 4, 2, f_force, jac_force, 4,
                   w=np.array([1e2, 1e2, 1.0, 1.0, 1e-1, 1e-1]), w_e=np.array([1e2, 1e2, 1.0, 1.0]),
                   lbx=np.array([-P_LIM, -P_LIM, -V_LIM, -V_LIM]), ubx=np.array([P_LIM, P_LIM, V_LIM, V_LIM]),
                   lbu=np.array([MIN_F, MIN_F]), ubu=np.array([MAX_F, MAX_F]), N=N, **kw)


def jerk_ocp(N=N_HORIZON, **kw):
    """reference src/jerk_model/ocp.py:20-95.  ERK with sim_method_num_stages = 1 == explicit Euler."""
    return OcpSpec('jerk', 6, 2, f_jerk, jac_jerk, 1,
                   w=np.array([1e2, 1e2, 1.0, 1.0, 0.0, 0.0, 1e-1, 1e-1]),
                   w_e=np.array([1e2, 1e2, 1.0, 1.0, 0.0, 0.0]),
                   lbx=np.array([-P_LIM, -P_LIM, -V_LIM, -V_LIM, -A_LIM, -A_LIM + GRAVITY_ACC]),
                   ubx=np.array([P_LIM, P_LIM, V_LIM, V_LIM, A_LIM, A_LIM + GRAVITY_ACC]),
                   lbu=np.array([-JERK_LIM, -JERK_LIM]), ubu=np.array([JERK_LIM, JERK_LIM]), N=N, **kw)


def thrust_ocp(N=N_HORIZON, **kw):
    """NOT in the reference: the plant model (src/plant.py:27-33, inputs theta, Fd) used directly as controller model.
    A nonlinear OCP that exercises the general SQP path (state/input dependent sensitivities, several SQP iterations)."""
    return OcpSpec('thrust', 4, 2, f_plant, jac_plant, 4,
                   w=np.array([1e2, 1e2, 1.0, 1.0, 1e-1, 1e-1]), w_e=np.array([1e2, 1e2, 1.0, 1.0]),
                   lbx=np.array([-P_LIM, -P_LIM, -V_LIM, -V_LIM]), ubx=np.array([P_LIM, P_LIM, V_LIM, V_LIM]),
                   lbu=np.array([-1.0, 0.05]), ubu=np.array([1.0, 0.6]), N=N, **kw)


def att_ocp(N=N_HORIZON, **kw):
    """NOT in the reference: the OCP of the 3-D attitude model (same numbers as bnmpc_config_default(BNMPC_MODEL_ATT)):
    position weight 100, velocity 1, quaternion vector part 10, inputs 0.1; |p| <= 1.2, |v| <= 1, thrust within [0.1, 2] m g,
    body rates within +-6 rad/s; ERK4."""
    GR = MASS * GRAVITY_ACC
    w = np.array([1e2, 1e2, 1e2, 1.0, 1.0, 1.0, 0.0, 10.0, 10.0, 10.0, 1e-1, 1e-1, 1e-1, 1e-1])
    ub = np.array([P_LIM, P_LIM, P_LIM, V_LIM, V_LIM, V_LIM, 1.5, 1.5, 1.5, 1.5])
    return OcpSpec('att', 10, 4, f_att, jac_att, 4, w=w, w_e=w[:10].copy(), lbx=-ub, ubx=ub,
                   lbu=np.array([0.1 * GR, -6.0, -6.0, -6.0]), ubu=np.array([2.0 * GR, 6.0, 6.0, 6.0]), N=N, **kw)


# ----------------------------------------------------------------------------------------------------------------------
# HPIPM-style IPM on the stage-structured QP, dense KKT solves
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class QpResult:
    dz: np.ndarray
    pi: np.ndarray
    lam_lb: np.ndarray
    lam_ub: np.ndarray
    t_lb: np.ndarray
    t_ub: np.ndarray
    iters: int
    status: int          # HPIPM: 0 success, 1 max iter, 2 min step, 3 NaN
    res: tuple = ()


def qp_ipm_dense(Hd, q, G, bvec, lb, ub, spec: OcpSpec):
    """min 0.5 z'diag(Hd)z + q'z  s.t.  G z + bvec = 0,  lb <= z <= ub   (all variables bounded unless +-inf)

    HPIPM conventions: residuals res_g = H z + q + G'pi - lam_lb + lam_ub, res_b = G z + bvec,
    res_d_lb = lb - z + t_lb, res_d_ub = z - ub + t_ub, res_m = lam*t.
    """
    nv, ne = len(q), len(bvec)
    bounded = np.isfinite(lb) & np.isfinite(ub)
    nb = int(bounded.sum())
    ib = np.where(bounded)[0]
    lbb, ubb = lb[ib], ub[ib]
    nc = 2 * nb
    # ---- INIT_VAR_OCP_QP, cold start
    z = np.zeros(nv)
    pi = np.zeros(ne)
    t_lb = z[ib] - lbb
    t_ub = ubb - z[ib]
    thr0, mu0 = spec.thr0, spec.mu0
    for j in range(nb):
        if t_lb[j] < thr0:
            if t_ub[j] < thr0:
                z[ib[j]] = 0.5 * (lbb[j] + ubb[j])
                t_lb[j] = thr0; t_ub[j] = thr0
            else:
                t_lb[j] = thr0
                z[ib[j]] = lbb[j] + thr0
        elif t_ub[j] < thr0:
            t_ub[j] = thr0
            z[ib[j]] = ubb[j] - thr0
    lam_lb = mu0 / t_lb
    lam_ub = mu0 / t_ub

    def residuals():
        rg = Hd * z + q + G.T @ pi
        rg[ib] += lam_ub - lam_lb
        rb = G @ z + bvec
        rd_lb = lbb - z[ib] + t_lb
        rd_ub = z[ib] - ubb + t_ub
        rm_lb = lam_lb * t_lb
        rm_ub = lam_ub * t_ub
        mu = (rm_lb.sum() + rm_ub.sum()) / nc
        return rg, rb, rd_lb, rd_ub, rm_lb, rm_ub, mu

    def norms(rg, rb, rd_lb, rd_ub, rm_lb, rm_ub):
        return (np.max(np.abs(rg)), np.max(np.abs(rb)) if ne else 0.0,
                max(np.max(np.abs(rd_lb)), np.max(np.abs(rd_ub))),
                max(np.max(np.abs(rm_lb)), np.max(np.abs(rm_ub))))

    K = np.zeros((nv + ne, nv + ne))
    K[:nv, nv:] = G.T
    K[nv:, :nv] = G

    rg, rb, rd_lb, rd_ub, rm_lb, rm_ub, mu = residuals()
    nrm = norms(rg, rb, rd_lb, rd_ub, rm_lb, rm_ub)
    alpha = 1.0
    it = 0
    tols = tuple(spec.tol if v is None else v for v in (spec.qp_tol_stat, spec.qp_tol_eq, spec.qp_tol_ineq, spec.qp_tol_comp))
    unconverged = lambda n: n[0] > tols[0] or n[1] > tols[1] or n[2] > tols[2] or n[3] > tols[3]
    while it < spec.qp_max_iter and alpha > spec.alpha_min and unconverged(nrm):
        tinv_lb, tinv_ub = 1.0 / t_lb, 1.0 / t_ub
        Gam = np.zeros(nv)
        Gam[ib] = lam_lb * tinv_lb + lam_ub * tinv_ub
        K[np.arange(nv), np.arange(nv)] = Hd + Gam

        def solve(rm_lb_, rm_ub_):
            gam_lb = tinv_lb * (rm_lb_ - lam_lb * rd_lb)
            gam_ub = tinv_ub * (rm_ub_ - lam_ub * rd_ub)
            gt = rg.copy()
            gt[ib] += gam_lb - gam_ub
            sol = np.linalg.solve(K, np.concatenate([-gt, -rb]))
            dz_, dpi_ = sol[:nv], sol[nv:]
            dt_lb_ = dz_[ib] - rd_lb
            dt_ub_ = -dz_[ib] - rd_ub
            dlam_lb_ = -tinv_lb * (lam_lb * dt_lb_ + rm_lb_)
            dlam_ub_ = -tinv_ub * (lam_ub * dt_ub_ + rm_ub_)
            return dz_, dpi_, dt_lb_, dt_ub_, dlam_lb_, dlam_ub_

        def step_len(dt_lb_, dt_ub_, dlam_lb_, dlam_ub_):
            # COMPUTE_ALPHA_QP: largest alpha <= 1 keeping lam, t >= 0
            a = 1.0
            for v_, dv_ in ((lam_lb, dlam_lb_), (lam_ub, dlam_ub_), (t_lb, dt_lb_), (t_ub, dt_ub_)):
                neg = dv_ < 0
                if np.any(neg):
                    a = min(a, np.min(-v_[neg] / dv_[neg]))
            return a

        def mu_aff_of(a, dt_lb_, dt_ub_, dlam_lb_, dlam_ub_):
            a = a * spec.mu_aff_shrink  # COMPUTE_MU_AFF_QP: "this affects the minimum value of sigma"
            return (np.sum((lam_lb + a * dlam_lb_) * (t_lb + a * dt_lb_)) +
                    np.sum((lam_ub + a * dlam_ub_) * (t_ub + a * dt_ub_))) / nc

        # predictor (affine) step
        step = solve(rm_lb, rm_ub)
        alpha = step_len(*step[2:])
        mu_aff = mu_aff_of(alpha, *step[2:])
        sigma = (mu_aff / mu) ** 3
        sigma_mu = max(sigma * mu, spec.t_min)
        # corrector: res_m <- res_m_bkp + dt_aff*dlam_aff - sigma*mu
        cm_lb = rm_lb + step[2] * step[4] - sigma_mu
        cm_ub = rm_ub + step[3] * step[5] - sigma_mu
        step = solve(cm_lb, cm_ub)
        alpha = step_len(*step[2:])
        # conditional predictor-corrector: fall back to a pure centering step when the corrector is poor
        mu_aff_c = mu_aff_of(alpha, *step[2:])
        if mu_aff_c > 2.0 * mu_aff:
            step = solve(rm_lb - sigma_mu, rm_ub - sigma_mu)
            alpha = step_len(*step[2:])
        # UPDATE_VAR_QP
        a = alpha * ((1.0 - alpha) * 0.99 + alpha * 0.9999999)
        dz, dpi, dt_lb, dt_ub, dlam_lb, dlam_ub = step
        z = z + a * dz
        pi = pi + a * dpi
        lam_lb = np.maximum(lam_lb + a * dlam_lb, spec.lam_min)
        lam_ub = np.maximum(lam_ub + a * dlam_ub, spec.lam_min)
        t_lb = np.maximum(t_lb + a * dt_lb, spec.t_min)
        t_ub = np.maximum(t_ub + a * dt_ub, spec.t_min)
        rg, rb, rd_lb, rd_ub, rm_lb, rm_ub, mu = residuals()
        nrm = norms(rg, rb, rd_lb, rd_ub, rm_lb, rm_ub)
        it += 1

    if not np.all(np.isfinite(z)):
        status = 3
    elif it >= spec.qp_max_iter and unconverged(nrm):
        status = 1
    elif alpha <= spec.alpha_min:
        status = 2
    else:
        status = 0
    full = lambda v: _scatter(v, ib, nv)
    return QpResult(z, pi, full(lam_lb), full(lam_ub), full(t_lb), full(t_ub), it, status, nrm)


def _scatter(v, ib, n):
    out = np.zeros(n)
    out[ib] = v
    return out


# ----------------------------------------------------------------------------------------------------------------------
# acados-style SQP solver object (AcadosOcpSolver surface used by the reference)
# ----------------------------------------------------------------------------------------------------------------------
ACADOS_SUCCESS, ACADOS_FAILURE, ACADOS_MAXITER, ACADOS_MINSTEP, ACADOS_QP_FAILURE = 0, 1, 2, 3, 4   # src/Readme.md:14-20


class OracleOcpSolver:
    """set(stage, field, value) / solve() / get(stage, field), the calls the reference makes
    (src/force_model/controller.py:30-39, src/force_model/ocp.py:120-122)."""

    def __init__(self, spec: OcpSpec, p=(MASS, GRAVITY_ACC), rti=False):
        self.spec = spec
        s = spec
        self.p = np.array(p, float)
        self.rti = rti
        self.x = np.zeros((s.N + 1, s.nx))
        self.u = np.zeros((s.N, s.nu))
        self.pi = np.zeros((s.N, s.nx))
        self.yref = np.zeros((s.N, s.nx + s.nu))
        self.yref_e = np.zeros(s.nx)
        self.x0 = np.zeros(s.nx)
        # per-stage boxes (acados accepts 'lbu' / 'ubu' at every stage and 'lbx' / 'ubx' at stages 1..N-1 through set();
        # the reference leaves them at the values of create_ocp, src/force_model/ocp.py:62-76)
        self.lbu_k = np.tile(s.lbu, (s.N, 1)); self.ubu_k = np.tile(s.ubu, (s.N, 1))
        self.lbx_k = np.tile(s.lbx, (s.N, 1)); self.ubx_k = np.tile(s.ubx, (s.N, 1))
        self.lam = None
        self.status = 0
        self.sqp_iter = 0
        self.qp_iter = 0
        self.qp_iters = []
        self.res = (0, 0, 0, 0)
        # variable layout of the x0-eliminated QP (HPIPM ordering, u before x): [u_0 | u_1 x_1 | ... | x_N]
        off, self.off_u, self.off_x = 0, [], [None] * (s.N + 1)
        for k in range(s.N):
            self.off_u.append(off); off += s.nu
            if k >= 1:
                self.off_x[k] = off; off += s.nx
        self.off_x[s.N] = off; off += s.nx
        self.nv = off

    # -- reference-facing surface ---------------------------------------------------------------------------------
    def set(self, stage, fld, val):
        val = np.asarray(val, float)
        s = self.spec
        if fld == 'yref':
            if stage == s.N:
                self.yref_e[:] = val
            else:
                self.yref[stage] = val
        elif fld in ('lbx', 'ubx') and stage == 0:
            # the reference only sets stage 0 (x0 embedding, controller.py:30-31)
            self.x0[:] = val
        elif fld in ('lbx', 'ubx', 'lbu', 'ubu'):
            assert (1 if fld.endswith('x') else 0) <= stage < s.N
            getattr(self, fld + '_k')[stage] = val
        elif fld == 'x':
            self.x[stage] = val
        elif fld == 'u':
            self.u[stage] = val
        elif fld == 'p':
            self.p[:] = val
        else:
            raise ValueError(fld)

    def get(self, stage, fld):
        if fld == 'x':
            return self.x[stage].copy()
        if fld == 'u':
            return self.u[stage].copy()
        if fld == 'pi':
            return self.pi[stage].copy()
        raise ValueError(fld)

    # -- one SQP run ------------------------------------------------------------------------------------------------
    def _linearise(self):
        s = self.spec
        A = np.zeros((s.N, s.nx, s.nx)); B = np.zeros((s.N, s.nx, s.nu)); b = np.zeros((s.N, s.nx))
        for k in range(s.N):
            xn, S = erk_step(s.f, s.jac, self.x[k], self.u[k], self.p, s.dt, s.erk_stages)
            A[k], B[k] = S[:, :s.nx], S[:, s.nx:]
            b[k] = xn - self.x[k + 1]
        return A, B, b

    def _build_qp(self, A, B, b):
        """x0-eliminated delta QP around the current iterate."""
        s = self.spec
        nv, ne = self.nv, s.N * s.nx
        Hd = np.zeros(nv); q = np.zeros(nv); lb = np.full(nv, -np.inf); ub = np.full(nv, np.inf)
        G = np.zeros((ne, nv)); bv = np.zeros(ne)
        dx0 = self.x0 - self.x[0]
        for k in range(s.N):
            ou = self.off_u[k]
            Hd[ou:ou + s.nu] = s.dt * s.w[s.nx:]
            q[ou:ou + s.nu] = s.dt * s.w[s.nx:] * (self.u[k] - self.yref[k, s.nx:])
            lb[ou:ou + s.nu] = self.lbu_k[k] - self.u[k]
            ub[ou:ou + s.nu] = self.ubu_k[k] - self.u[k]
            if k >= 1:
                ox = self.off_x[k]
                Hd[ox:ox + s.nx] = s.dt * s.w[:s.nx]
                q[ox:ox + s.nx] = s.dt * s.w[:s.nx] * (self.x[k] - self.yref[k, :s.nx])
                lb[ox:ox + s.nx] = self.lbx_k[k] - self.x[k]
                ub[ox:ox + s.nx] = self.ubx_k[k] - self.x[k]
            rows = slice(k * s.nx, (k + 1) * s.nx)
            G[rows, ou:ou + s.nu] = B[k]
            if k >= 1:
                G[rows, self.off_x[k]:self.off_x[k] + s.nx] = A[k]
                bv[rows] = b[k]
            else:
                bv[rows] = b[k] + A[k] @ dx0          # x0 eliminated (the stage-0 cost gradient term is constant)
            on = self.off_x[k + 1]
            G[rows, on:on + s.nx] = -np.eye(s.nx)
        ox = self.off_x[s.N]
        Hd[ox:ox + s.nx] = s.w_e
        q[ox:ox + s.nx] = s.w_e * (self.x[s.N] - self.yref_e)
        return Hd, q, G, bv, lb, ub, dx0

    def _nlp_residuals(self, A, B, b):
        """acados ocp_nlp_res_compute at the current iterate with the current multipliers."""
        s = self.spec
        res_eq = np.max(np.abs(b)) if s.N else 0.0
        res_eq = max(res_eq, np.max(np.abs(self.x0 - self.x[0])))
        if self.lam is None:
            return np.inf, res_eq, np.inf, np.inf
        lam_lb, lam_ub = self.lam
        stat, ineq, comp = 0.0, 0.0, 0.0
        for k in range(s.N + 1):
            if k < s.N:
                ou = self.off_u[k]
                gu = s.dt * s.w[s.nx:] * (self.u[k] - self.yref[k, s.nx:]) + B[k].T @ self.pi[k] \
                    - lam_lb[ou:ou + s.nu] + lam_ub[ou:ou + s.nu]
                stat = max(stat, np.max(np.abs(gu)))
                for lo, hi, v, ll, lu in ((self.lbu_k[k], self.ubu_k[k], self.u[k], lam_lb[ou:ou + s.nu], lam_ub[ou:ou + s.nu]),):
                    ineq = max(ineq, np.max(np.maximum(lo - v, 0)), np.max(np.maximum(v - hi, 0)))
                    comp = max(comp, np.max(np.abs(ll * (lo - v))), np.max(np.abs(lu * (v - hi))))
            if k >= 1:
                ox = self.off_x[k]
                if k < s.N:
                    gx = s.dt * s.w[:s.nx] * (self.x[k] - self.yref[k, :s.nx]) + A[k].T @ self.pi[k] - self.pi[k - 1] \
                        - lam_lb[ox:ox + s.nx] + lam_ub[ox:ox + s.nx]
                    v = self.x[k]
                    ineq = max(ineq, np.max(np.maximum(self.lbx_k[k] - v, 0)), np.max(np.maximum(v - self.ubx_k[k], 0)))
                    comp = max(comp, np.max(np.abs(lam_lb[ox:ox + s.nx] * (self.lbx_k[k] - v))),
                               np.max(np.abs(lam_ub[ox:ox + s.nx] * (v - self.ubx_k[k]))))
                else:
                    gx = s.w_e * (self.x[k] - self.yref_e) - self.pi[k - 1]
                stat = max(stat, np.max(np.abs(gx)))
        return stat, res_eq, ineq, comp

    def solve(self):
        s = self.spec
        self.qp_iters = []
        self.sqp_iter = 0
        if not (np.all(np.isfinite(self.x0)) and np.all(np.isfinite(self.yref)) and np.all(np.isfinite(self.yref_e))):
            self.status = ACADOS_FAILURE
            return self.status
        max_it = 1 if self.rti else s.sqp_max_iter
        for it in range(max_it + 1):
            A, B, b = self._linearise()
            if not self.rti:
                res = self._nlp_residuals(A, B, b)
                self.res = res
                if all(r < s.tol for r in res):
                    self.status = ACADOS_SUCCESS
                    break
                if it >= max_it:
                    self.status = ACADOS_MAXITER
                    break
            elif it >= 1:
                break
            Hd, q, G, bv, lb, ub, dx0 = self._build_qp(A, B, b)
            r = qp_ipm_dense(Hd, q, G, bv, lb, ub, s)
            self.qp_iters.append(r.iters)
            self.last_qp = r
            if r.status not in (0, 1):
                self.status = ACADOS_QP_FAILURE
                self.sqp_iter = it + 1
                break
            # full step
            self.x[0] = self.x[0] + dx0
            for k in range(s.N):
                ou = self.off_u[k]
                self.u[k] += r.dz[ou:ou + s.nu]
                if k >= 1:
                    self.x[k] += r.dz[self.off_x[k]:self.off_x[k] + s.nx]
            self.x[s.N] += r.dz[self.off_x[s.N]:self.off_x[s.N] + s.nx]
            self.pi = r.pi.reshape(s.N, s.nx).copy()
            self.lam = (r.lam_lb, r.lam_ub)
            self.sqp_iter = it + 1
            if self.rti:
                self.status = ACADOS_SUCCESS if r.status == 0 else ACADOS_MAXITER
        self.qp_iter = int(sum(self.qp_iters))
        return self.status


# ----------------------------------------------------------------------------------------------------------------------
# Converters, plant step, trajectory, metrics, closed loops
# ----------------------------------------------------------------------------------------------------------------------
def convert_force(u):
    """reference src/force_model/dynamics.py:66-70"""
    return np.array([np.arctan2(u[0], u[1]), np.sqrt(u[0] * u[0] + u[1] * u[1])])


def convert_jerk(h, a_i, mass=MASS):
    """reference src/jerk_model/dynamics.py:76-83 (a_i is advanced in place over the 10 sub-steps)"""
    a = np.array(a_i, float)
    out = np.zeros((CTRLS_PER_SAMPLE, 2))
    for j in range(CTRLS_PER_SAMPLE):
        a = a + h * DT_CONV
        fx, fz = mass * a[0], mass * a[1]
        out[j] = (np.arctan2(fx, fz), np.sqrt(fx * fx + fz * fz))
    return out, a


def plant_step_force(x, u_plant, p, eps):
    """reference src/force_model/ocp.py:98-115: ERK4, one step of dt, then one scalar eps added to all states"""
    return erk_step(f_plant, jac_plant, x, u_plant, p, DT, 4, 1, sens=False) + eps


def plant_step_jerk(x, u_plant10, p, eps):
    """reference src/jerk_model/ocp.py:97-116: Euler, dt_conv, 10 sub-steps each with its own input"""
    xi = np.array(x, float)
    for j in range(CTRLS_PER_SAMPLE):
        xi = erk_step(f_plant, jac_plant, xi, u_plant10[j], p, DT_CONV, 1, 1, sens=False)
    return xi + eps


def gen_circle_traj(n_steps=N_STEPS, n_horizon=N_HORIZON, center=(0.0, 0.0), radius=1.0, phase=0.0):
    """reference src/generate_trajectory.py:7-28 (nx=6, nu=2 layout: 8 columns); phase is our extension"""
    ref = np.zeros((n_steps + n_horizon, 8))
    omega = 2 * np.pi / T_END
    t = np.linspace(0, T_END, n_steps)
    a = omega * t + phase
    ref[:n_steps, 0] = center[0] + radius * np.cos(a)
    ref[:n_steps, 1] = center[1] + radius * np.sin(a)
    ref[:n_steps, 2] = -radius * omega * np.sin(a)
    ref[:n_steps, 3] = radius * omega * np.cos(a)
    ref[:n_steps, 4] = -radius * omega ** 2 * np.cos(a)
    ref[:n_steps, 5] = -radius * omega ** 2 * np.sin(a) + GRAVITY_ACC
    ref[n_steps:] = ref[:n_horizon]
    return ref


def calc_aed(pref, psim):
    """reference src/store_results.py:233-236 (mean absolute coordinate error)"""
    return float(np.mean(np.sqrt((pref - psim) ** 2)))


_WCOST = np.array([1e2, 1e2, 1.0, 1.0])


def follow_trajectory_force(xref, uref, x0, eps, n_steps=N_STEPS, spec=None, p_ctrl=(MASS, GRAVITY_ACC),
                            p_plant=(MASS, GRAVITY_ACC), rti=False, raise_on_fail=True):
    """reference src/force_model/controller.py:8-56; eps[i] is the noise draw of step i (0 for noise=False)"""
    spec = spec or force_ocp()
    sol = OracleOcpSolver(spec, p_ctrl, rti=rti)
    N = spec.N
    Xsim = np.zeros((n_steps + 1, 4)); U_plant = np.zeros((n_steps, 2)); a = np.zeros((n_steps, 2))
    U_ctrl = np.zeros((n_steps, 2)); stat = np.zeros(n_steps, int); qpit = np.zeros(n_steps, int)
    Xsim[0] = x0
    cost_sum = 0.0
    for i in range(n_steps):
        for k in range(N):
            sol.set(k, 'yref', np.hstack((xref[i + k], uref[i + k])))
        sol.set(N, 'yref', xref[i + N])
        sol.set(0, 'lbx', Xsim[i]); sol.set(0, 'ubx', Xsim[i])
        status = sol.solve()
        stat[i], qpit[i] = status, sol.qp_iter
        if status != 0 and raise_on_fail:
            raise RuntimeError(f'Failed in iteration {i}: status {status}')
        u = sol.get(0, 'u')
        U_ctrl[i] = u
        a[i] = u / MASS
        xo = sol.get(0, 'x')
        d = xo[:4] - xref[i, :4]
        cost_sum += float(d @ (_WCOST * d))
        U_plant[i] = convert_force(u)
        Xsim[i + 1] = plant_step_force(Xsim[i], U_plant[i], p_plant, eps[i])
    return dict(cost=cost_sum, Xsim=Xsim, a=a, U_plant=U_plant, U_ctrl=U_ctrl, status=stat, qp_iter=qpit)


def follow_trajectory_jerk(xref, uref, x0, eps, n_steps=N_STEPS, spec=None, p_ctrl=(MASS, GRAVITY_ACC),
                           p_plant=(MASS, GRAVITY_ACC), rti=False, raise_on_fail=True):
    """reference src/jerk_model/controller.py:8-58"""
    spec = spec or jerk_ocp()
    sol = OracleOcpSolver(spec, p_ctrl, rti=rti)
    N = spec.N
    Xsim = np.zeros((n_steps + 1, 4)); U_plant = np.zeros((n_steps, 2)); a = np.zeros((n_steps, 2))
    U_ctrl = np.zeros((n_steps, 2)); stat = np.zeros(n_steps, int); qpit = np.zeros(n_steps, int)
    a_i = np.array([0.0, GRAVITY_ACC])
    Xsim[0] = x0
    cost_sum = 0.0
    for i in range(n_steps):
        for k in range(N):
            sol.set(k, 'yref', np.hstack((xref[i + k], uref[i + k])))
        sol.set(N, 'yref', xref[i + N])
        x0_bar = np.hstack((Xsim[i], a_i))
        sol.set(0, 'lbx', x0_bar); sol.set(0, 'ubx', x0_bar)
        status = sol.solve()
        stat[i], qpit[i] = status, sol.qp_iter
        if status != 0 and raise_on_fail:
            raise RuntimeError(f'Failed in iteration {i}: status {status}')
        h = sol.get(0, 'u')
        U_ctrl[i] = h
        xo = sol.get(1, 'x')
        d = xo[:4] - xref[i, :4]
        cost_sum += float(d @ (_WCOST * d))
        u_tmp, a_i = convert_jerk(h, a_i, p_ctrl[0])
        a[i] = a_i
        U_plant[i] = u_tmp[-1]
        Xsim[i + 1] = plant_step_jerk(Xsim[i], u_tmp, p_plant, eps[i])
    return dict(cost=cost_sum, Xsim=Xsim, a=a, U_plant=U_plant, U_ctrl=U_ctrl, status=stat, qp_iter=qpit)


def main_py_noise(seed=42, n=2 * N_STEPS):
    """The noise stream of reference src/main.py:43-46: np.random.seed(42), then one normal(0, 0.01) per control
    step; the force run consumes draws 0..499, the jerk run draws 500..999."""
    rs = np.random.RandomState(seed)
    return rs.normal(0, NOISE_STD, size=n)
