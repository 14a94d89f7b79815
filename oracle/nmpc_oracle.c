/* CPU oracle (plain C, FP64) for the NMPC hot path of BroilerCompiler/drone-attitude-control.
 *
 * TEST INFRASTRUCTURE ONLY.  The product library (libbnmpc.so) never links, loads or calls this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, as the checker and as
 * the reported CPU baseline ("port": the reference's own arithmetic lives in acados/HPIPM/BLASFEO/CasADi, which
 * are neither in /root/reference nor installable here - see oracle/nmpc_oracle.py for the pins).
 *
 * Same algorithm as oracle/nmpc_oracle.py (which is pinned to the decoded acados run, tests/test_oracle_golden.py),
 * but the Newton systems are solved with a stage-wise Riccati recursion like HPIPM does, instead of a dense KKT
 * solve; tests/test_oracle_c.py checks the two against each other and this file against the golden series.
 *
 * What it restates, per control step (reference file:line):
 *   OCP.set_up_ocp            src/force_model/ocp.py:117-122, src/jerk_model/ocp.py:118-123   (yref window)
 *   x0 embedding              src/force_model/controller.py:29-31, src/jerk_model/controller.py:30-32
 *   AcadosOcpSolver.solve()   OCP of src/force_model/ocp.py:21-96 / src/jerk_model/ocp.py:20-95, models
 *                             src/force_model/dynamics.py:32-37, src/jerk_model/dynamics.py:35-42
 *   Converter.convert         src/force_model/dynamics.py:66-70, src/jerk_model/dynamics.py:76-83
 *   OCP.simulate_next_x       src/force_model/ocp.py:98-115, src/jerk_model/ocp.py:97-116; plant src/plant.py:27-33
 *   logged cost / a           src/force_model/controller.py:38-41, src/jerk_model/controller.py:39-46
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NXM 12
#define NUM 4
#define NSM (NXM + NUM)

/* MODEL_PLANT doubles as the controller model of the "thrust" OCP (inputs theta, Fd: a nonlinear OCP that is not in the
 * reference - it exercises the general SQP path: state/input dependent sensitivities, several SQP iterations) */
enum { MODEL_FORCE = 0, MODEL_JERK = 1, MODEL_PLANT = 2, MODEL_ATT = 3 };
/* MODEL_ATT (NOT in the reference; the north-star's 3-D attitude-and-total-thrust model): x = (p, v, q) with
 * the attitude quaternion q = (w, x, y, z) body->world, u = (T, wx, wy, wz):  pdot = v, vdot = (T/m) R(q) e3 - g e3,
 * qdot = 1/2 q (x) (0, w).  The planar plant of src/plant.py:27-33 is its restriction to the x-z plane.  PARITY UNPINNED for
 * this model: the reference holds no run of it; the two oracles check each other (tests/test_att.py). */

typedef struct {
    int model;         /* MODEL_FORCE | MODEL_JERK */
    int N;             /* horizon */
    int erk_stages;    /* OCP integrator: 4 (force; stands in for IRK, exact for the affine model), 1 (jerk); 0 = IRK (irk_gl4_step) */
    int sqp_max_iter;  /* acados nlp_solver_max_iter (100) */
    int qp_max_iter;   /* acados qp_solver_iter_max (50) */
    int rti;           /* 1: exactly one QP per call, no NLP residual test (SQP_RTI) */
    double dt;
    double w[NSM];     /* diag W, order [x; u] (LINEAR_LS with Vx=[I;0], Vu=[0;I]) */
    double w_e[NXM];   /* diag W_e */
    double lbx[NXM], ubx[NXM], lbu[NUM], ubu[NUM];
    double tol[4];     /* NLP tolerances stat, eq, ineq, comp (1e-6) */
    double qp_tol[4];  /* QP tolerances res_g, res_b, res_d, res_m (acados hands the NLP ones to HPIPM) */
    double mu0, thr0, alpha_min, lam_min, t_min;
} orc_opts;

/* Timing builds of bench.py fix the dimensions at compile time (-DORC_FIXED_NX=4 -DORC_FIXED_NU=2 ...) so that the
 * compiler unrolls and vectorises the stage kernels; the default build reads them from the model at run time.  Same
 * arithmetic either way. */
#ifdef ORC_FIXED_NX
#define ORC_NX(v) ORC_FIXED_NX
#define ORC_NU(v) ORC_FIXED_NU
#else
#define ORC_NX(v) (v)
#define ORC_NU(v) (v)
#endif

static void model_dims(int model, int *nx, int *nu) {
    if (model == MODEL_ATT) { *nx = ORC_NX(10); *nu = ORC_NU(4); }
    else if (model == MODEL_JERK) { *nx = ORC_NX(6); *nu = ORC_NU(2); } else { *nx = ORC_NX(4); *nu = ORC_NU(2); }
}

/* xdot = f(x,u,p), p = (mass, g) */
static void model_f(int model, const double *x, const double *u, const double *p, double *xd) {
    const double m = p[0], g = p[1];
    switch (model) {
    case MODEL_FORCE: xd[0] = x[2]; xd[1] = x[3]; xd[2] = u[0] / m; xd[3] = u[1] / m - g; break;
    case MODEL_JERK:  xd[0] = x[2]; xd[1] = x[3]; xd[2] = x[4]; xd[3] = x[5] - g; xd[4] = u[0]; xd[5] = u[1]; break;
    case MODEL_ATT: {
        const double qw = x[6], qx = x[7], qy = x[8], qz = x[9], a = u[0] / m, wx = u[1], wy = u[2], wz = u[3];
        xd[0] = x[3]; xd[1] = x[4]; xd[2] = x[5];
        xd[3] = 2.0 * (qx * qz + qw * qy) * a; xd[4] = 2.0 * (qy * qz - qw * qx) * a; xd[5] = (1.0 - 2.0 * (qx * qx + qy * qy)) * a - g;
        xd[6] = 0.5 * (-qx * wx - qy * wy - qz * wz); xd[7] = 0.5 * (qw * wx + qy * wz - qz * wy);
        xd[8] = 0.5 * (qw * wy - qx * wz + qz * wx);  xd[9] = 0.5 * (qw * wz + qx * wy - qy * wx);
    } break;
    default:          xd[0] = x[2]; xd[1] = x[3]; xd[2] = u[1] * sin(u[0]) / m; xd[3] = u[1] * cos(u[0]) / m - g; break;
    }
}

/* fx (nx*nx, row-major), fu (nx*nu) */
static void model_jac(int model, const double *x, const double *u, const double *p, double *fx, double *fu) {
    const double m = p[0];
    int nx, nu; model_dims(model, &nx, &nu);
    (void)x;
    memset(fx, 0, sizeof(double) * nx * nx); memset(fu, 0, sizeof(double) * nx * nu);
    switch (model) {
    case MODEL_FORCE: fx[0 * 4 + 2] = 1; fx[1 * 4 + 3] = 1; fu[2 * 2 + 0] = 1 / m; fu[3 * 2 + 1] = 1 / m; break;
    case MODEL_JERK:  fx[0 * 6 + 2] = 1; fx[1 * 6 + 3] = 1; fx[2 * 6 + 4] = 1; fx[3 * 6 + 5] = 1; fu[4 * 2 + 0] = 1; fu[5 * 2 + 1] = 1; break;
    case MODEL_ATT: {
        const double qw = x[6], qx = x[7], qy = x[8], qz = x[9], a = u[0] / m, wx = u[1], wy = u[2], wz = u[3];
#define FX(r, c) fx[(r) * 10 + (c)]
#define FU(r, c) fu[(r) * 4 + (c)]
        FX(0, 3) = 1; FX(1, 4) = 1; FX(2, 5) = 1;
        FX(3, 6) = 2 * qy * a;  FX(3, 7) = 2 * qz * a;  FX(3, 8) = 2 * qw * a; FX(3, 9) = 2 * qx * a;
        FX(4, 6) = -2 * qx * a; FX(4, 7) = -2 * qw * a; FX(4, 8) = 2 * qz * a; FX(4, 9) = 2 * qy * a;
        FX(5, 7) = -4 * qx * a; FX(5, 8) = -4 * qy * a;
        FX(6, 7) = -0.5 * wx; FX(6, 8) = -0.5 * wy; FX(6, 9) = -0.5 * wz;
        FX(7, 6) = 0.5 * wx;  FX(7, 8) = 0.5 * wz;  FX(7, 9) = -0.5 * wy;
        FX(8, 6) = 0.5 * wy;  FX(8, 7) = -0.5 * wz; FX(8, 9) = 0.5 * wx;
        FX(9, 6) = 0.5 * wz;  FX(9, 7) = 0.5 * wy;  FX(9, 8) = -0.5 * wx;
        FU(3, 0) = 2.0 * (qx * qz + qw * qy) / m; FU(4, 0) = 2.0 * (qy * qz - qw * qx) / m; FU(5, 0) = (1.0 - 2.0 * (qx * qx + qy * qy)) / m;
        FU(6, 1) = -0.5 * qx; FU(6, 2) = -0.5 * qy; FU(6, 3) = -0.5 * qz;
        FU(7, 1) = 0.5 * qw;  FU(7, 2) = -0.5 * qz; FU(7, 3) = 0.5 * qy;
        FU(8, 1) = 0.5 * qz;  FU(8, 2) = 0.5 * qw;  FU(8, 3) = -0.5 * qx;
        FU(9, 1) = -0.5 * qy; FU(9, 2) = 0.5 * qx;  FU(9, 3) = 0.5 * qw;
#undef FX
#undef FU
    } break;
    default:
        fx[0 * 4 + 2] = 1; fx[1 * 4 + 3] = 1;
        fu[2 * 2 + 0] = u[1] * cos(u[0]) / m;  fu[2 * 2 + 1] = sin(u[0]) / m;
        fu[3 * 2 + 0] = -u[1] * sin(u[0]) / m; fu[3 * 2 + 1] = cos(u[0]) / m; break;
    }
}

static const double ERK_A[5][4][4] = {
    {{0}}, {{0}},
    {{0, 0, 0, 0}, {0.5, 0, 0, 0}},
    {{0, 0, 0, 0}, {0.5, 0, 0, 0}, {-1.0, 2.0, 0, 0}},
    {{0, 0, 0, 0}, {0.5, 0, 0, 0}, {0, 0.5, 0, 0}, {0, 0, 1.0, 0}}};
static const double ERK_B[5][4] = {{0}, {1.0}, {0.0, 1.0}, {1.0 / 6, 2.0 / 3, 1.0 / 6}, {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6}};

/* acados sim_erk: num_steps explicit RK steps over [0,T]; S (nx x (nx+nu), row-major) = forward sensitivities, may be NULL */
static void erk_step(int model, const double *x0, const double *u, const double *p, double T, int ns, int num_steps,
                     double *xn, double *S) {
    int nx, nu; model_dims(model, &nx, &nu);
    const int nc = nx + nu;
    const double h = T / num_steps;
    double x[NXM], K[4][NXM], SK[4][NXM * NSM], xi[NXM], Si[NXM * NSM], fx[NXM * NXM], fu[NXM * NUM];
    memcpy(x, x0, sizeof(double) * nx);
    if (S) { memset(S, 0, sizeof(double) * nx * nc); for (int i = 0; i < nx; i++) S[i * nc + i] = 1.0; }
    for (int st = 0; st < num_steps; st++) {
        for (int i = 0; i < ns; i++) {
            for (int r = 0; r < nx; r++) { double a = 0; for (int j = 0; j < i; j++) a += ERK_A[ns][i][j] * K[j][r]; xi[r] = x[r] + h * a; }
            model_f(model, xi, u, p, K[i]);
            if (S) {
                for (int e = 0; e < nx * nc; e++) { double a = 0; for (int j = 0; j < i; j++) a += ERK_A[ns][i][j] * SK[j][e]; Si[e] = S[e] + h * a; }
                model_jac(model, xi, u, p, fx, fu);
                for (int r = 0; r < nx; r++) for (int c = 0; c < nc; c++) {
                    double a = 0; for (int l = 0; l < nx; l++) a += fx[r * nx + l] * Si[l * nc + c];
                    if (c >= nx) a += fu[r * nu + (c - nx)];
                    SK[i][r * nc + c] = a;
                }
            }
        }
        for (int r = 0; r < nx; r++) { double a = 0; for (int i = 0; i < ns; i++) a += ERK_B[ns][i] * K[i][r]; x[r] += h * a; }
        if (S) for (int e = 0; e < nx * nc; e++) { double a = 0; for (int i = 0; i < ns; i++) a += ERK_B[ns][i] * SK[i][e]; S[e] += h * a; }
    }
    memcpy(xn, x, sizeof(double) * nx);
}


/* acados sim_irk restated (integrator_type 'IRK', reference src/force_model/ocp.py:85; same algorithm as
 * oracle/nmpc_oracle.py irk_gl4_step): Gauss-Legendre collocation with 4 stages, Newton from K = 0 with the exact Jacobian
 * I - h (A (x) f_x) re-evaluated in each of 3 iterations (dense LU, partial pivoting), sensitivities from the implicit
 * function theorem at the final iterate.  Parity unpinned beyond the affine case (for which it is the exact discretisation). */
static const double GL4_A[4][4] = {
    {0.086963711284363464343, -0.026604180084998793313, 0.012627462689404724515, -0.0035551496857956831569},
    {0.18811811749986807165, 0.16303628871563653566, -0.027880428602470895224, 0.0067355005945381555154},
    {0.16719192197418877317, 0.35395300603374396654, 0.16303628871563653566, -0.014190694931141142964},
    {0.17748257225452261184, 0.3134451147418683468, 0.35267675751627186463, 0.086963711284363464343}};
static const double GL4_B[4] = {0.17392742256872692869, 0.32607257743127307131, 0.32607257743127307131, 0.17392742256872692869};

static void lu_solve(int D, double *G, double *R, int ldr, int nrhs) {
    for (int c = 0; c < D; c++) {
        int pv = c; double best = fabs(G[c * D + c]);
        for (int r = c + 1; r < D; r++) if (fabs(G[r * D + c]) > best) { best = fabs(G[r * D + c]); pv = r; }
        if (pv != c) {
            for (int j = 0; j < D; j++) { double t = G[c * D + j]; G[c * D + j] = G[pv * D + j]; G[pv * D + j] = t; }
            for (int j = 0; j < nrhs; j++) { double t = R[c * ldr + j]; R[c * ldr + j] = R[pv * ldr + j]; R[pv * ldr + j] = t; }
        }
        const double inv = 1.0 / G[c * D + c];
        for (int r = c + 1; r < D; r++) {
            const double l = G[r * D + c] * inv;
            if (l == 0.0) continue;
            for (int j = c + 1; j < D; j++) G[r * D + j] -= l * G[c * D + j];
            for (int j = 0; j < nrhs; j++) R[r * ldr + j] -= l * R[c * ldr + j];
        }
    }
    for (int c = D - 1; c >= 0; c--) {
        const double inv = 1.0 / G[c * D + c];
        for (int j = 0; j < nrhs; j++) {
            double a = R[c * ldr + j];
            for (int l = c + 1; l < D; l++) a -= G[c * D + l] * R[l * ldr + j];
            R[c * ldr + j] = a * inv;
        }
    }
}

static void irk_gl4_step(int model, const double *x0, const double *u, const double *p, double h, double *xn, double *S) {
    int nx, nu; model_dims(model, &nx, &nu);
    const int nc = nx + nu, D = 4 * nx, ldr = S ? nc : 1;
    double K[4 * NXM], G[16 * NXM * NXM], R[4 * NXM * NSM], xi[NXM], fi[NXM], fx[NXM * NXM], fu[NXM * NUM];
    memset(K, 0, sizeof(K));
    for (int it = 0; it < 3 + (S ? 1 : 0); it++) {
        const int last = it == 3;
        for (int i = 0; i < 4; i++) {
            for (int r = 0; r < nx; r++) { double a = 0; for (int j = 0; j < 4; j++) a += GL4_A[i][j] * K[j * nx + r]; xi[r] = x0[r] + h * a; }
            model_f(model, xi, u, p, fi);
            model_jac(model, xi, u, p, fx, fu);
            for (int j = 0; j < 4; j++) {
                const double ha = h * GL4_A[i][j];
                for (int r = 0; r < nx; r++) for (int c = 0; c < nx; c++)
                    G[(i * nx + r) * D + j * nx + c] = ((i == j && r == c) ? 1.0 : 0.0) - ha * fx[r * nx + c];
            }
            for (int r = 0; r < nx; r++) {
                if (!last) R[(i * nx + r) * ldr] = fi[r] - K[i * nx + r];
                else {
                    for (int c = 0; c < nx; c++) R[(i * nx + r) * ldr + c] = fx[r * nx + c];
                    for (int c = 0; c < nu; c++) R[(i * nx + r) * ldr + nx + c] = fu[r * nu + c];
                }
            }
        }
        lu_solve(D, G, R, ldr, last ? nc : 1);
        if (!last) for (int e = 0; e < D; e++) K[e] += R[e * ldr];
    }
    for (int r = 0; r < nx; r++) { double a = 0; for (int i = 0; i < 4; i++) a += GL4_B[i] * K[i * nx + r]; xn[r] = x0[r] + h * a; }
    if (S) for (int r = 0; r < nx; r++) for (int c = 0; c < nc; c++) {
        double a = 0; for (int i = 0; i < 4; i++) a += GL4_B[i] * R[(i * nx + r) * ldr + c];
        S[r * nc + c] = ((r == c) ? 1.0 : 0.0) + h * a;
    }
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* one solver instance                                                                                              */
typedef struct {
    int nx, nu, N;
    /* iterate (kept between calls, like acados) */
    double *x, *u, *pi, *lam_lbu, *lam_ubu, *lam_lbx, *lam_ubx;
    /* linearisation */
    double *A, *B, *b;
    /* QP data + IPM state */
    double *qu, *qx, *zu, *zx, *ppi, *t_lbu, *t_ubu, *t_lbx, *t_ubx, *l_lbu, *l_ubu, *l_lbx, *l_ubx;
    double *dzu, *dzx, *dpi, *du_aff, *dx_aff;
    double *P, *pv, *Lr, *K;          /* Riccati factors */
    double *Phi, *wv;                 /* closed-loop matrices A + B K (nx x nx per stage), w_k = gu + B'P rb */
    double *gu, *gx;                  /* modified gradients */
    double *rgu, *rgx, *rb;           /* residuals */
    int have_mult;
} inst_t;

static double *dal(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

static inst_t *inst_new(int nx, int nu, int N) {
    inst_t *s = (inst_t *)calloc(1, sizeof(inst_t));
    s->nx = nx; s->nu = nu; s->N = N;
    size_t X = (size_t)(N + 1) * nx, U = (size_t)N * nu;
    s->x = dal(X); s->u = dal(U); s->pi = dal(X);
    s->lam_lbu = dal(U); s->lam_ubu = dal(U); s->lam_lbx = dal(X); s->lam_ubx = dal(X);
    s->A = dal((size_t)N * nx * nx); s->B = dal((size_t)N * nx * nu); s->b = dal(X);
    s->qu = dal(U); s->qx = dal(X); s->zu = dal(U); s->zx = dal(X); s->ppi = dal(X);
    s->t_lbu = dal(U); s->t_ubu = dal(U); s->t_lbx = dal(X); s->t_ubx = dal(X);
    s->l_lbu = dal(U); s->l_ubu = dal(U); s->l_lbx = dal(X); s->l_ubx = dal(X);
    s->dzu = dal(U); s->dzx = dal(X); s->dpi = dal(X); s->du_aff = dal(U); s->dx_aff = dal(X);
    s->P = dal((size_t)(N + 1) * nx * nx); s->pv = dal(X); s->Lr = dal((size_t)N * nu * nu); s->K = dal((size_t)N * nu * nx);
    s->Phi = dal((size_t)N * nx * nx); s->wv = dal(U);
    s->gu = dal(U); s->gx = dal(X); s->rgu = dal(U); s->rgx = dal(X); s->rb = dal(X);
    return s;
}

static void inst_free(inst_t *s) {
    double **f[] = {&s->x, &s->u, &s->pi, &s->lam_lbu, &s->lam_ubu, &s->lam_lbx, &s->lam_ubx, &s->A, &s->B, &s->b, &s->qu, &s->qx,
                    &s->zu, &s->zx, &s->ppi, &s->t_lbu, &s->t_ubu, &s->t_lbx, &s->t_ubx, &s->l_lbu, &s->l_ubu, &s->l_lbx, &s->l_ubx,
                    &s->dzu, &s->dzx, &s->dpi, &s->du_aff, &s->dx_aff, &s->P, &s->pv, &s->Lr, &s->K, &s->gu, &s->gx, &s->rgu, &s->rgx, &s->rb, &s->Phi, &s->wv};
    for (size_t i = 0; i < sizeof(f) / sizeof(f[0]); i++) free(*f[i]);
    free(s);
}

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }

/* relative bounds of the delta QP */
static inline double LBU(const orc_opts *o, const inst_t *s, int k, int j) { return o->lbu[j] - s->u[k * s->nu + j]; }
static inline double UBU(const orc_opts *o, const inst_t *s, int k, int j) { return o->ubu[j] - s->u[k * s->nu + j]; }
static inline double LBX(const orc_opts *o, const inst_t *s, int k, int j) { return o->lbx[j] - s->x[k * s->nx + j]; }
static inline double UBX(const orc_opts *o, const inst_t *s, int k, int j) { return o->ubx[j] - s->x[k * s->nx + j]; }

/* HPIPM residuals of the QP at (z, pi, lam, t); returns mu; norms[4] = inf-norms of res_g, res_b, res_d, res_m */
static double qp_residuals(const orc_opts *o, inst_t *s, double *norms) {
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N;
    double ng = 0, nb = 0, nd = 0, nm = 0, musum = 0; int nc = 0;
    for (int k = 0; k <= N; k++) {
        if (k < N) {
            for (int j = 0; j < nu; j++) {
                int i = k * nu + j;
                double r = o->dt * o->w[nx + j] * s->zu[i] + s->qu[i] - s->l_lbu[i] + s->l_ubu[i];
                for (int l = 0; l < nx; l++) r += s->B[(k * nx + l) * nu + j] * s->ppi[k * nx + l];
                s->rgu[i] = r; ng = dmax(ng, fabs(r));
                double dl = LBU(o, s, k, j) - s->zu[i] + s->t_lbu[i], du = s->zu[i] - UBU(o, s, k, j) + s->t_ubu[i];
                nd = dmax(nd, dmax(fabs(dl), fabs(du)));
                double ml = s->l_lbu[i] * s->t_lbu[i], mu_ = s->l_ubu[i] * s->t_ubu[i];
                nm = dmax(nm, dmax(fabs(ml), fabs(mu_))); musum += ml + mu_; nc += 2;
            }
            for (int r_ = 0; r_ < nx; r_++) {
                double r = s->b[k * nx + r_] - s->zx[(k + 1) * nx + r_];
                if (k >= 1) for (int l = 0; l < nx; l++) r += s->A[(k * nx + r_) * nx + l] * s->zx[k * nx + l];
                for (int l = 0; l < nu; l++) r += s->B[(k * nx + r_) * nu + l] * s->zu[k * nu + l];
                s->rb[k * nx + r_] = r; nb = dmax(nb, fabs(r));
            }
        }
        if (k >= 1) {
            for (int j = 0; j < nx; j++) {
                int i = k * nx + j;
                double r;
                if (k < N) {
                    r = o->dt * o->w[j] * s->zx[i] + s->qx[i] - s->ppi[(k - 1) * nx + j] - s->l_lbx[i] + s->l_ubx[i];
                    for (int l = 0; l < nx; l++) r += s->A[(k * nx + l) * nx + j] * s->ppi[k * nx + l];
                    double dl = LBX(o, s, k, j) - s->zx[i] + s->t_lbx[i], du = s->zx[i] - UBX(o, s, k, j) + s->t_ubx[i];
                    nd = dmax(nd, dmax(fabs(dl), fabs(du)));
                    double ml = s->l_lbx[i] * s->t_lbx[i], mu_ = s->l_ubx[i] * s->t_ubx[i];
                    nm = dmax(nm, dmax(fabs(ml), fabs(mu_))); musum += ml + mu_; nc += 2;
                } else {
                    r = o->w_e[j] * s->zx[i] + s->qx[i] - s->ppi[(k - 1) * nx + j];
                }
                s->rgx[i] = r; ng = dmax(ng, fabs(r));
            }
        }
    }
    norms[0] = ng; norms[1] = nb; norms[2] = nd; norms[3] = nm;
    return musum / nc;
}

/* Riccati factorisation (if fact) + solve for (dz, dpi) of the Newton system with complementarity rhs
 *   rm = res_m (predictor), or the corrected one passed through the callbacks below.
 * mode 0: rm = lam*t ; mode 1: rm = lam*t + dt_aff*dlam_aff - sigma_mu ; mode 2: rm = lam*t - sigma_mu */
static inline void bound_terms(double lam, double t, double rd, double rm, double *Gam, double *gam) {
    double ti = 1.0 / t;
    *Gam = ti * lam;
    *gam = ti * (rm - lam * rd);
}

static inline double rm_of(int mode, double lam, double t, double rd, double dz_aff_signed, double sigma_mu) {
    /* dz_aff_signed = +dz for a lower bound, -dz for an upper bound */
    double rm = lam * t;
    if (mode == 1) {
        double dt = dz_aff_signed - rd;
        double dl = -(lam * dt + rm) / t;
        rm += dt * dl - sigma_mu;
    } else if (mode == 2) rm -= sigma_mu;
    return rm;
}

static void kkt_solve(const orc_opts *o, inst_t *s, int fact, int mode, double sigma_mu) {
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N;
    double Hu[NUM], Hx[NXM];
    /* modified gradients g~ = res_g + gamma_lb - gamma_ub, and barrier diagonals (kept in P workspace on the fly) */
    /* terminal stage */
    double *PN = s->P + (size_t)N * nx * nx;
    if (fact) { memset(PN, 0, sizeof(double) * nx * nx); for (int j = 0; j < nx; j++) PN[j * nx + j] = o->w_e[j]; }
    for (int j = 0; j < nx; j++) s->pv[N * nx + j] = s->rgx[N * nx + j];
    for (int k = N - 1; k >= 0; k--) {
        const double *A = s->A + (size_t)k * nx * nx, *B = s->B + (size_t)k * nx * nu;
        const double *Pn = s->P + (size_t)(k + 1) * nx * nx, *pn = s->pv + (k + 1) * nx;
        double *Lr = s->Lr + (size_t)k * nu * nu, *K = s->K + (size_t)k * nu * nx;
        double gu[NUM], gx[NXM];
        for (int j = 0; j < nu; j++) {
            int i = k * nu + j;
            double rdl = LBU(o, s, k, j) - s->zu[i] + s->t_lbu[i], rdu = s->zu[i] - UBU(o, s, k, j) + s->t_ubu[i];
            double G1, g1, G2, g2;
            bound_terms(s->l_lbu[i], s->t_lbu[i], rdl, rm_of(mode, s->l_lbu[i], s->t_lbu[i], rdl, s->du_aff[i], sigma_mu), &G1, &g1);
            bound_terms(s->l_ubu[i], s->t_ubu[i], rdu, rm_of(mode, s->l_ubu[i], s->t_ubu[i], rdu, -s->du_aff[i], sigma_mu), &G2, &g2);
            Hu[j] = o->dt * o->w[nx + j] + G1 + G2;
            gu[j] = s->rgu[i] + g1 - g2;
        }
        if (k >= 1) for (int j = 0; j < nx; j++) {
            int i = k * nx + j;
            double rdl = LBX(o, s, k, j) - s->zx[i] + s->t_lbx[i], rdu = s->zx[i] - UBX(o, s, k, j) + s->t_ubx[i];
            double G1, g1, G2, g2;
            bound_terms(s->l_lbx[i], s->t_lbx[i], rdl, rm_of(mode, s->l_lbx[i], s->t_lbx[i], rdl, s->dx_aff[i], sigma_mu), &G1, &g1);
            bound_terms(s->l_ubx[i], s->t_ubx[i], rdu, rm_of(mode, s->l_ubx[i], s->t_ubx[i], rdu, -s->dx_aff[i], sigma_mu), &G2, &g2);
            Hx[j] = o->dt * o->w[j] + G1 + G2;
            gx[j] = s->rgx[i] + g1 - g2;
        }
        /* The solve is written as a linear recurrence in p: p_k = c_k + Phi_k' p_{k+1} with Phi_k = A + B K_k, where
         * everything that does not depend on p_{k+1} (Prb = P_{k+1} res_b_k, w = gu + B'Prb, c) is formed first; the
         * feed-forward follows once p_{k+1} is known.  Same mathematics as r~ = gu + B'(P rb + p), p_k = gx + A'(P rb + p)
         * + K'r~, re-associated so that the CUDA kernel can form Prb, w, c for all stages in parallel. */
        double Prb[NXM];
        for (int r = 0; r < nx; r++) { double a = 0; for (int l = 0; l < nx; l++) a += Pn[r * nx + l] * s->rb[k * nx + l]; Prb[r] = a; }
        double PA[NXM * NXM], PB[NXM * NUM];
        if (fact) {
            for (int r = 0; r < nx; r++) {
                for (int c = 0; c < nx; c++) { double a = 0; for (int l = 0; l < nx; l++) a += Pn[r * nx + l] * A[l * nx + c]; PA[r * nx + c] = a; }
                for (int c = 0; c < nu; c++) { double a = 0; for (int l = 0; l < nx; l++) a += Pn[r * nx + l] * B[l * nu + c]; PB[r * nu + c] = a; }
            }
            /* R~ = Hu + B'PB, Cholesky (lower) */
            double R[NUM * NUM];
            for (int r = 0; r < nu; r++) for (int c = 0; c < nu; c++) {
                double a = (r == c) ? Hu[r] : 0.0; for (int l = 0; l < nx; l++) a += B[l * nu + r] * PB[l * nu + c]; R[r * nu + c] = a; }
            for (int c = 0; c < nu; c++) {
                double d = R[c * nu + c]; for (int l = 0; l < c; l++) d -= Lr[c * nu + l] * Lr[c * nu + l];
                d = sqrt(d); Lr[c * nu + c] = d;
                for (int r = c + 1; r < nu; r++) { double a = R[r * nu + c]; for (int l = 0; l < c; l++) a -= Lr[r * nu + l] * Lr[c * nu + l]; Lr[r * nu + c] = a / d; }
            }
        }
        /* w = gu + B'Prb ; r~ = w + B'p_{k+1} ; kff = -R~^{-1} r~ */
        double wv[NUM], rt[NUM];
        for (int r = 0; r < nu; r++) { double a = gu[r]; for (int l = 0; l < nx; l++) a += B[l * nu + r] * Prb[l]; wv[r] = a; }
        for (int r = 0; r < nu; r++) { double a = wv[r]; for (int l = 0; l < nx; l++) a += B[l * nu + r] * pn[l]; rt[r] = a; }
        double kff[NUM];
        for (int r = 0; r < nu; r++) { double a = -rt[r]; for (int l = 0; l < r; l++) a -= Lr[r * nu + l] * kff[l]; kff[r] = a / Lr[r * nu + r]; }
        for (int r = nu - 1; r >= 0; r--) { double a = kff[r]; for (int l = r + 1; l < nu; l++) a -= Lr[l * nu + r] * kff[l]; kff[r] = a / Lr[r * nu + r]; }
        for (int j = 0; j < nu; j++) s->gu[k * nu + j] = kff[j];          /* feed-forward kept in gu */
        if (k >= 1) {
            double St[NUM * NXM]; /* S~ = B'PA (nu x nx) */
            if (fact) {
                for (int r = 0; r < nu; r++) for (int c = 0; c < nx; c++) { double a = 0; for (int l = 0; l < nx; l++) a += B[l * nu + r] * PA[l * nx + c]; St[r * nx + c] = a; }
                /* K = -R~^{-1} S~ */
                for (int c = 0; c < nx; c++) {
                    double y[NUM];
                    for (int r = 0; r < nu; r++) { double a = -St[r * nx + c]; for (int l = 0; l < r; l++) a -= Lr[r * nu + l] * y[l]; y[r] = a / Lr[r * nu + r]; }
                    for (int r = nu - 1; r >= 0; r--) { double a = y[r]; for (int l = r + 1; l < nu; l++) a -= Lr[l * nu + r] * y[l]; y[r] = a / Lr[r * nu + r]; }
                    for (int r = 0; r < nu; r++) K[r * nx + c] = y[r];
                }
                /* P_k = Hx + A'PA + S~'K */
                double *Pk = s->P + (size_t)k * nx * nx;
                for (int r = 0; r < nx; r++) for (int c = 0; c < nx; c++) {
                    double a = (r == c) ? Hx[r] : 0.0;
                    for (int l = 0; l < nx; l++) a += A[l * nx + r] * PA[l * nx + c];
                    for (int l = 0; l < nu; l++) a += St[l * nx + r] * K[l * nx + c];
                    Pk[r * nx + c] = a;
                }
                for (int r = 0; r < nx; r++) for (int c = 0; c < r; c++) { double a = 0.5 * (Pk[r * nx + c] + Pk[c * nx + r]); Pk[r * nx + c] = Pk[c * nx + r] = a; }
                /* Phi_k = A + B K */
                double *Phi = s->Phi + (size_t)k * nx * nx;
                for (int r = 0; r < nx; r++) for (int c = 0; c < nx; c++) { double a = A[r * nx + c]; for (int l = 0; l < nu; l++) a += B[r * nu + l] * K[l * nx + c]; Phi[r * nx + c] = a; }
            }
            /* c = gx + A'Prb + K'w ; p_k = c + Phi_k' p_{k+1} */
            const double *Phi = s->Phi + (size_t)k * nx * nx;
            for (int r = 0; r < nx; r++) {
                double a = gx[r]; for (int l = 0; l < nx; l++) a += A[l * nx + r] * Prb[l];
                for (int l = 0; l < nu; l++) a += K[l * nx + r] * wv[l];
                for (int l = 0; l < nx; l++) a += Phi[l * nx + r] * pn[l];
                s->pv[k * nx + r] = a;
            }
        }
    }
    /* forward */
    double dx[NXM]; memset(dx, 0, sizeof(dx));
    for (int k = 0; k < N; k++) {
        const double *A = s->A + (size_t)k * nx * nx, *B = s->B + (size_t)k * nx * nu, *K = s->K + (size_t)k * nu * nx;
        double du[NUM], dxn[NXM];
        /* dx_{k+1} = e_k + Phi_k dx_k with e_k = res_b_k + B kff_k ; du_k = kff_k + K_k dx_k */
        const double *Phi = s->Phi + (size_t)k * nx * nx;
        for (int r = 0; r < nu; r++) { double a = s->gu[k * nu + r]; if (k >= 1) for (int l = 0; l < nx; l++) a += K[r * nx + l] * dx[l]; du[r] = a; s->dzu[k * nu + r] = a; }
        for (int r = 0; r < nx; r++) {
            double a = s->rb[k * nx + r];
            for (int l = 0; l < nu; l++) a += B[r * nu + l] * s->gu[k * nu + l];
            if (k >= 1) for (int l = 0; l < nx; l++) a += Phi[r * nx + l] * dx[l];
            dxn[r] = a;
        }
        const double *Pn = s->P + (size_t)(k + 1) * nx * nx, *pn = s->pv + (k + 1) * nx;
        for (int r = 0; r < nx; r++) { double a = pn[r]; for (int l = 0; l < nx; l++) a += Pn[r * nx + l] * dxn[l]; s->dpi[k * nx + r] = a; }
        memcpy(dx, dxn, sizeof(double) * nx);
        memcpy(s->dzx + (k + 1) * nx, dxn, sizeof(double) * nx);
    }
}

/* after kkt_solve: walk all bounds, compute (dlam, dt) for the given rm mode and return alpha (COMPUTE_ALPHA_QP);
 * also accumulates the three sums that give mu_aff(alpha) */
typedef struct { double alpha; double s0, s1, s2; int nc; } step_info;

static inline void bound_step(int mode, double lam, double t, double rd, double dz_signed, double dz_aff_signed, double sigma_mu,
                              double *dlam, double *dt) {
    double rm = rm_of(mode, lam, t, rd, dz_aff_signed, sigma_mu);
    *dt = dz_signed - rd;
    *dlam = -(lam * (*dt) + rm) / t;
}

#define FOR_ALL_BOUNDS(...)                                                                                            \
    for (int k = 0; k < N; k++) {                                                                                      \
        for (int j = 0; j < nu; j++) {                                                                                 \
            int i = k * nu + j;                                                                                        \
            { double *lam = &s->l_lbu[i], *t = &s->t_lbu[i]; double rd = LBU(o, s, k, j) - s->zu[i] + *t, dz = s->dzu[i], dza = s->du_aff[i]; __VA_ARGS__ } \
            { double *lam = &s->l_ubu[i], *t = &s->t_ubu[i]; double rd = s->zu[i] - UBU(o, s, k, j) + *t, dz = -s->dzu[i], dza = -s->du_aff[i]; __VA_ARGS__ } \
        }                                                                                                              \
        if (k >= 1) for (int j = 0; j < nx; j++) {                                                                     \
            int i = k * nx + j;                                                                                        \
            { double *lam = &s->l_lbx[i], *t = &s->t_lbx[i]; double rd = LBX(o, s, k, j) - s->zx[i] + *t, dz = s->dzx[i], dza = s->dx_aff[i]; __VA_ARGS__ } \
            { double *lam = &s->l_ubx[i], *t = &s->t_ubx[i]; double rd = s->zx[i] - UBX(o, s, k, j) + *t, dz = -s->dzx[i], dza = -s->dx_aff[i]; __VA_ARGS__ } \
        }                                                                                                              \
    }

static step_info step_length(const orc_opts *o, inst_t *s, int mode, double sigma_mu) {
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N;
    step_info si = {1.0, 0, 0, 0, 0};
    double a_lam = -1.0, a_t = -1.0;   /* HPIPM keeps -alpha */
    FOR_ALL_BOUNDS({
        double dlam, dt; bound_step(mode, *lam, *t, rd, dz, dza, sigma_mu, &dlam, &dt);
        if (a_lam * dlam > *lam) a_lam = *lam / dlam;
        if (a_t * dt > *t) a_t = *t / dt;
        si.s0 += (*lam) * (*t); si.s1 += (*lam) * dt + (*t) * dlam; si.s2 += dlam * dt; si.nc++;
    })
    double a = a_lam > a_t ? a_lam : a_t;
    si.alpha = -a;
    return si;
}

static void update_vars(const orc_opts *o, inst_t *s, int mode, double sigma_mu, double alpha) {
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N;
    const double a = alpha * ((1.0 - alpha) * 0.99 + alpha * 0.9999999);   /* UPDATE_VAR_QP */
    /* lam, t first: they need the old z through rd */
    FOR_ALL_BOUNDS({
        double dlam, dt; bound_step(mode, *lam, *t, rd, dz, dza, sigma_mu, &dlam, &dt);
        double ln = *lam + a * dlam, tn = *t + a * dt;
        *lam = ln <= o->lam_min ? o->lam_min : ln;
        *t = tn <= o->t_min ? o->t_min : tn;
    })
    for (int i = 0; i < N * nu; i++) s->zu[i] += a * s->dzu[i];
    for (int i = nx; i < (N + 1) * nx; i++) s->zx[i] += a * s->dzx[i];
    for (int i = 0; i < N * nx; i++) s->ppi[i] += a * s->dpi[i];
}

/* HPIPM d_ocp_qp_ipm_solve, cold start.  returns HPIPM status (0 ok, 1 max iter, 2 min step, 3 NaN) */
static int qp_ipm(const orc_opts *o, inst_t *s, int *iters) {
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N;
    memset(s->zu, 0, sizeof(double) * N * nu); memset(s->zx, 0, sizeof(double) * (N + 1) * nx);
    memset(s->ppi, 0, sizeof(double) * (N + 1) * nx);
    memset(s->du_aff, 0, sizeof(double) * N * nu); memset(s->dx_aff, 0, sizeof(double) * (N + 1) * nx);
    /* INIT_VAR_OCP_QP */
#define INIT_BOUND(z, lb, ub, tl, tu, ll, lu)                                                                          \
    { double t_lb = (z) - (lb), t_ub = (ub) - (z);                                                                     \
      if (t_lb < o->thr0) { if (t_ub < o->thr0) { (z) = 0.5 * ((lb) + (ub)); t_lb = o->thr0; t_ub = o->thr0; }         \
                            else { t_lb = o->thr0; (z) = (lb) + o->thr0; } }                                           \
      else if (t_ub < o->thr0) { t_ub = o->thr0; (z) = (ub) - o->thr0; }                                               \
      (tl) = t_lb; (tu) = t_ub; (ll) = o->mu0 / t_lb; (lu) = o->mu0 / t_ub; }
    for (int k = 0; k < N; k++) {
        for (int j = 0; j < nu; j++) { int i = k * nu + j; INIT_BOUND(s->zu[i], LBU(o, s, k, j), UBU(o, s, k, j), s->t_lbu[i], s->t_ubu[i], s->l_lbu[i], s->l_ubu[i]) }
        if (k >= 1) for (int j = 0; j < nx; j++) { int i = k * nx + j; INIT_BOUND(s->zx[i], LBX(o, s, k, j), UBX(o, s, k, j), s->t_lbx[i], s->t_ubx[i], s->l_lbx[i], s->l_ubx[i]) }
    }
    double nrm[4];
    double mu = qp_residuals(o, s, nrm);
    double alpha = 1.0;
    int it = 0;
#define UNCONV (nrm[0] > o->qp_tol[0] || nrm[1] > o->qp_tol[1] || nrm[2] > o->qp_tol[2] || nrm[3] > o->qp_tol[3])
    while (it < o->qp_max_iter && alpha > o->alpha_min && UNCONV) {
        /* predictor */
        kkt_solve(o, s, 1, 0, 0.0);
        step_info si = step_length(o, s, 0, 0.0);
        alpha = si.alpha;
        double mu_aff = (si.s0 + alpha * si.s1 + alpha * alpha * si.s2) / si.nc;
        double sigma = mu_aff / mu; sigma = sigma * sigma * sigma;
        double sigma_mu = sigma * mu; if (sigma_mu < o->t_min) sigma_mu = o->t_min;
        /* corrector */
        memcpy(s->du_aff, s->dzu, sizeof(double) * N * nu); memcpy(s->dx_aff, s->dzx, sizeof(double) * (N + 1) * nx);
        kkt_solve(o, s, 0, 1, sigma_mu);
        si = step_length(o, s, 1, sigma_mu);
        alpha = si.alpha;
        int mode = 1;
        /* conditional predictor-corrector */
        double mu_aff_c = (si.s0 + alpha * si.s1 + alpha * alpha * si.s2) / si.nc;
        if (mu_aff_c > 2.0 * mu_aff) {
            kkt_solve(o, s, 0, 2, sigma_mu);
            si = step_length(o, s, 2, sigma_mu);
            alpha = si.alpha; mode = 2;
        }
        update_vars(o, s, mode, sigma_mu, alpha);
        mu = qp_residuals(o, s, nrm);
        it++;
    }
    *iters = it;
    int nan = 0;
    for (int i = 0; i < N * nu; i++) if (!isfinite(s->zu[i])) nan = 1;
    for (int i = 0; i < (N + 1) * nx; i++) if (!isfinite(s->zx[i])) nan = 1;
    if (nan) return 3;
    if (it >= o->qp_max_iter && UNCONV) return 1;
    if (alpha <= o->alpha_min) return 2;
    return 0;
}

static void linearise(const orc_opts *o, inst_t *s, const double *p) {
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N;
    double S[NXM * NSM], xn[NXM];
    for (int k = 0; k < N; k++) {
        if (o->erk_stages == 0) irk_gl4_step(o->model, s->x + k * nx, s->u + k * nu, p, o->dt, xn, S);
        else erk_step(o->model, s->x + k * nx, s->u + k * nu, p, o->dt, o->erk_stages, 1, xn, S);
        for (int r = 0; r < nx; r++) {
            for (int c = 0; c < nx; c++) s->A[(k * nx + r) * nx + c] = S[r * (nx + nu) + c];
            for (int c = 0; c < nu; c++) s->B[(k * nx + r) * nu + c] = S[r * (nx + nu) + nx + c];
            s->b[k * nx + r] = xn[r] - s->x[(k + 1) * nx + r];
        }
    }
}

/* acados ocp_nlp_res_compute; yref = [N][nx+nu] then [nx] */
static void nlp_residuals(const orc_opts *o, inst_t *s, const double *yref, const double *x0, double *res) {
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N, ny = nx + nu;
    double stat = 0, eq = 0, ineq = 0, comp = 0;
    for (int i = 0; i < N * nx; i++) eq = dmax(eq, fabs(s->b[i]));
    for (int j = 0; j < nx; j++) eq = dmax(eq, fabs(x0[j] - s->x[j]));
    if (!s->have_mult) { res[0] = INFINITY; res[1] = eq; res[2] = INFINITY; res[3] = INFINITY; return; }
    for (int k = 0; k <= N; k++) {
        if (k < N) for (int j = 0; j < nu; j++) {
            int i = k * nu + j; double v = s->u[i];
            double g = o->dt * o->w[nx + j] * (v - yref[k * ny + nx + j]) - s->lam_lbu[i] + s->lam_ubu[i];
            for (int l = 0; l < nx; l++) g += s->B[(k * nx + l) * nu + j] * s->pi[k * nx + l];
            stat = dmax(stat, fabs(g));
            ineq = dmax(ineq, dmax(dmax(o->lbu[j] - v, 0), dmax(v - o->ubu[j], 0)));
            comp = dmax(comp, dmax(fabs(s->lam_lbu[i] * (o->lbu[j] - v)), fabs(s->lam_ubu[i] * (v - o->ubu[j]))));
        }
        if (k >= 1) for (int j = 0; j < nx; j++) {
            int i = k * nx + j; double v = s->x[i], g;
            if (k < N) {
                g = o->dt * o->w[j] * (v - yref[k * ny + j]) - s->pi[(k - 1) * nx + j] - s->lam_lbx[i] + s->lam_ubx[i];
                for (int l = 0; l < nx; l++) g += s->A[(k * nx + l) * nx + j] * s->pi[k * nx + l];
                ineq = dmax(ineq, dmax(dmax(o->lbx[j] - v, 0), dmax(v - o->ubx[j], 0)));
                comp = dmax(comp, dmax(fabs(s->lam_lbx[i] * (o->lbx[j] - v)), fabs(s->lam_ubx[i] * (v - o->ubx[j]))));
            } else g = o->w_e[j] * (v - yref[N * ny + j]) - s->pi[(k - 1) * nx + j];
            stat = dmax(stat, fabs(g));
        }
    }
    res[0] = stat; res[1] = eq; res[2] = ineq; res[3] = comp;
}

/* acados SQP; returns acados status (0 ok, 1 failure/NaN input, 2 max iter, 3 min step, 4 QP failure) */
static int sqp_solve(const orc_opts *o, inst_t *s, const double *x0, const double *yref, const double *p, int *sqp_iter, int *qp_iter) {
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N, ny = nx + nu;
    *sqp_iter = 0; *qp_iter = 0;
    for (int j = 0; j < nx; j++) if (!isfinite(x0[j])) return 1;
    for (int i = 0; i < N * ny + nx; i++) if (!isfinite(yref[i])) return 1;
    const int max_it = o->rti ? 1 : o->sqp_max_iter;
    int status = 0;
    for (int it = 0; it <= max_it; it++) {
        linearise(o, s, p);
        if (!o->rti) {
            double res[4]; nlp_residuals(o, s, yref, x0, res);
            if (res[0] < o->tol[0] && res[1] < o->tol[1] && res[2] < o->tol[2] && res[3] < o->tol[3]) { status = 0; break; }
            if (it >= max_it) { status = 2; break; }
        } else if (it >= 1) break;
        /* delta QP around the iterate, x0 eliminated */
        double dx0[NXM];
        for (int j = 0; j < nx; j++) dx0[j] = x0[j] - s->x[j];
        for (int k = 0; k < N; k++) {
            for (int j = 0; j < nu; j++) s->qu[k * nu + j] = o->dt * o->w[nx + j] * (s->u[k * nu + j] - yref[k * ny + nx + j]);
            if (k >= 1) for (int j = 0; j < nx; j++) s->qx[k * nx + j] = o->dt * o->w[j] * (s->x[k * nx + j] - yref[k * ny + j]);
        }
        for (int j = 0; j < nx; j++) s->qx[N * nx + j] = o->w_e[j] * (s->x[N * nx + j] - yref[N * ny + j]);
        for (int r = 0; r < nx; r++) { double a = s->b[r]; for (int l = 0; l < nx; l++) a += s->A[r * nx + l] * dx0[l]; s->b[r] = a; }
        int qi = 0;
        int qs = qp_ipm(o, s, &qi);
        *qp_iter += qi;
        *sqp_iter = it + 1;
        if (qs != 0 && qs != 1) { status = 4; break; }
        /* full step, QP multipliers */
        for (int j = 0; j < nx; j++) s->x[j] += dx0[j];
        for (int i = 0; i < N * nu; i++) s->u[i] += s->zu[i];
        for (int i = nx; i < (N + 1) * nx; i++) s->x[i] += s->zx[i];
        memcpy(s->pi, s->ppi, sizeof(double) * N * nx);
        memcpy(s->lam_lbu, s->l_lbu, sizeof(double) * N * nu); memcpy(s->lam_ubu, s->l_ubu, sizeof(double) * N * nu);
        memcpy(s->lam_lbx, s->l_lbx, sizeof(double) * (N + 1) * nx); memcpy(s->lam_ubx, s->l_ubx, sizeof(double) * (N + 1) * nx);
        s->have_mult = 1;
        if (o->rti) status = (qs == 0) ? 0 : 2;
    }
    return status;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* exported batch API (AoS, row-major [B][...])                                                                     */

void orc_default_opts(int model, orc_opts *o) {
    memset(o, 0, sizeof(*o));
    o->model = model; o->N = 30; o->dt = 1.0 / 50; o->sqp_max_iter = 100; o->qp_max_iter = 50; o->rti = 0;
    const double MASS = 0.03277, G = 9.81, GR = G * MASS;
    for (int i = 0; i < 4; i++) { o->tol[i] = 1e-6; o->qp_tol[i] = 1e-6; }
    o->mu0 = 1.0; o->thr0 = 0.1; o->alpha_min = 1e-8; o->lam_min = 1e-16; o->t_min = 1e-16;
    const double wx[4] = {1e2, 1e2, 1.0, 1.0};
    if (model == MODEL_ATT) {     /* 3-D attitude OCP (our extension; same numbers as bnmpc_config_default) */
        o->erk_stages = 4;
        const double w[14] = {1e2, 1e2, 1e2, 1.0, 1.0, 1.0, 0.0, 10.0, 10.0, 10.0, 1e-1, 1e-1, 1e-1, 1e-1};
        for (int i = 0; i < 14; i++) o->w[i] = w[i];
        for (int i = 0; i < 10; i++) o->w_e[i] = w[i];
        const double lb[10] = {-1.2, -1.2, -1.2, -1, -1, -1, -1.5, -1.5, -1.5, -1.5};
        for (int i = 0; i < 10; i++) { o->lbx[i] = lb[i]; o->ubx[i] = -lb[i]; }
        o->lbu[0] = 0.1 * GR; o->ubu[0] = 2.0 * GR;
        for (int i = 1; i < 4; i++) { o->lbu[i] = -6.0; o->ubu[i] = 6.0; }
    } else if (model == MODEL_PLANT) {   /* thrust OCP (our extension): same state cost and boxes as the force model, u = (theta, Fd) */
        o->erk_stages = 4;
        for (int i = 0; i < 4; i++) { o->w[i] = wx[i]; o->w_e[i] = wx[i]; }
        o->w[4] = o->w[5] = 1e-1;
        const double lb[4] = {-1.2, -1.2, -1, -1}, ub[4] = {1.2, 1.2, 1, 1};
        memcpy(o->lbx, lb, sizeof(lb)); memcpy(o->ubx, ub, sizeof(ub));
        o->lbu[0] = -1.0; o->ubu[0] = 1.0; o->lbu[1] = 0.05; o->ubu[1] = 0.6;
    } else if (model == MODEL_JERK) {
        o->erk_stages = 1;
        for (int i = 0; i < 4; i++) { o->w[i] = wx[i]; o->w_e[i] = wx[i]; }
        o->w[4] = o->w[5] = 0; o->w_e[4] = o->w_e[5] = 0; o->w[6] = o->w[7] = 1e-1;
        const double lb[6] = {-1.2, -1.2, -1, -1, -5, -5 + G}, ub[6] = {1.2, 1.2, 1, 1, 5, 5 + G};
        memcpy(o->lbx, lb, sizeof(lb)); memcpy(o->ubx, ub, sizeof(ub));
        o->lbu[0] = o->lbu[1] = -5; o->ubu[0] = o->ubu[1] = 5;
    } else {
        o->erk_stages = 4;
        for (int i = 0; i < 4; i++) { o->w[i] = wx[i]; o->w_e[i] = wx[i]; }
        o->w[4] = o->w[5] = 1e-1;
        const double lb[4] = {-1.2, -1.2, -1, -1}, ub[4] = {1.2, 1.2, 1, 1};
        memcpy(o->lbx, lb, sizeof(lb)); memcpy(o->ubx, ub, sizeof(ub));
        o->lbu[0] = o->lbu[1] = -0.2 * GR; o->ubu[0] = o->ubu[1] = 1.3 * GR;
    }
}

int orc_sizeof_opts(void) { return (int)sizeof(orc_opts); }

/* ---- tiny pthread parallel-for: each worker owns one inst_t and pulls instance indices from a shared counter ---- */
typedef void (*inst_fn)(void *ctx, inst_t *s, double *scratch, int i);
typedef struct { inst_fn fn; void *ctx; int B, nx, nu, N; volatile int *next; } pf_t;

static void *pf_worker(void *arg) {
    pf_t *w = (pf_t *)arg;
    inst_t *s = inst_new(w->nx, w->nu, w->N);
    double *scratch = dal((size_t)w->N * (w->nx + w->nu) + w->nx);
    for (;;) {
        int i = __sync_fetch_and_add(w->next, 1);
        if (i >= w->B) break;
        w->fn(w->ctx, s, scratch, i);
    }
    free(scratch);
    inst_free(s);
    return NULL;
}

static void parallel_for(int B, int nthreads, int nx, int nu, int N, inst_fn fn, void *ctx) {
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads > B) nthreads = B;
    if (nthreads < 1) nthreads = 1;
    volatile int next = 0;
    pf_t w = {fn, ctx, B, nx, nu, N, &next};
    if (nthreads == 1) { pf_worker(&w); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, pf_worker, &w);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
}

int orc_num_cores(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }

typedef struct {
    const orc_opts *o; const double *x0, *yref, *p; double *x, *u, *pi, *lam_out; int *status, *sqp_iter, *qp_iter;
} solve_ctx;

static void solve_one(void *vc, inst_t *s, double *scratch, int i) {
    solve_ctx *c = (solve_ctx *)vc;
    const orc_opts *o = c->o;
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N, ny = nx + nu;
    const size_t X = (size_t)(N + 1) * nx, U = (size_t)N * nu;
    (void)scratch;
    memcpy(s->x, c->x + i * X, sizeof(double) * X); memcpy(s->u, c->u + i * U, sizeof(double) * U);
    s->have_mult = 0;
    c->status[i] = sqp_solve(o, s, c->x0 + (size_t)i * nx, c->yref + (size_t)i * (N * ny + nx), c->p + (size_t)i * 2,
                             &c->sqp_iter[i], &c->qp_iter[i]);
    memcpy(c->x + i * X, s->x, sizeof(double) * X); memcpy(c->u + i * U, s->u, sizeof(double) * U);
    if (c->pi) memcpy(c->pi + (size_t)i * N * nx, s->pi, sizeof(double) * N * nx);
    if (c->lam_out) {
        double *l = c->lam_out + (size_t)i * 2 * (U + X);
        memcpy(l, s->lam_lbu, sizeof(double) * U); memcpy(l + U, s->lam_ubu, sizeof(double) * U);
        memcpy(l + 2 * U, s->lam_lbx, sizeof(double) * X); memcpy(l + 2 * U + X, s->lam_ubx, sizeof(double) * X);
    }
}

/* One solve per instance from the given iterate (x,u in/out).  lam_out: [B][2*(N*nu + (N+1)*nx)] = lbu|ubu|lbx|ubx */
int orc_solve_batch(const orc_opts *o, int B, const double *x0, const double *yref, const double *p, double *x, double *u,
                    double *pi, double *lam_out, int *status, int *sqp_iter, int *qp_iter, int nthreads) {
    int nx, nu; model_dims(o->model, &nx, &nu);
    solve_ctx c = {o, x0, yref, p, x, u, pi, lam_out, status, sqp_iter, qp_iter};
    parallel_for(B, nthreads, nx, nu, o->N, solve_one, &c);
    return 0;
}

/* AcadosSimSolver of the plant: x_next = ERK(ns stages, nsub substeps of T each with its own input) (no noise) */
int orc_sim_batch(int B, int ns, int nsub, double T, const double *x, const double *u, const double *p, double *xn) {
    for (int i = 0; i < B; i++) {
        double xi[4]; memcpy(xi, x + (size_t)i * 4, sizeof(xi));
        for (int j = 0; j < nsub; j++) erk_step(MODEL_PLANT, xi, u + ((size_t)i * nsub + j) * 2, p + (size_t)i * 2, T, ns, 1, xi, NULL);
        memcpy(xn + (size_t)i * 4, xi, sizeof(xi));
    }
    return 0;
}

/* the same for any model as its own plant (MODEL_ATT: x [B][10], u [B][nsub][4]) */
int orc_sim_batch_model(int model, int B, int ns, int nsub, double T, const double *x, const double *u, const double *p, double *xn) {
    int nx, nu; model_dims(model, &nx, &nu);
    for (int i = 0; i < B; i++) {
        double xi[NXM]; memcpy(xi, x + (size_t)i * nx, sizeof(double) * nx);
        for (int j = 0; j < nsub; j++) erk_step(model, xi, u + ((size_t)i * nsub + j) * nu, p + (size_t)i * 2, T, ns, 1, xi, NULL);
        memcpy(xn + (size_t)i * nx, xi, sizeof(double) * nx);
    }
    return 0;
}

typedef struct {
    const orc_opts *o; int B, n_steps, rows, ref_shared; const double *ref, *x0, *noise, *p_ctrl, *p_plant;
    double *Xsim, *U_plant, *U_ctrl, *a_log, *cost; int *status, *qp_iter;
} cl_ctx;

static void closed_loop_one(void *vc, inst_t *s, double *yref, int i) {
    cl_ctx *c = (cl_ctx *)vc;
    const orc_opts *o = c->o;
    const int nx = ORC_NX(s->nx), nu = ORC_NU(s->nu), N = s->N, ny = nx + nu, n_steps = c->n_steps, B = c->B;
    const double wc[4] = {1e2, 1e2, 1.0, 1.0};
    const double *rt = c->ref + (c->ref_shared ? 0 : (size_t)i * c->rows * 8);
    const double *pc = c->p_ctrl + (size_t)i * 2, *pp = c->p_plant + (size_t)i * 2;
    memset(s->x, 0, sizeof(double) * (N + 1) * nx); memset(s->u, 0, sizeof(double) * N * nu); s->have_mult = 0;
    double xs[4]; memcpy(xs, c->x0 + (size_t)i * 4, sizeof(xs));
    double ai[2] = {0.0, pc[1]};       /* jerk_model/controller.py:23 (hover) */
    double csum = 0;
    if (c->Xsim) memcpy(c->Xsim + (size_t)i * (n_steps + 1) * 4, xs, sizeof(xs));
    for (int st = 0; st < n_steps; st++) {
        /* set_up_ocp: yref_k = [xref[st+k], uref[st+k]]; force: ref[:, :4] | ref[:, 4:6]; jerk: ref[:, :6] | ref[:, 6:] */
        for (int k = 0; k < N; k++) for (int j = 0; j < ny; j++) yref[k * ny + j] = rt[(size_t)(st + k) * 8 + j];
        for (int j = 0; j < nx; j++) yref[N * ny + j] = rt[(size_t)(st + N) * 8 + j];
        double x0b[NXM]; memcpy(x0b, xs, sizeof(xs));
        if (o->model == MODEL_JERK) { x0b[4] = ai[0]; x0b[5] = ai[1]; }
        int si_, qi_;
        int stt = sqp_solve(o, s, x0b, yref, pc, &si_, &qi_);
        if (c->status) c->status[(size_t)i * n_steps + st] = stt;
        if (c->qp_iter) c->qp_iter[(size_t)i * n_steps + st] = qi_;
        const double *u0 = s->u;
        const double *xo = (o->model == MODEL_JERK) ? s->x + nx : s->x;
        for (int j = 0; j < 4; j++) { double d = xo[j] - rt[(size_t)st * 8 + j]; csum += wc[j] * d * d; }
        double up[10][2]; int nsub; double xn[4];
        size_t o2 = ((size_t)i * n_steps + st) * 2;
        if (o->model == MODEL_JERK) {
            nsub = 10;
            for (int j = 0; j < 10; j++) {
                ai[0] += u0[0] * (1.0 / 500); ai[1] += u0[1] * (1.0 / 500);
                double fx = pc[0] * ai[0], fz = pc[0] * ai[1];
                up[j][0] = atan2(fx, fz); up[j][1] = sqrt(fx * fx + fz * fz);
            }
            memcpy(xn, xs, sizeof(xs));
            for (int j = 0; j < 10; j++) erk_step(MODEL_PLANT, xn, up[j], pp, 1.0 / 500, 1, 1, xn, NULL);
            if (c->a_log) { c->a_log[o2] = ai[0]; c->a_log[o2 + 1] = ai[1]; }
        } else if (o->model == MODEL_PLANT) {   /* thrust OCP: u0 already is the plant input */
            nsub = 1;
            up[0][0] = u0[0]; up[0][1] = u0[1];
            erk_step(MODEL_PLANT, xs, up[0], pp, o->dt, 4, 1, xn, NULL);
            if (c->a_log) { c->a_log[o2] = u0[1] * sin(u0[0]) / 0.03277; c->a_log[o2 + 1] = u0[1] * cos(u0[0]) / 0.03277; }
        } else {
            nsub = 1;
            up[0][0] = atan2(u0[0], u0[1]); up[0][1] = sqrt(u0[0] * u0[0] + u0[1] * u0[1]);
            erk_step(MODEL_PLANT, xs, up[0], pp, o->dt, 4, 1, xn, NULL);
            if (c->a_log) { c->a_log[o2] = u0[0] / 0.03277; c->a_log[o2 + 1] = u0[1] / 0.03277; }
        }
        const double eps = c->noise ? c->noise[(size_t)st * B + i] : 0.0;
        for (int j = 0; j < 4; j++) xs[j] = xn[j] + eps;
        if (c->U_ctrl) { c->U_ctrl[o2] = u0[0]; c->U_ctrl[o2 + 1] = u0[1]; }
        if (c->U_plant) { c->U_plant[o2] = up[nsub - 1][0]; c->U_plant[o2 + 1] = up[nsub - 1][1]; }
        if (c->Xsim) memcpy(c->Xsim + ((size_t)i * (n_steps + 1) + st + 1) * 4, xs, sizeof(xs));
    }
    if (c->cost) c->cost[i] = csum;
}

/* Closed loop of follow_trajectory for B instances.
 *   ref      [B][rows][8]  (or one shared table when ref_shared != 0), rows >= n_steps + N
 *   x0       [B][4];  noise [n_steps][B];  p_ctrl, p_plant [B][2]
 *   outputs (any may be NULL): Xsim [B][n_steps+1][4], U_plant [B][n_steps][2], U_ctrl [B][n_steps][2], a_log [B][n_steps][2],
 *   cost [B], status [B][n_steps], qp_iter [B][n_steps] */
int orc_closed_loop(const orc_opts *o, int B, int n_steps, int rows, const double *ref, int ref_shared, const double *x0,
                    const double *noise, const double *p_ctrl, const double *p_plant, double *Xsim, double *U_plant,
                    double *U_ctrl, double *a_log, double *cost, int *status, int *qp_iter, int nthreads) {
    int nx, nu; model_dims(o->model, &nx, &nu);
    if (rows < n_steps + o->N) return -1;
    cl_ctx c = {o, B, n_steps, rows, ref_shared, ref, x0, noise, p_ctrl, p_plant, Xsim, U_plant, U_ctrl, a_log, cost, status, qp_iter};
    parallel_for(B, nthreads, nx, nu, o->N, closed_loop_one, &c);
    return 0;
}

/* Closed loop of the 3-D attitude model (follow_trajectory with the controller model as its own plant, the shape of
 * src/force_model/controller.py:25-54): per step yref window from ref, x0 embedding, solve, u0 held over one ERK4 plant step of
 * length dt with the PLANT parameters, eps added to position and velocity.
 *   ref [B][rows][14] = [x (10); u (4)] per row (or one shared table); x0 [B][10]; noise [n_steps][B] or NULL; p_* [B][2]
 *   outputs (any may be NULL): Xsim [B][n_steps+1][10], U_ctrl [B][n_steps][4], cost [B], status / qp_iter / sqp_iter [B][n_steps] */
typedef struct {
    const orc_opts *o; int B, n_steps, rows, ref_shared; const double *ref, *x0, *noise, *p_ctrl, *p_plant;
    double *Xsim, *U_ctrl, *cost; int *status, *qp_iter, *sqp_iter;
} cla_ctx;

static void closed_loop_att_one(void *vc, inst_t *s, double *yref, int i) {
    cla_ctx *c = (cla_ctx *)vc;
    const orc_opts *o = c->o;
    const int nx = s->nx, nu = s->nu, N = s->N, ny = nx + nu, n_steps = c->n_steps, B = c->B;
    const double *rt = c->ref + (c->ref_shared ? 0 : (size_t)i * c->rows * ny);
    const double *pc = c->p_ctrl + (size_t)i * 2, *pp = c->p_plant + (size_t)i * 2;
    memset(s->x, 0, sizeof(double) * (N + 1) * nx); memset(s->u, 0, sizeof(double) * N * nu); s->have_mult = 0;
    for (int k = 0; k <= N; k++) s->x[k * nx + 6] = 1.0;      /* start iterate: identity attitude, hover thrust */
    for (int k = 0; k < N; k++) s->u[k * nu] = pc[0] * pc[1];
    double xs[NXM]; memcpy(xs, c->x0 + (size_t)i * nx, sizeof(double) * nx);
    double csum = 0;
    if (c->Xsim) memcpy(c->Xsim + (size_t)i * (n_steps + 1) * nx, xs, sizeof(double) * nx);
    for (int st = 0; st < n_steps; st++) {
        for (int k = 0; k < N; k++) for (int j = 0; j < ny; j++) yref[k * ny + j] = rt[(size_t)(st + k) * ny + j];
        for (int j = 0; j < nx; j++) yref[N * ny + j] = rt[(size_t)(st + N) * ny + j];
        int si_, qi_;
        int stt = sqp_solve(o, s, xs, yref, pc, &si_, &qi_);
        if (c->status) c->status[(size_t)i * n_steps + st] = stt;
        if (c->qp_iter) c->qp_iter[(size_t)i * n_steps + st] = qi_;
        if (c->sqp_iter) c->sqp_iter[(size_t)i * n_steps + st] = si_;
        for (int j = 0; j < 6; j++) { double d = s->x[j] - rt[(size_t)st * ny + j]; csum += (j < 3 ? 1e2 : 1.0) * d * d; }
        double xn[NXM];
        erk_step(MODEL_ATT, xs, s->u, pp, o->dt, 4, 1, xn, NULL);
        const double eps = c->noise ? c->noise[(size_t)st * B + i] : 0.0;
        for (int j = 0; j < nx; j++) xs[j] = xn[j] + (j < 6 ? eps : 0.0);
        if (c->U_ctrl) memcpy(c->U_ctrl + ((size_t)i * n_steps + st) * nu, s->u, sizeof(double) * nu);
        if (c->Xsim) memcpy(c->Xsim + ((size_t)i * (n_steps + 1) + st + 1) * nx, xs, sizeof(double) * nx);
    }
    if (c->cost) c->cost[i] = csum;
}

int orc_closed_loop_att(const orc_opts *o, int B, int n_steps, int rows, const double *ref, int ref_shared, const double *x0,
                        const double *noise, const double *p_ctrl, const double *p_plant, double *Xsim, double *U_ctrl, double *cost,
                        int *status, int *qp_iter, int *sqp_iter, int nthreads) {
    int nx, nu; model_dims(o->model, &nx, &nu);
    if (o->model != MODEL_ATT || rows < n_steps + o->N) return -1;
    cla_ctx c = {o, B, n_steps, rows, ref_shared, ref, x0, noise, p_ctrl, p_plant, Xsim, U_ctrl, cost, status, qp_iter, sqp_iter};
    parallel_for(B, nthreads, nx, nu, o->N, closed_loop_att_one, &c);
    return 0;
}
