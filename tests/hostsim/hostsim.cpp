// hostsim - TEST HARNESS ONLY.  Compiles the solver arithmetic of the product (csrc/bnmpc_core.cuh, bnmpc_loop.cuh,
// the same __host__ __device__ templates the CUDA kernels instantiate) for the host CPU so that the device code's logic
// can be checked against the oracle in the CPU-only test run.  It is NOT part of libbnmpc.so and nothing in the
// product package can reach it: the product has no CPU path.  The warp that owns an instance on the GPU is emulated
// by a single "lane" (group size 1) that walks all (stage, block) items in order.
#define __host__
#define __device__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#include <string.h>
#include <stdlib.h>
#include <vector>

#include "../../drone_attitude_control_b200/csrc/bnmpc_lockstep.cuh"

using namespace bnmpc;

struct HostGroup {
    static constexpr int L = 1;
    static constexpr bool PAR_SCAN = false;
    int lane = 0;
    template <class T> T max(T v) const { return v; }
    template <class T> T sum(T v) const { return v; }
    bool any(bool p) const { return p; }
    bool all(bool p) const { return p; }
    void sync() const {}
};

template <class M, class T>
struct Sim {
    Gs<T> gs; std::vector<T> buf, sm; std::vector<int32_t> ints; std::vector<double> u0; int B, N;
    Sim(int B_, int N_) : B(B_), N(N_) {
        constexpr int SG = M::NU + M::NX;
        const size_t nV = (size_t)(N + 1) * SG, nPI = (size_t)N * M::NX, nL = (size_t)N * 2 * SG;
        const size_t nY = (size_t)N * SG + M::NX;
        buf.assign((size_t)B * (nV + nY + nPI + nL + M::NX + M::NP), T(0));
        T* p = buf.data();
        gs.V = p; p += B * nV; gs.PI = p; p += B * nPI; gs.LAM = p; p += B * nL; gs.YREF = p; p += B * nY; gs.X0 = p; p += (size_t)B * M::NX; gs.PAR = p;
        ints.assign((size_t)4 * B, 0);
        gs.status = ints.data(); gs.sqp_iter = gs.status + B; gs.qp_iter = gs.sqp_iter + B; gs.have_mult = gs.qp_iter + B;
        gs.B = B; gs.N = N;
        u0.assign((size_t)B * M::NU, 0.0); gs.U0 = u0.data(); gs.BND = nullptr;
        sm.assign(SmLayout<M, true>::elems(N), T(0));
    }
    void set(int inst, int field, int k, const double* v) {
        const int dim = field_dim(M::NX, M::NU, M::NP, field, k, N);
        for (int j = 0; j < dim; j++) *field_ptr(gs, M::NX, M::NU, M::NP, inst, field, k, j) = T(v[j]);
    }
    void get(int inst, int field, int k, double* v) {
        const int dim = field_dim(M::NX, M::NU, M::NP, field, k, N);
        for (int j = 0; j < dim; j++) v[j] = double(*field_ptr(gs, M::NX, M::NU, M::NP, inst, field, k, j));
    }
};

template <class M, class T>
static int solve_batch_t(const Opts* o, int B, const double* x0, const double* yref, const double* p, double* x, double* u,
                         double* pi, int* status, int* sqp_iter, int* qp_iter, const double* bnd) {
    constexpr int NX = M::NX, NU = M::NU, ny = NX + NU;
    const int N = o->N;
    Sim<M, T> sim(B, N);
    sim.gs.BND = const_cast<double*>(bnd);          // per-stage bounds [B][N][2][NU+NX] or NULL
    HostGroup g;
    for (int i = 0; i < B; i++) {
        for (int k = 0; k <= N; k++) sim.set(i, F_X, k, x + ((size_t)i * (N + 1) + k) * NX);
        for (int k = 0; k < N; k++) sim.set(i, F_U, k, u + ((size_t)i * N + k) * NU);
        for (int k = 0; k <= N; k++) sim.set(i, F_YREF, k, yref + (size_t)i * (N * ny + NX) + (size_t)k * ny);
        sim.set(i, F_LBX, 0, x0 + (size_t)i * NX);
        sim.set(i, F_P, 0, p + (size_t)i * 2);
        const SmemPriv<M, T> ps;
        Solver<M, T, HostGroup, SmemPriv<M, T>> sv(sim.sm.data(), 0, *o, g, ps);
        api_solve<M, T>(sv, i, sim.gs);
        for (int k = 0; k <= N; k++) sim.get(i, F_X, k, x + ((size_t)i * (N + 1) + k) * NX);
        for (int k = 0; k < N; k++) sim.get(i, F_U, k, u + ((size_t)i * N + k) * NU);
        if (pi) for (int k = 0; k < N; k++) sim.get(i, F_PI, k, pi + ((size_t)i * N + k) * NX);
        status[i] = sim.gs.status[i]; sqp_iter[i] = sim.gs.sqp_iter[i]; qp_iter[i] = sim.gs.qp_iter[i];
    }
    return 0;
}

template <class M, class T>
static int closed_loop_t(const Opts* o, int kind, int B, int n_steps, int rows, const double* ref, int ref_shared, const double* x0,
                         const double* noise, const double* p_ctrl, const double* p_plant, double* Xsim, double* U_plant,
                         double* U_ctrl, double* a_log, double* cost, double* abs_err, int* status, int* qp_iter) {
    const int N = o->N;
    if (rows < n_steps + N) return -1;
    Sim<M, T> sim(B, N);
    HostGroup g;
    const size_t Bp = B;
    std::vector<double> xs(4 * Bp), acc(2 * Bp), pp(2 * Bp);
    for (int i = 0; i < B; i++) {
        for (int j = 0; j < 4; j++) { xs[j * Bp + i] = x0[j * B + i]; if (Xsim) Xsim[(size_t)j * B + i] = x0[j * B + i]; }
        acc[i] = 0.0; acc[Bp + i] = p_ctrl[B + i];
        pp[i] = p_plant[i]; pp[Bp + i] = p_plant[B + i];
        cost[i] = 0; abs_err[i] = 0;
        const double pc[2] = {p_ctrl[i], p_ctrl[B + i]};
        sim.set(i, F_P, 0, pc);
    }
    for (int i = 0; i < B; i++) {
        const SmemPriv<M, T> ps;
        Solver<M, T, HostGroup, SmemPriv<M, T>> sv(sim.sm.data(), 0, *o, g, ps);
        for (int st = 0; st < n_steps; st++) {
            LoopArgs a{};
            a.step = st; a.kind = kind; a.ref_layout = ref_shared; a.ref_rows = rows; a.log_stride = n_steps; a.batch = B; a.Bp = Bp;
            a.ref = ref; a.noise = noise; a.Xsim = Xsim; a.U_plant = U_plant; a.U_ctrl = U_ctrl; a.a_log = a_log;
            a.status = status; a.qp_iter = qp_iter; a.xs = xs.data(); a.acc = acc.data(); a.cost = cost; a.abs_err = abs_err;
            a.p_plant = pp.data();
            closed_loop_step<M, T>(sv, i, sim.gs, a);
        }
    }
    return 0;
}

// The slotted lockstep schedule of bnmpc_lockstep.cuh with one emulated CTA of W "warps": the pass slots run warp after
// warp (the CTA barriers of the kernel are implicit), the sweep slots call the instance's own sweep functions.  Checks the
// state machine (ls_pslot), the ticket / chunk hand-over and the multi-step residency against the per-step path.
struct HostTickets {
    int next = 0; std::vector<int>* ns;
    int take() { return next++; }
    bool ready(int inst, int step) const { return (*ns)[inst] == step; }
    void wait(int inst, int step) const { if (!ready(inst, step)) abort(); }    // tickets run in order here: never blocks
    void publish(int inst, int nx) { (*ns)[inst] = nx; }
};

template <class M, class T>
static int closed_loop_ls_t(const Opts* o, int kind, int B, int n_steps, int chunk, int W, int rows, const double* ref, int ref_shared,
                            const double* x0, const double* noise, const double* p_ctrl, const double* p_plant, double* Xsim,
                            double* U_plant, double* U_ctrl, double* a_log, double* cost, double* abs_err, int* status, int* qp_iter,
                            int* fail_count, unsigned long long philox_seed, double philox_std, long long inst0) {
    const int N = o->N;
    if (rows < n_steps + N) return -1;
    using SV = Solver<M, T, HostGroup, SmemPriv<M, T>>;
    Sim<M, T> sim(B, N);
    HostGroup g;
    const size_t Bp = B;
    std::vector<double> xs(4 * Bp), acc(2 * Bp), pp(2 * Bp);
    std::vector<int> next_step(B, 0);
    for (int i = 0; i < B; i++) {
        for (int j = 0; j < 4; j++) { xs[j * Bp + i] = x0[j * B + i]; if (Xsim) Xsim[(size_t)j * B + i] = x0[j * B + i]; }
        acc[i] = 0.0; acc[Bp + i] = p_ctrl[B + i];
        pp[i] = p_plant[i]; pp[Bp + i] = p_plant[B + i];
        cost[i] = 0; abs_err[i] = 0; if (fail_count) fail_count[i] = 0;
        const double pc[2] = {p_ctrl[i], p_ctrl[B + i]};
        sim.set(i, F_P, 0, pc);
    }
    LoopArgs a{};
    a.step = 0; a.n_steps = n_steps; a.chunk = chunk; a.kind = kind; a.ref_layout = ref_shared; a.ref_rows = rows; a.log_stride = n_steps;
    a.batch = B; a.Bp = Bp; a.ref = ref; a.noise = noise; a.Xsim = Xsim; a.U_plant = U_plant; a.U_ctrl = U_ctrl; a.a_log = a_log;
    a.status = status; a.qp_iter = qp_iter; a.xs = xs.data(); a.acc = acc.data(); a.cost = cost; a.abs_err = abs_err; a.p_plant = pp.data();
    a.fail_count = fail_count; a.next_step = next_step.data();
    if (philox_std > 0) { a.noise = nullptr; a.noise_philox = 1; a.noise_seed = philox_seed; a.noise_std = philox_std; a.inst0 = inst0; }
    const SmemPriv<M, T> ps;
    if (W == 0) {      // the free-running multi-step path of k_loop_step: one group, tickets in order
        std::vector<T> sm1(SmLayout<M, true>::elems(N), T(0));
        SV sv1(sm1.data(), 0, *o, g, ps);
        HostTickets wq1; wq1.ns = &next_step;
        const int total = B * ((n_steps + chunk - 1) / chunk);
        for (int t = wq1.take(); t < total; t = wq1.take()) closed_loop_chunk<M, T>(sv1, t, sim.gs, a, wq1);
        return 0;
    }
    std::vector<std::vector<T>> sm(W, std::vector<T>(SmLayout<M, true>::elems(N), T(0)));
    std::vector<SV> sv; sv.reserve(W);
    std::vector<LsWarp<M, T, HostGroup, SmemPriv<M, T>>> w(W);
    for (int i = 0; i < W; i++) { sv.emplace_back(sm[i].data(), 0, *o, g, ps); w[i].reset(); }
    HostTickets wq; wq.ns = &next_step;
    std::vector<int> ws(W);
    for (long half = 0;; half++) {
        if (half > 400L * n_steps * ((B + W - 1) / W) + 1000) return -3;      // the schedule must terminate
        for (int i = 0; i < W; i++) ls_pslot<M, T>(1, w[i], sv[i], sim.gs, a, wq);
        bool alive = false;
        for (int i = 0; i < W; i++) { ws[i] = w[i].published(); alive = alive || (ws[i] & LS_ALIVE); }
        if (!alive) break;
        for (int i = 0; i < W; i++) if ((ws[i] & 7) == 1) sv[i].kkt_factor();
        for (int i = 0; i < W; i++) ls_pslot<M, T>(2, w[i], sv[i], sim.gs, a, wq);
        for (int i = 0; i < W; i++) if ((ws[i] & 7) != 0) sv[i].back_scan();
        for (int i = 0; i < W; i++) ls_pslot<M, T>(3, w[i], sv[i], sim.gs, a, wq);
        for (int i = 0; i < W; i++) if ((ws[i] & 7) != 0) sv[i].fwd_scan((ws[i] & 7) - 1);
        for (int i = 0; i < W; i++) ls_pslot<M, T>(4, w[i], sv[i], sim.gs, a, wq);
    }
    for (int i = 0; i < B; i++) if (next_step[i] != n_steps) return -4;
    return 0;
}

#ifdef HS_ONLY_ATT      // quick build while working on the 3-D model: only its instantiations
#define DISPATCH(fn, model, prec, ...) return -2;
#else
#define DISPATCH(fn, model, prec, ...)                                                     \
    switch ((model) * 2 + (prec)) {                                                        \
    case 0: return fn<Model_force, double>(__VA_ARGS__);                                   \
    case 1: return fn<Model_force, float>(__VA_ARGS__);                                    \
    case 2: return fn<Model_jerk, double>(__VA_ARGS__);                                    \
    case 3: return fn<Model_jerk, float>(__VA_ARGS__);                                     \
    case 4: return fn<Model_force_dense, double>(__VA_ARGS__);                             \
    case 5: return fn<Model_force_dense, float>(__VA_ARGS__);                              \
    case 6: return fn<Model_jerk_dense, double>(__VA_ARGS__);                              \
    case 7: return fn<Model_jerk_dense, float>(__VA_ARGS__);                               \
    case 8: return fn<Model_plant, double>(__VA_ARGS__);                                   \
    case 9: return fn<Model_plant, float>(__VA_ARGS__);                                    \
    }                                                                                      \
    return -2;
#endif

extern "C" {
int hs_sizeof_opts(void) { return (int)sizeof(Opts); }
// table [B][rows][8] of the on-the-fly circle reference (params [B][4]), n = samples per revolution
void hs_circle_table(int B, int rows, int n, const double* prm, double* out) {
    for (int i = 0; i < B; i++) for (int r = 0; r < rows; r++) for (int c = 0; c < 8; c++)
        out[((size_t)i * rows + r) * 8 + c] = circle_ref(prm + (size_t)i * 4, r, c, n);
}
// AoS in/out like the C oracle's orc_solve_batch: x [B][N+1][NX], u [B][N][NU] (start iterate in, solution out)
int hs_solve_batch(int model, int prec, const Opts* o, int B, const double* x0, const double* yref, const double* p, double* x,
                   double* u, double* pi, int* status, int* sqp_iter, int* qp_iter, const double* bnd) {
    // the 3-D attitude model has the solve() surface only (no fused closed loop)
    if (model == 5) return prec ? solve_batch_t<Model_att, float>(o, B, x0, yref, p, x, u, pi, status, sqp_iter, qp_iter, bnd)
                                : solve_batch_t<Model_att, double>(o, B, x0, yref, p, x, u, pi, status, sqp_iter, qp_iter, bnd);
    DISPATCH(solve_batch_t, model, prec, o, B, x0, yref, p, x, u, pi, status, sqp_iter, qp_iter, bnd)
}
// batch-minor in/out exactly like bnmpc_closed_loop_* (x0 [4][B], p_* [2][B], ref shared [rows][8] or [rows][8][B], logs [steps][dim][B])
int hs_closed_loop(int model, int prec, const Opts* o, int B, int n_steps, int rows, const double* ref, int ref_shared,
                   const double* x0, const double* noise, const double* p_ctrl, const double* p_plant, double* Xsim, double* U_plant,
                   double* U_ctrl, double* a_log, double* cost, double* abs_err, int* status, int* qp_iter) {
    const int kind = (model == 1 || model == 3) ? KIND_JERK : (model == 4 ? KIND_THRUST : KIND_FORCE);
    DISPATCH(closed_loop_t, model, prec, o, kind, B, n_steps, rows, ref, ref_shared, x0, noise, p_ctrl, p_plant, Xsim, U_plant, U_ctrl,
             a_log, cost, abs_err, status, qp_iter)
}
// the same closed loop through the slotted lockstep schedule (one emulated CTA of W warps, tickets of `chunk` steps);
// philox_std > 0: the noise is drawn by philox_normal(seed, inst0 + i, step) instead of read from `noise`
int hs_closed_loop_ls(int model, int prec, const Opts* o, int B, int n_steps, int chunk, int W, int rows, const double* ref, int ref_shared,
                      const double* x0, const double* noise, const double* p_ctrl, const double* p_plant, double* Xsim, double* U_plant,
                      double* U_ctrl, double* a_log, double* cost, double* abs_err, int* status, int* qp_iter, int* fail_count,
                      unsigned long long philox_seed, double philox_std, long long inst0) {
    const int kind = (model == 1 || model == 3) ? KIND_JERK : (model == 4 ? KIND_THRUST : KIND_FORCE);
    DISPATCH(closed_loop_ls_t, model, prec, o, kind, B, n_steps, chunk, W, rows, ref, ref_shared, x0, noise, p_ctrl, p_plant, Xsim, U_plant,
             U_ctrl, a_log, cost, abs_err, status, qp_iter, fail_count, philox_seed, philox_std, inst0)
}
// noise_std * N(0,1) of the device-side generator, [n_steps][B]
void hs_philox_noise(int B, int n_steps, int first_step, unsigned long long seed, double std_, long long inst0, double* out) {
    for (int s = 0; s < n_steps; s++) for (int i = 0; i < B; i++) out[(size_t)s * B + i] = std_ * philox_normal(seed, inst0 + i, first_step + s);
}
void hs_philox4x32(const unsigned* ctr, const unsigned* key, unsigned* out) { philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out); }
}
