// hostsim - TEST HARNESS ONLY.  Compiles the solver arithmetic of the product (csrc/bnmpc_core.cuh, bnmpc_loop.cuh,
// the same __host__ __device__ templates the CUDA kernels instantiate) for the host CPU so that the device code's logic
// can be checked against the oracle in the CPU-only test run.  It is NOT part of libbnmpc.so and nothing in the
// product package can reach it: the product has no CPU path.  The NBLK threads of an instance (warp lanes on the
// GPU) are emulated by NBLK host threads that exchange the interior-point scalars through a barrier.
#define __host__
#define __device__
#define __forceinline__ inline __attribute__((always_inline))
#include <pthread.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <thread>

#include "../../drone_attitude_control_b200/csrc/bnmpc_loop.cuh"

using namespace bnmpc;

template <int NBLK>
struct HostShared { pthread_barrier_t bar; double slot[NBLK]; };

template <int NBLK>
struct HostXchg {
    HostShared<NBLK>* sh; int b;
    template <class T> T exchange(T v, int mode, int src = 0) const {
        if (NBLK == 1) return v;
        sh->slot[b] = (double)v;
        pthread_barrier_wait(&sh->bar);
        double r = sh->slot[0];
        if (mode == 0) { for (int i = 1; i < NBLK; i++) r = sh->slot[i] > r ? sh->slot[i] : r; }
        else if (mode == 1) { r = 0; for (int i = 0; i < NBLK; i++) r += sh->slot[i]; }
        else r = sh->slot[src];
        pthread_barrier_wait(&sh->bar);
        return (T)r;
    }
    template <class T> T max(T v) const { return exchange(v, 0); }
    template <class T> T sum(T v) const { return exchange(v, 1); }
    bool any_in_instance(bool p) const { return exchange((double)p, 0) != 0.0; }
    bool any_in_group(bool p) const { return any_in_instance(p); }
    template <class T> T from_block(T v, int src) const { return exchange(v, 2, src); }
    void sync() const { if (NBLK > 1) pthread_barrier_wait(&sh->bar); }
};

template <class M, class T>
struct Sim {
    Ws<T> w; std::vector<T> buf; std::vector<int32_t> ints; int B, N;
    Sim(int B_, int N_) : B(B_), N(N_) {
        w.B = B; w.S = (size_t)((B * M::NBLK + 31) / 32) * 32;
        const int rows = WsLayout<M>::fill(N, w.off);
        buf.assign((size_t)rows * w.S, T(0)); w.base = buf.data();
        ints.assign((size_t)4 * B, 0);
        w.status = ints.data(); w.sqp_iter = w.status + B; w.qp_iter = w.sqp_iter + B; w.have_mult = w.qp_iter + B;
    }
};

// run fn(inst, b, xchg) for every (instance, block); blocks of one instance run as concurrent host threads
template <class M, class F>
static void for_all(int B, F fn) {
    constexpr int NBLK = M::NBLK;
    for (int i = 0; i < B; i++) {
        HostShared<NBLK> sh;
        if (NBLK > 1) pthread_barrier_init(&sh.bar, NULL, NBLK);
        std::vector<std::thread> th;
        for (int b = 1; b < NBLK; b++) th.emplace_back([&, b] { HostXchg<NBLK> xc{&sh, b}; fn(i, b, xc); });
        HostXchg<NBLK> xc{&sh, 0};
        fn(i, 0, xc);
        for (auto& t : th) t.join();
        if (NBLK > 1) pthread_barrier_destroy(&sh.bar);
    }
}

template <class M, class T>
static int solve_batch_t(const Opts* o, int B, const double* x0, const double* yref, const double* p, double* x, double* u,
                         double* pi, int* status, int* sqp_iter, int* qp_iter) {
    constexpr int NBLK = M::NBLK, NX = M::NX, NU = M::NU, ny = NX + NU;
    const int N = o->N;
    Sim<M, T> sim(B, N);
    for_all<M>(B, [&](int i, int b, HostXchg<NBLK>& xc) {
        const size_t slot = (size_t)i * NBLK + b;
        for (int k = 0; k <= N; k++) field_xfer<M, T, true>(sim.w, slot, b, F_X, k, N, x + ((size_t)i * (N + 1) + k) * NX);
        for (int k = 0; k < N; k++) field_xfer<M, T, true>(sim.w, slot, b, F_U, k, N, u + ((size_t)i * N + k) * NU);
        field_xfer<M, T, true>(sim.w, slot, b, F_LBX, 0, N, const_cast<double*>(x0) + (size_t)i * NX);
        field_xfer<M, T, true>(sim.w, slot, b, F_P, 0, N, const_cast<double*>(p) + (size_t)i * 2);
        yref_all_to_ws<M, T>(sim.w, slot, b, N, yref + (size_t)i * (N * ny + NX));
        BlockSolver<M, T, HostXchg<NBLK>> bs(sim.w, *o, xc, slot, b);
        bs.sqp_solve(true, i);
        for (int k = 0; k <= N; k++) field_xfer<M, T, false>(sim.w, slot, b, F_X, k, N, x + ((size_t)i * (N + 1) + k) * NX);
        for (int k = 0; k < N; k++) field_xfer<M, T, false>(sim.w, slot, b, F_U, k, N, u + ((size_t)i * N + k) * NU);
        if (pi) for (int k = 0; k < N; k++) field_xfer<M, T, false>(sim.w, slot, b, F_PI, k, N, pi + ((size_t)i * N + k) * NX);
    });
    for (int i = 0; i < B; i++) { status[i] = sim.w.status[i]; sqp_iter[i] = sim.w.sqp_iter[i]; qp_iter[i] = sim.w.qp_iter[i]; }
    return 0;
}

template <class M, class T>
static int closed_loop_t(const Opts* o, int kind, int B, int n_steps, int rows, const double* ref, int ref_shared, const double* x0,
                         const double* noise, const double* p_ctrl, const double* p_plant, double* Xsim, double* U_plant,
                         double* U_ctrl, double* a_log, double* cost, double* abs_err, int* status, int* qp_iter) {
    constexpr int NBLK = M::NBLK;
    const int N = o->N;
    if (rows < n_steps + N) return -1;
    Sim<M, T> sim(B, N);
    const size_t Bp = B;
    std::vector<double> xs(4 * Bp), acc(2 * Bp), pp(2 * Bp);
    for (int i = 0; i < B; i++) {
        for (int j = 0; j < 4; j++) { xs[j * Bp + i] = x0[j * B + i]; if (Xsim) Xsim[(size_t)j * B + i] = x0[j * B + i]; }
        acc[i] = 0.0; acc[Bp + i] = p_ctrl[B + i];
        pp[i] = p_plant[i]; pp[Bp + i] = p_plant[B + i];
        cost[i] = 0; abs_err[i] = 0;
    }
    for_all<M>(B, [&](int i, int b, HostXchg<NBLK>& xc) {
        const size_t slot = (size_t)i * NBLK + b;
        double pc[2] = {p_ctrl[i], p_ctrl[B + i]};
        field_xfer<M, T, true>(sim.w, slot, b, F_P, 0, N, pc);
        BlockSolver<M, T, HostXchg<NBLK>> bs(sim.w, *o, xc, slot, b);
        for (int st = 0; st < n_steps; st++) {
            LoopArgs a{};
            a.step = st; a.kind = kind; a.ref_shared = ref_shared; a.log_stride = n_steps; a.batch = B; a.Bp = Bp;
            a.ref = ref; a.noise = noise; a.Xsim = Xsim; a.U_plant = U_plant; a.U_ctrl = U_ctrl; a.a_log = a_log;
            a.status = status; a.qp_iter = qp_iter; a.xs = xs.data(); a.acc = acc.data(); a.cost = cost; a.abs_err = abs_err;
            a.p_plant = pp.data();
            closed_loop_step<M, T>(bs, true, i, a);
        }
    });
    return 0;
}

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
    }                                                                                      \
    return -2;

extern "C" {
int hs_sizeof_opts(void) { return (int)sizeof(Opts); }
// AoS in/out like the C oracle's orc_solve_batch: x [B][N+1][NX], u [B][N][NU] (start iterate in, solution out)
int hs_solve_batch(int model, int prec, const Opts* o, int B, const double* x0, const double* yref, const double* p, double* x,
                   double* u, double* pi, int* status, int* sqp_iter, int* qp_iter) {
    DISPATCH(solve_batch_t, model, prec, o, B, x0, yref, p, x, u, pi, status, sqp_iter, qp_iter)
}
// batch-minor in/out exactly like bnmpc_closed_loop_* (x0 [4][B], p_* [2][B], ref shared [rows][8] or [rows][8][B], logs [steps][dim][B])
int hs_closed_loop(int model, int prec, const Opts* o, int B, int n_steps, int rows, const double* ref, int ref_shared,
                   const double* x0, const double* noise, const double* p_ctrl, const double* p_plant, double* Xsim, double* U_plant,
                   double* U_ctrl, double* a_log, double* cost, double* abs_err, int* status, int* qp_iter) {
    const int kind = (model == 1 || model == 3) ? KIND_JERK : KIND_FORCE;
    DISPATCH(closed_loop_t, model, prec, o, kind, B, n_steps, rows, ref, ref_shared, x0, noise, p_ctrl, p_plant, Xsim, U_plant, U_ctrl,
             a_log, cost, abs_err, status, qp_iter)
}
}
