// bnmpc_loop.cuh - the glue either side of solve(): acados-style field access (set/get), yref windowing, converters,
// plant step, logged quantities.  Reference (BroilerCompiler/drone-attitude-control):
//   OCP.set_up_ocp        src/force_model/ocp.py:117-122, src/jerk_model/ocp.py:118-123
//   follow_trajectory     src/force_model/controller.py:18-56, src/jerk_model/controller.py:18-58
//   Converter.convert     src/force_model/dynamics.py:66-70, src/jerk_model/dynamics.py:76-83
//   OCP.simulate_next_x   src/force_model/ocp.py:106-115, src/jerk_model/ocp.py:106-116
#pragma once
#include "bnmpc_core.cuh"

namespace bnmpc {

// field ids of the C-ABI (include/bnmpc.h)
enum { F_X = 0, F_U = 1, F_YREF = 2, F_LBX = 3, F_UBX = 4, F_P = 5, F_PI = 6, F_LAM = 7 };
enum { KIND_FORCE = 0, KIND_JERK = 1 };

// which block / local index holds global input g, state g
template <class M> BN_HD constexpr int ublk(int g) { for (int b = 0; b < M::NBLK; b++) for (int j = 0; j < M::NUB; j++) if (M::ug(b, j) == g) return b; return 0; }
template <class M> BN_HD constexpr int uloc(int g) { for (int b = 0; b < M::NBLK; b++) for (int j = 0; j < M::NUB; j++) if (M::ug(b, j) == g) return j; return 0; }
template <class M> BN_HD constexpr int xblk(int g) { for (int b = 0; b < M::NBLK; b++) for (int j = 0; j < M::NXB; j++) if (M::xg(b, j) == g) return b; return 0; }
template <class M> BN_HD constexpr int xloc(int g) { for (int b = 0; b < M::NBLK; b++) for (int j = 0; j < M::NXB; j++) if (M::xg(b, j) == g) return j; return 0; }

// AoS dimension of a field at a stage (0 = not valid there)
template <class M>
BN_HD int field_dim(int field, int stage, int N) {
    switch (field) {
    case F_X: return (stage >= 0 && stage <= N) ? M::NX : 0;
    case F_U: return (stage >= 0 && stage < N) ? M::NU : 0;
    case F_YREF: return (stage >= 0 && stage < N) ? M::NX + M::NU : (stage == N ? M::NX : 0);
    case F_LBX: case F_UBX: return stage == 0 ? M::NX : 0;
    case F_P: return M::NP;
    case F_PI: return (stage >= 0 && stage < N) ? M::NX : 0;
    case F_LAM: return (stage == 0) ? 2 * M::NU : ((stage > 0 && stage < N) ? 2 * (M::NU + M::NX) : 0);
    }
    return 0;
}

// One thread (instance, block) moves its share of an AoS vector [dim] (FP64, one instance) to / from the workspace.
template <class M, class T, bool TO_WS>
BN_HD void field_xfer(const Ws<T>& w, size_t slot, int b, int field, int k, int N, double* aos) {
    constexpr int n = M::NXB, m = M::NUB, s = n + m, NX = M::NX, NU = M::NU;
    auto at = [&](int arr, int i) -> T& { return w.base[(size_t)(w.off[arr] + i) * w.S + slot]; };
    auto mv = [&](T& ws, double& a) { if (TO_WS) ws = T(a); else a = double(ws); };
    switch (field) {
    case F_X:
#pragma unroll
        for (int j = 0; j < n; j++) mv(at(A_V, k * s + m + j), aos[M::xg(b, j)]);
        break;
    case F_U:
#pragma unroll
        for (int j = 0; j < m; j++) mv(at(A_V, k * s + j), aos[M::ug(b, j)]);
        break;
    case F_YREF:
#pragma unroll
        for (int j = 0; j < n; j++) mv(at(A_YREF, k * s + m + j), aos[M::xg(b, j)]);
        if (k < N) {
#pragma unroll
            for (int j = 0; j < m; j++) mv(at(A_YREF, k * s + j), aos[NX + M::ug(b, j)]);
        }
        break;
    case F_LBX: case F_UBX:
#pragma unroll
        for (int j = 0; j < n; j++) mv(at(A_X0, j), aos[M::xg(b, j)]);
        break;
    case F_P:
#pragma unroll
        for (int j = 0; j < M::NP; j++) { if (TO_WS) at(A_PAR, j) = T(aos[j]); else if (b == 0) aos[j] = double(at(A_PAR, j)); }
        break;
    case F_PI:
#pragma unroll
        for (int j = 0; j < n; j++) mv(at(A_PI, k * n + j), aos[M::xg(b, j)]);
        break;
    case F_LAM: {   // [lbu, lbx, ubu, ubx]; stage 0 has no x bounds
        const int half = (k == 0) ? NU : NU + NX;
#pragma unroll
        for (int j = 0; j < m; j++) { mv(at(A_LAM, k * 2 * s + j), aos[M::ug(b, j)]); mv(at(A_LAM, k * 2 * s + s + j), aos[half + M::ug(b, j)]); }
        if (k >= 1) {
#pragma unroll
            for (int j = 0; j < n; j++) { mv(at(A_LAM, k * 2 * s + m + j), aos[NU + M::xg(b, j)]); mv(at(A_LAM, k * 2 * s + s + m + j), aos[half + NU + M::xg(b, j)]); }
        }
    } break;
    }
}

// whole yref window [N*ny + ny_e] of one instance (bnmpc_set_yref_all = OCP.set_up_ocp)
template <class M, class T>
BN_HD void yref_all_to_ws(const Ws<T>& w, size_t slot, int b, int N, const double* aos) {
    constexpr int ny = M::NX + M::NU;
    for (int k = 0; k <= N; k++) field_xfer<M, T, true>(w, slot, b, F_YREF, k, N, const_cast<double*>(aos) + (size_t)k * ny);
}

// ---------------------------------------------------------------------------------------------------------------------
// fused closed loop
// ---------------------------------------------------------------------------------------------------------------------
struct LoopArgs {
    int step;            // absolute control-step index (row of ref / noise / logs)
    int kind;            // KIND_FORCE | KIND_JERK (converter + logged quantities)
    int ref_shared, log_stride, batch, pad;
    size_t Bp;           // stride of the handle-owned state arrays
    const double* ref; const double* noise;
    double *Xsim, *U_plant, *U_ctrl, *a_log; int32_t *status, *qp_iter;
    // handle-owned loop state
    double *xs, *acc, *cost, *abs_err; const double *p_plant;
};

template <class M, class T, class X>
BN_HD void closed_loop_post(BlockSolver<M, T, X>& bs, int inst, const LoopArgs& a, const double* u0, const double* xo);

template <class M, class T, class X>
BN_HD void closed_loop_step(BlockSolver<M, T, X>& bs, bool act, int inst, const LoopArgs& a) {
    constexpr int n = M::NXB, m = M::NUB, s = n + m, NX = M::NX, NU = M::NU;
    const int N = bs.N, b = bs.b;
    const size_t Bp = a.Bp, Bt = (size_t)a.batch;
    auto refv = [&](int row, int col) -> double {
        return a.ref_shared ? a.ref[(size_t)row * 8 + col] : a.ref[((size_t)row * 8 + col) * Bt + inst];
    };
    if (act) {
        // set_up_ocp: yref_k = [xref[i+k], uref[i+k]], yref_N = xref[i+N]; xref = ref[:, :NX], uref = ref[:, NX:NX+NU]
        for (int k = 0; k <= N; k++) {
#pragma unroll
            for (int j = 0; j < n; j++) bs.at(A_YREF, k * s + m + j) = T(refv(a.step + k, M::xg(b, j)));
            if (k < N) {
#pragma unroll
                for (int j = 0; j < m; j++) bs.at(A_YREF, k * s + j) = T(refv(a.step + k, NX + M::ug(b, j)));
            }
        }
        // x0_bar = Xsim[i] (+ a_i for the jerk model)
#pragma unroll
        for (int j = 0; j < n; j++) {
            const int g = M::xg(b, j);
            bs.at(A_X0, j) = T(g < 4 ? a.xs[(size_t)g * Bp + inst] : a.acc[(size_t)(g - 4) * Bp + inst]);
        }
    }
    bs.sqp_solve(act, inst);
    // collect u0 and the state entering the logged cost from the block threads
    double u0[NU], xo[4];
    const int cost_stage = (a.kind == KIND_JERK) ? 1 : 0;   // controller.py:39 get(0,'x') / jerk controller.py:39 get(1,'x')
#pragma unroll
    for (int g = 0; g < NU; g++) {
        const double v = act ? double(bs.at(A_V, uloc<M>(g))) : 0.0;
        u0[g] = bs.xc.from_block(v, ublk<M>(g));
    }
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const double v = act ? double(bs.at(A_V, cost_stage * s + m + xloc<M>(g))) : 0.0;
        xo[g] = bs.xc.from_block(v, xblk<M>(g));
    }
    // ---- leader thread of the instance: cost, converter, plant step, noise, logs ------------------------------------
    if (act && b == 0) closed_loop_post<M, T, X>(bs, inst, a, u0, xo);
    bs.xc.sync();   // the other block threads read the new plant state at the start of the next step
}

template <class M, class T, class X>
BN_HD void closed_loop_post(BlockSolver<M, T, X>& bs, int inst, const LoopArgs& a, const double* u0, const double* xo) {
    const size_t Bp = a.Bp, Bt = (size_t)a.batch;
    auto refv = [&](int row, int col) -> double {
        return a.ref_shared ? a.ref[(size_t)row * 8 + col] : a.ref[((size_t)row * 8 + col) * Bt + inst];
    };
    const double pc_mass = double(bs.par[0]);
    double xs[4], pp[2];
#pragma unroll
    for (int j = 0; j < 4; j++) xs[j] = a.xs[(size_t)j * Bp + inst];
    pp[0] = a.p_plant[inst]; pp[1] = a.p_plant[Bp + inst];
    const double wc[4] = {1e2, 1e2, 1.0, 1.0};            // controller.py:40
    double c = 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) { const double d = xo[j] - refv(a.step, j); c += wc[j] * d * d; }
    a.cost[inst] += c;
    a.abs_err[inst] += fabs(refv(a.step, 0) - xs[0]) + fabs(refv(a.step, 1) - xs[1]);   // store_results.py:233-236
    double up[2], alog[2];
    if (a.kind == KIND_JERK) {
        double ai[2] = {a.acc[inst], a.acc[Bp + inst]};
        const double hc = bs.o.sim_dt;
        for (int j = 0; j < bs.o.sim_substeps; j++) {
            ai[0] += u0[0] * hc; ai[1] += u0[1] * hc;
            const double fx = pc_mass * ai[0], fz = pc_mass * ai[1];
            up[0] = atan2(fx, fz); up[1] = sqrt(fx * fx + fz * fz);
            plant_step<double>(bs.o.sim_erk_stages, pp, hc, up, xs);
        }
        a.acc[inst] = ai[0]; a.acc[Bp + inst] = ai[1];
        alog[0] = ai[0]; alog[1] = ai[1];
    } else {
        up[0] = atan2(u0[0], u0[1]); up[1] = sqrt(u0[0] * u0[0] + u0[1] * u0[1]);
        for (int j = 0; j < bs.o.sim_substeps; j++) plant_step<double>(bs.o.sim_erk_stages, pp, bs.o.sim_dt, up, xs);
        alog[0] = u0[0] / 0.03277; alog[1] = u0[1] / 0.03277;   // controller.py:38 (drone.MASS, params.py:42)
    }
    const double eps = a.noise ? a.noise[(size_t)a.step * Bt + inst] : 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) { xs[j] += eps; a.xs[(size_t)j * Bp + inst] = xs[j]; }
    const size_t l2 = (size_t)a.step * 2 * Bt + inst;
    if (a.U_ctrl) { a.U_ctrl[l2] = u0[0]; a.U_ctrl[l2 + Bt] = u0[1]; }
    if (a.U_plant) { a.U_plant[l2] = up[0]; a.U_plant[l2 + Bt] = up[1]; }
    if (a.a_log) { a.a_log[l2] = alog[0]; a.a_log[l2 + Bt] = alog[1]; }
    if (a.Xsim) {
#pragma unroll
        for (int j = 0; j < 4; j++) a.Xsim[((size_t)(a.step + 1) * 4 + j) * Bt + inst] = xs[j];
    }
    if (a.status) a.status[(size_t)a.step * Bt + inst] = bs.w.status[inst];
    if (a.qp_iter) a.qp_iter[(size_t)a.step * Bt + inst] = bs.w.qp_iter[inst];
}

}  // namespace bnmpc
