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
enum { KIND_FORCE = 0, KIND_JERK = 1, KIND_THRUST = 2 };
enum { REF_BATCH_MINOR = 0, REF_SHARED = 1, REF_INSTANCE_MAJOR = 2, REF_CIRCLE = 3 };   // bnmpc_closed_loop_args.ref_shared

// which block / local index holds global input g, state g
template <class M> BN_HD constexpr int ublk(int g) { for (int b = 0; b < M::NBLK; b++) for (int j = 0; j < M::NUB; j++) if (M::ug(b, j) == g) return b; return 0; }
template <class M> BN_HD constexpr int uloc(int g) { for (int b = 0; b < M::NBLK; b++) for (int j = 0; j < M::NUB; j++) if (M::ug(b, j) == g) return j; return 0; }
template <class M> BN_HD constexpr int xblk(int g) { for (int b = 0; b < M::NBLK; b++) for (int j = 0; j < M::NXB; j++) if (M::xg(b, j) == g) return b; return 0; }
template <class M> BN_HD constexpr int xloc(int g) { for (int b = 0; b < M::NBLK; b++) for (int j = 0; j < M::NXB; j++) if (M::xg(b, j) == g) return j; return 0; }

// AoS dimension of a field at a stage (0 = not valid there)
BN_HD int field_dim(int NX, int NU, int NP, int field, int stage, int N) {
    switch (field) {
    case F_X: return (stage >= 0 && stage <= N) ? NX : 0;
    case F_U: return (stage >= 0 && stage < N) ? NU : 0;
    case F_YREF: return (stage >= 0 && stage < N) ? NX + NU : (stage == N ? NX : 0);
    case F_LBX: case F_UBX: return stage == 0 ? NX : 0;
    case F_P: return NP;
    case F_PI: return (stage >= 0 && stage < N) ? NX : 0;
    case F_LAM: return (stage == 0) ? 2 * NU : ((stage > 0 && stage < N) ? 2 * (NU + NX) : 0);
    }
    return 0;
}

// element j of the AoS vector of (field, stage) <-> its place in the persistent state of instance `inst`
template <class T>
BN_HD T* field_ptr(const Gs<T>& gs, int NX, int NU, int NP, int inst, int field, int k, int j) {
    const int N = gs.N, SG = NU + NX;
    switch (field) {
    case F_X: return gs.V + (size_t)inst * (N + 1) * SG + k * SG + NU + j;
    case F_U: return gs.V + (size_t)inst * (N + 1) * SG + k * SG + j;
    case F_YREF:   // acados order [x-part; u-part] (Vx = [I; 0], Vu = [0; I]), stored as is
        return gs.YREF + (size_t)inst * (N * SG + NX) + k * SG + j;
    case F_LBX: case F_UBX: return gs.X0 + (size_t)inst * NX + j;
    case F_P: return gs.PAR + (size_t)inst * NP + j;
    case F_PI: return gs.PI + (size_t)inst * N * NX + k * NX + j;
    case F_LAM: {  // acados order [lbu, lbx, ubu, ubx]; stage 0 has no x bounds
        const int half = (k == 0) ? NU : SG;
        const int side = j >= half, v = j - side * half;
        return gs.LAM + (size_t)inst * N * 2 * SG + k * 2 * SG + side * SG + v;
    }
    }
    return nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// fused closed loop
// ---------------------------------------------------------------------------------------------------------------------
struct LoopArgs {
    int step;            // absolute control-step index (row of ref / noise / logs)
    int kind;            // KIND_FORCE | KIND_JERK (converter + logged quantities)
    int ref_layout, log_stride, batch, ref_rows;   // ref_layout: REF_BATCH_MINOR | REF_SHARED | REF_INSTANCE_MAJOR
    size_t Bp;           // stride of the handle-owned state arrays
    const double* ref; const double* noise;
    double *Xsim, *U_plant, *U_ctrl, *a_log; int32_t *status, *qp_iter;
    // handle-owned loop state
    double *xs, *acc, *cost, *abs_err; const double *p_plant;
};

// one instance, one control step, executed by the lanes of the solver's group
template <class M, class T, class G, class PS>
BN_HD void closed_loop_step(Solver<M, T, G, PS>& sv, int inst, const Gs<T>& gs, const LoopArgs& a) {
    using SL = typename Solver<M, T, G, PS>::SL;
    constexpr int n = M::NXB, m = M::NUB, NX = M::NX, NU = M::NU, NBLK = M::NBLK, NP = M::NP;
    const size_t Bp = a.Bp, Bt = (size_t)a.batch;
    YrefSrc ys;
    ys.yref = nullptr; ys.ref = a.ref; ys.row0 = a.step; ys.circle_n = 0;
    if (a.ref_layout == REF_CIRCLE) { ys.ref_rs = 0; ys.ref_cs = 0; ys.ref_off = (size_t)inst * 4; ys.circle_n = a.ref_rows - sv.N; }
    else if (a.ref_layout == REF_SHARED) { ys.ref_rs = 8; ys.ref_cs = 1; ys.ref_off = 0; }
    else if (a.ref_layout == REF_INSTANCE_MAJOR) { ys.ref_rs = 8; ys.ref_cs = 1; ys.ref_off = (size_t)inst * a.ref_rows * 8; }
    else { ys.ref_rs = 8 * Bt; ys.ref_cs = Bt; ys.ref_off = (size_t)inst; }
    {
        T p[NP];
#pragma unroll
        for (int j = 0; j < NP; j++) p[j] = gs.PAR[(size_t)inst * NP + j];
        sv.set_par(p);
        // x0_bar = Xsim[i] (+ a_i for the jerk model): controller.py:29-31 / jerk controller.py:30-32
        for (int gi = sv.g.lane; gi < NX; gi += G::L)
            sv.X0S(gi) = T(gi < 4 ? a.xs[(size_t)gi * Bp + inst] : a.acc[(size_t)(gi - 4) * Bp + inst]);
    }
    sv.g.sync();
    sv.template sqp_solve<T>(inst, gs, ys);
    // ---- leader lane of the instance: cost, converter, plant step, noise, logs ---------------------------------------
    if (sv.g.lane == 0) {
        auto refv = [&](int row, int col) -> double {
            if (ys.circle_n > 0) return circle_ref(a.ref + ys.ref_off, row, col, ys.circle_n);
            return a.ref[(size_t)row * ys.ref_rs + (size_t)col * ys.ref_cs + ys.ref_off];
        };
        const int cost_stage = (a.kind == KIND_JERK) ? 1 : 0;   // controller.py:39 get(0,'x') / jerk controller.py:39 get(1,'x')
        double u0[NU], xo[4];
#pragma unroll
        for (int gi = 0; gi < NU; gi++) u0[gi] = double(sv.S(SL::VAL + uloc<M>(gi), ublk<M>(gi)));
#pragma unroll
        for (int gi = 0; gi < 4; gi++) xo[gi] = double(sv.S(SL::VAL + m + xloc<M>(gi), cost_stage * NBLK + xblk<M>(gi)));
        const double pc_mass = double(sv.par[0]);
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
            const double hc = sv.o.sim_dt;
            for (int j = 0; j < sv.o.sim_substeps; j++) {     // jerk dynamics.py:76-83 + jerk ocp.py:106-113
                ai[0] += u0[0] * hc; ai[1] += u0[1] * hc;
                const double fx = pc_mass * ai[0], fz = pc_mass * ai[1];
                up[0] = atan2(fx, fz); up[1] = sqrt(fx * fx + fz * fz);
                plant_step<double>(sv.o.sim_erk_stages, pp, hc, up, xs);
            }
            a.acc[inst] = ai[0]; a.acc[Bp + inst] = ai[1];
            alog[0] = ai[0]; alog[1] = ai[1];
        } else if (a.kind == KIND_THRUST) {               // the OCP input already is the plant input (theta, Fd)
            up[0] = u0[0]; up[1] = u0[1];
            for (int j = 0; j < sv.o.sim_substeps; j++) plant_step<double>(sv.o.sim_erk_stages, pp, sv.o.sim_dt, up, xs);
            alog[0] = u0[1] * sin(u0[0]) / 0.03277; alog[1] = u0[1] * cos(u0[0]) / 0.03277;
        } else {
            up[0] = atan2(u0[0], u0[1]); up[1] = sqrt(u0[0] * u0[0] + u0[1] * u0[1]);     // dynamics.py:66-70
            for (int j = 0; j < sv.o.sim_substeps; j++) plant_step<double>(sv.o.sim_erk_stages, pp, sv.o.sim_dt, up, xs);
            alog[0] = u0[0] / 0.03277; alog[1] = u0[1] / 0.03277;   // controller.py:38 (drone.MASS, params.py:42)
        }
        const double eps = a.noise ? a.noise[(size_t)a.step * Bt + inst] : 0.0;           // ocp.py:114-115
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
        if (a.status) a.status[(size_t)a.step * Bt + inst] = gs.status[inst];
        if (a.qp_iter) a.qp_iter[(size_t)a.step * Bt + inst] = gs.qp_iter[inst];
    }
    sv.g.sync();
}

// one instance, one ocp_solver.solve() with the x0 / yref / p stored through the API
template <class M, class T, class G, class PS>
BN_HD void api_solve(Solver<M, T, G, PS>& sv, int inst, const Gs<T>& gs) {
    constexpr int NX = M::NX, NU = M::NU, NP = M::NP;
    YrefSrc ys;
    ys.ref = nullptr; ys.ref_rs = 0; ys.ref_cs = 0; ys.ref_off = 0; ys.row0 = 0; ys.circle_n = 0;
    {
        ys.yref = gs.YREF + (size_t)inst * (gs.N * (NU + NX) + NX);
        T p[NP];
#pragma unroll
        for (int j = 0; j < NP; j++) p[j] = gs.PAR[(size_t)inst * NP + j];
        sv.set_par(p);
        for (int gi = sv.g.lane; gi < NX; gi += G::L) sv.X0S(gi) = gs.X0[(size_t)inst * NX + gi];
    }
    sv.g.sync();
    sv.template sqp_solve<T>(inst, gs, ys);
    if (sv.g.lane == 0) {      // ocp_solver.get(0, 'u') for free: the solution is still in shared memory
        using SL = typename Solver<M, T, G, PS>::SL;
#pragma unroll
        for (int gi = 0; gi < NU; gi++) gs.U0[(size_t)inst * NU + gi] = double(sv.S(SL::VAL + uloc<M>(gi), ublk<M>(gi)));
    }
    sv.g.sync();
}

}  // namespace bnmpc
