// bnmpc_lockstep.cuh - the slotted lockstep schedule of the fused closed loop.
//
// Why.  With one warp per OCP instance (bnmpc_core.cuh) the Riccati factorisation and the two scans of every Newton solve
// are sequential in the stage index and run on NBLK lanes of the instance's warp: a third of all issued warp instructions
// execute with 2 of 32 lanes active (profiles/r01_v7_source_regions.txt).  Here the warps of the one CTA per SM walk the
// interior-point loop SIDE BY SIDE - every warp still owns one instance and runs its 32-lane passes on its own working
// set - and the sequential sweeps of ALL instances of the CTA are executed together by one warp, NBLK lanes per instance
// (force model: 16 instances x 2 blocks = 32 lanes).  A sweep costs the CTA the same dependent chain as before, but it is
// issued once instead of sixteen times.
//
// Schedule.  Time is cut into half-rounds of four pass slots P1..P4 separated by CTA barriers, with a sweep slot after
// P1 (factorisation), P2 (backward scan) and P3 (forward scan):
//
//      P1 | S:factor | P2 | S:back | P3 | S:fwd | P4                 (one Newton solve = one half-round)
//
//   instance inside the interior-point loop (mode = predictor / corrector / centering fallback):
//      P1 residual pass (+ convergence test in predictor mode)   P2 stage-local part of the backward solve (predictor)
//      P3 feed-forward                                            P4 step length, mu_aff [, variable update]
//   instance between two QPs (the pieces of Solver::sqp_solve and closed_loop_step, one per slot):
//      P1 QP exit + full step | begin (queue ticket, state load)   P2 [input check] linearise
//      P3 NLP residual test                                        P4 finish (results, plant step, logs) | build QP + IPM init
// A solve that ends in P1 of a predictor half-round is back in the loop with its next problem exactly two half-rounds
// (= one interior-point iteration) later, in phase with the other instances, so factorisations keep falling into the
// same sweep slot.  Nothing of this changes the arithmetic of an instance: results are bit-identical to the one-warp-
// per-instance kernels (tests/test_gpu_parity.py::test_results_do_not_depend_on_the_launch_shape).
//
// Work items.  A queue ticket is (instance, chunk of control steps): an instance keeps its working set on chip for
// `chunk` consecutive control steps (only the first one loads the persistent state from HBM) and is then handed back so
// that any warp of any SM can run its next chunk; next_step[instance] orders the chunks of an instance.  Tickets are
// issued instance-round-robin, so the predecessor of a ticket was taken B tickets earlier and is finished or in flight
// on a resident warp - a warp whose predecessor is not done yet polls once per half-round without holding up its CTA.
//
// ls_pslot() is __host__ __device__: tests/hostsim emulates a CTA (warps one after the other, barriers implicit) so the
// state machine is checked against the oracle in the CPU-only test run.
#pragma once
#include "bnmpc_loop.cuh"

namespace bnmpc {

enum { PC_IDLE = 0, PC_BEGIN, PC_LIN, PC_NLPRES, PC_FINISH, PC_BUILD, PC_IPM };
enum { LS_ALIVE = 8 };   // bit of the published warp state; bits 0..2: 0 = not in the IPM loop, 1 + mode otherwise

template <class M, class T, class G, class PS>
struct LsWarp {
    using SV = Solver<M, T, G, PS>;
    int pc;                 // what this warp's instance does next (PC_*)
    int inst;               // instance, -1 = none
    int step, step_end;     // current control step, end of the chunk
    int ticket;             // queue ticket taken but not started yet (its predecessor chunk is still running), -1 = none
    bool have_mult, check_inputs;
    int status, sqp_it, qp_it;
    typename SV::IpmState q;
    YrefSrc ys;
    BN_HD void reset() { pc = PC_BEGIN; inst = -1; step = 0; step_end = 0; ticket = -1; have_mult = false; check_inputs = false;
                         status = 0; sqp_it = 0; qp_it = 0; q.mode = 0; q.it = 0; }
    BN_HD int published() const { return (pc != PC_IDLE ? LS_ALIVE : 0) | (pc == PC_IPM ? 1 + q.mode : 0); }
};

// Pass slot `slot` (1..4) of the half-round for the instance of warp state `w`.  `wq` hands out tickets:
//   int  take()                      next ticket (uniform over the group), >= total() when the queue is empty
//   bool ready(inst, step)           has the previous chunk of `inst` been published?
//   void publish(inst, next_step)    all writes of this group to the state of `inst` become visible, then next_step[inst]
template <class M, class T, class G, class PS, class WQ>
BN_HD void ls_pslot(const int slot, LsWarp<M, T, G, PS>& w, Solver<M, T, G, PS>& sv, const Gs<T>& gs, const LoopArgs& a, WQ& wq,
                    const bool others_in_ipm = false) {
    using SV = Solver<M, T, G, PS>;
    const Opts& o = sv.o;
    // generation mode: an instance that has left the interior-point loop waits until the others of the CTA have too, so that
    // the pieces between two QPs (global-memory latency, scalar plant step) never stretch a slot of the interior-point loop
    if (a.ls_generation && others_in_ipm && w.pc != PC_IPM) return;
    if (slot == 1) {
        if (w.pc == PC_IPM) {
            T nr[4], ms;
            sv.residual_pass(w.q.mode, w.q.sigma_mu, nr, ms);
            sv.g.sync();
            if (w.q.mode == 0 && !sv.ipm_check(w.q, nr, ms)) {          // the QP is done: acados' SQP step
                const int qs = sv.ipm_status(w.q);
                w.qp_it += w.q.it; w.sqp_it += 1;
                if (qs != 0 && qs != 1) { w.status = ST_QP_FAILURE; w.pc = PC_FINISH; }
                else {
                    sv.full_step();
                    w.have_mult = true;
                    if (o.rti) { w.status = (qs == 0) ? ST_SUCCESS : ST_MAXITER; w.pc = PC_FINISH; }
                    else w.pc = PC_LIN;
                }
            }
        } else if (w.pc == PC_BEGIN) {
            const int B = gs.B, nchunk = (a.n_steps + a.chunk - 1) / a.chunk;
            if (w.ticket < 0) w.ticket = wq.take();
            if (w.ticket >= B * nchunk) { w.pc = PC_IDLE; return; }
            const int c = w.ticket / B, pos = w.ticket - c * B;
            const int inst = o.order ? o.order[pos] : pos;
            const int s0 = a.step + c * a.chunk;
            if (c > 0 && !wq.ready(inst, s0)) return;                    // its previous chunk is still running: ask again next half-round
            w.ticket = -1;
            w.inst = inst; w.step = s0; w.step_end = s0 + a.chunk < a.step + a.n_steps ? s0 + a.chunk : a.step + a.n_steps;
            w.ys = loop_yref(inst, s0, sv.N, a);
            loop_begin<M, T>(sv, inst, gs, a);
            w.have_mult = ld_cg(gs.have_mult + inst) != 0;
            sv.load_state(gs, inst, w.have_mult);
            w.status = ST_SUCCESS; w.sqp_it = 0; w.qp_it = 0; w.check_inputs = true;
            w.pc = PC_LIN;
        }
    } else if (slot == 2) {
        if (w.pc == PC_IPM) {
            if (w.q.mode == 0) sv.solve_pre();
        } else if (w.pc == PC_LIN) {
            if (w.check_inputs) {
                w.check_inputs = false;
                if (sv.g.any(!sv.template inputs_finite<T>(w.ys))) { w.status = ST_FAILURE; w.pc = PC_FINISH; return; }
            }
            sv.linearise();
            w.pc = o.rti ? PC_BUILD : PC_NLPRES;
        }
    } else if (slot == 3) {
        if (w.pc == PC_IPM) sv.solve_mid(w.q.mode);
        else if (w.pc == PC_NLPRES) {
            T res[4];
            sv.template nlp_residuals<T>(w.ys, w.have_mult, res);
            if (sv.nlp_converged(res)) { w.status = ST_SUCCESS; w.pc = PC_FINISH; }
            else if (w.sqp_it >= o.sqp_max_iter) { w.status = ST_MAXITER; w.pc = PC_FINISH; }
            else w.pc = PC_BUILD;
        }
    } else {
        if (w.pc == PC_IPM) {
            typename SV::StepInfo si;
            sv.step_pass(w.q.mode, w.q.sigma_mu, si);
            if (sv.ipm_step_decision(w.q, si)) {
                sv.qp_update(w.q.mode, w.q.sigma_mu, w.q.alpha);
                w.q.it++; w.q.mode = 0;
            }
        } else if (w.pc == PC_BUILD) {
            sv.template build_qp<T>(w.ys);
            sv.g.sync();
            sv.qp_init();
            sv.ipm_start(w.q);
            w.pc = PC_IPM;
        } else if (w.pc == PC_FINISH) {
            const int inst = w.inst;
            sv.store_result(gs, inst, w.status, w.sqp_it, w.qp_it, w.have_mult);
            const bool more = w.step + 1 < w.step_end;
            loop_finish<M, T>(sv, inst, w.step, w.status, w.qp_it, w.ys, a, more);
            if (more) {                                                  // the instance stays: next control step, state on chip
                // a failed QP leaves its own multipliers on chip; the next solve starts from those of the last good solve,
                // which is what the persistent state holds (store_state above did not replace them)
                if (w.status == ST_QP_FAILURE) { sv.g.sync(); sv.load_state(gs, inst, w.have_mult); }
                w.step++;
                w.ys.row0 = w.step;
                w.status = ST_SUCCESS; w.sqp_it = 0; w.qp_it = 0; w.check_inputs = true;
                w.pc = PC_LIN;
            } else {
                wq.publish(inst, w.step_end);
                w.inst = -1;
                w.pc = PC_BEGIN;
            }
        }
    }
}

}  // namespace bnmpc
