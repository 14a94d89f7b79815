// bnmpc_kernels.cuh - __global__ wrappers around the solver templates and the per-model launch table.
// Each model_*.cu instantiates BNMPC_DEFINE_MODEL_OPS for one generated model (both precisions); bnmpc_api.cu only
// sees the ModelOps table, so the heavy templates compile in parallel translation units.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include "bnmpc_loop.cuh"

namespace bnmpc {

// type-erased Ws<T> (identical layout for every T)
struct WsAny {
    void* base; size_t S; int B; int off[A_COUNT];
    int32_t *status, *sqp_iter, *qp_iter, *have_mult;
};
template <class T> inline Ws<T> ws_cast(const WsAny& a) {
    static_assert(sizeof(Ws<T>) == sizeof(WsAny), "layout");
    Ws<T> w; memcpy(&w, &a, sizeof(w)); return w;
}

struct ModelOps {
    const char* name;
    int nx, nu, np, nblk, nxb, nub, kind, jac_const, elem_size;
    int (*layout)(int N, int* off);                                                     // rows of the workspace
    int (*field_dim)(int field, int stage, int N);
    cudaError_t (*solve)(const WsAny&, const Opts&, int tpb, cudaStream_t);
    cudaError_t (*loop_step)(const WsAny&, const Opts&, const LoopArgs&, int tpb, cudaStream_t);
    // AoS [B][dim] (stride `aos_stride` doubles between instances; 0 = one vector for all) <-> workspace
    cudaError_t (*field)(const WsAny&, int field, int stage, int N, double* aos, int aos_stride, int to_ws, cudaStream_t);
    cudaError_t (*yref_all)(const WsAny&, int N, const double* aos, cudaStream_t);
    // p_ctrl [2][B] batch-minor -> workspace parameters
    cudaError_t (*par_from_bm)(const WsAny&, const double* p_ctrl, cudaStream_t);
};

template <class M, class T>
__global__ void k_solve(const __grid_constant__ Ws<T> w, const __grid_constant__ Opts o) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int inst = (int)(slot / M::NBLK), b = (int)(slot % M::NBLK);
    const WarpXchg<M::NBLK> xc;
    BlockSolver<M, T, WarpXchg<M::NBLK>> bs(w, o, xc, slot, b);
    bs.sqp_solve(inst < w.B, inst);
}

template <class M, class T>
__global__ void k_loop_step(const __grid_constant__ Ws<T> w, const __grid_constant__ Opts o, const __grid_constant__ LoopArgs a) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int inst = (int)(slot / M::NBLK), b = (int)(slot % M::NBLK);
    const WarpXchg<M::NBLK> xc;
    BlockSolver<M, T, WarpXchg<M::NBLK>> bs(w, o, xc, slot, b);
    closed_loop_step<M, T>(bs, inst < w.B, inst, a);
}

template <class M, class T, bool TO_WS>
__global__ void k_field(const __grid_constant__ Ws<T> w, int field, int stage, int N, double* aos, int aos_stride) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int inst = (int)(slot / M::NBLK), b = (int)(slot % M::NBLK);
    if (inst >= w.B) return;
    field_xfer<M, T, TO_WS>(w, slot, b, field, stage, N, aos + (size_t)inst * aos_stride);
}

template <class M, class T>
__global__ void k_yref_all(const __grid_constant__ Ws<T> w, int N, const double* aos) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int inst = (int)(slot / M::NBLK), b = (int)(slot % M::NBLK);
    if (inst >= w.B) return;
    yref_all_to_ws<M, T>(w, slot, b, N, aos + (size_t)inst * (N * (M::NX + M::NU) + M::NX));
}

template <class M, class T>
__global__ void k_par_from_bm(const __grid_constant__ Ws<T> w, const double* p_ctrl) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int inst = (int)(slot / M::NBLK);
    if (inst >= w.B) return;
#pragma unroll
    for (int j = 0; j < M::NP; j++) w.base[(size_t)(w.off[A_PAR] + j) * w.S + slot] = T(p_ctrl[(size_t)j * w.B + inst]);
}

template <class M, class T>
struct OpsImpl {
    static int grid(const WsAny& a, int tpb) { return (int)((a.S + tpb - 1) / tpb); }
    static cudaError_t solve(const WsAny& a, const Opts& o, int tpb, cudaStream_t st) {
        k_solve<M, T><<<grid(a, tpb), tpb, 0, st>>>(ws_cast<T>(a), o);
        return cudaGetLastError();
    }
    static cudaError_t loop_step(const WsAny& a, const Opts& o, const LoopArgs& la, int tpb, cudaStream_t st) {
        k_loop_step<M, T><<<grid(a, tpb), tpb, 0, st>>>(ws_cast<T>(a), o, la);
        return cudaGetLastError();
    }
    static cudaError_t field(const WsAny& a, int field, int stage, int N, double* aos, int stride, int to_ws, cudaStream_t st) {
        if (to_ws) k_field<M, T, true><<<grid(a, 128), 128, 0, st>>>(ws_cast<T>(a), field, stage, N, aos, stride);
        else k_field<M, T, false><<<grid(a, 128), 128, 0, st>>>(ws_cast<T>(a), field, stage, N, aos, stride);
        return cudaGetLastError();
    }
    static cudaError_t yref_all(const WsAny& a, int N, const double* aos, cudaStream_t st) {
        k_yref_all<M, T><<<grid(a, 128), 128, 0, st>>>(ws_cast<T>(a), N, aos);
        return cudaGetLastError();
    }
    static cudaError_t par_from_bm(const WsAny& a, const double* p, cudaStream_t st) {
        k_par_from_bm<M, T><<<grid(a, 128), 128, 0, st>>>(ws_cast<T>(a), p);
        return cudaGetLastError();
    }
    static int fdim(int field, int stage, int N) { return field_dim<M>(field, stage, N); }
    static ModelOps make(int kind) {
        return ModelOps{M::name(), M::NX, M::NU, M::NP, M::NBLK, M::NXB, M::NUB, kind, M::JAC_CONST ? 1 : 0, (int)sizeof(T),
                        &WsLayout<M>::fill, &fdim, &solve, &loop_step, &field, &yref_all, &par_from_bm};
    }
};

#define BNMPC_DEFINE_MODEL_OPS(MODEL, KIND, FN)                                               \
    namespace bnmpc {                                                                         \
    const ModelOps* FN(int precision) {                                                       \
        static const ModelOps d = OpsImpl<MODEL, double>::make(KIND);                          \
        static const ModelOps f = OpsImpl<MODEL, float>::make(KIND);                           \
        return precision == 0 ? &d : &f;                                                      \
    }                                                                                         \
    }

const ModelOps* ops_force(int precision);
const ModelOps* ops_jerk(int precision);
const ModelOps* ops_force_dense(int precision);
const ModelOps* ops_jerk_dense(int precision);

}  // namespace bnmpc
