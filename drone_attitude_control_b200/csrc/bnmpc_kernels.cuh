// bnmpc_kernels.cuh - __global__ wrappers around the solver templates and the per-model launch table.
// Each model_*.cu instantiates BNMPC_DEFINE_MODEL_OPS for one generated model (both precisions); bnmpc_api.cu only
// sees the ModelOps table, so the heavy templates compile in parallel translation units.
//
// Launch shape: one warp per OCP instance, `wpc` warps per CTA, each warp with its own slice of dynamic shared memory
// (SmLayout<M>::elems(N) elements).  There is no inter-warp synchronisation; the CTA is only a packing unit.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include "bnmpc_loop.cuh"

// resident one-warp CTAs per SM the register allocator should leave room for (shared memory allows 10 force / 7 jerk)
#ifndef BNMPC_MIN_BLOCKS
#define BNMPC_MIN_BLOCKS 8
#endif

namespace bnmpc {

// type-erased Gs<T> (identical layout for every T)
struct GsAny {
    void *V, *PI, *LAM, *YREF, *X0, *PAR;
    int32_t *status, *sqp_iter, *qp_iter, *have_mult;
    int B, N;
};
template <class T> inline Gs<T> gs_cast(const GsAny& a) {
    static_assert(sizeof(Gs<T>) == sizeof(GsAny), "layout");
    Gs<T> g; memcpy(&g, &a, sizeof(g)); return g;
}

struct ModelOps {
    const char* name;
    int nx, nu, np, nblk, nxb, nub, kind, jac_const, elem_size, sm_rows;
    size_t (*smem_bytes)(int N);                                    // dynamic shared memory of one instance
    cudaError_t (*solve)(const GsAny&, const Opts&, int wpc, cudaStream_t);
    cudaError_t (*loop_step)(const GsAny&, const Opts&, const LoopArgs&, int wpc, cudaStream_t);
};

template <class M, class T>
__device__ __forceinline__ T* warp_smem(int N) {
    extern __shared__ double4 smem_raw[];
    const size_t per = (SmLayout<M>::elems(N) * sizeof(T) + 15) / 16 * 16;
    return reinterpret_cast<T*>(reinterpret_cast<char*>(smem_raw) + per * (threadIdx.x >> 5));
}

template <class M, class T>
__global__ void __launch_bounds__(32, BNMPC_MIN_BLOCKS) k_solve(const __grid_constant__ Gs<T> gs, const __grid_constant__ Opts o) {
    const int inst = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (inst >= gs.B) return;          // whole warp
    const WarpGroup<32> g;
    Solver<M, T, WarpGroup<32>> sv(warp_smem<M, T>(o.N), o, g);
    api_solve<M, T>(sv, inst, gs);
}

template <class M, class T>
__global__ void __launch_bounds__(32, BNMPC_MIN_BLOCKS) k_loop_step(const __grid_constant__ Gs<T> gs, const __grid_constant__ Opts o, const __grid_constant__ LoopArgs a) {
    const int inst = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (inst >= gs.B) return;          // whole warp
    const WarpGroup<32> g;
    Solver<M, T, WarpGroup<32>> sv(warp_smem<M, T>(o.N), o, g);
    closed_loop_step<M, T>(sv, inst, gs, a);
}

template <class M, class T>
struct OpsImpl {
    static size_t smem_bytes(int N) { return (SmLayout<M>::elems(N) * sizeof(T) + 15) / 16 * 16; }
    template <class K>
    static cudaError_t prep(K kern, size_t bytes) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    static cudaError_t solve(const GsAny& a, const Opts& o, int wpc, cudaStream_t st) {
        const size_t bytes = smem_bytes(o.N) * wpc;
        cudaError_t e = prep(k_solve<M, T>, bytes);
        if (e != cudaSuccess) return e;
        k_solve<M, T><<<(a.B + wpc - 1) / wpc, 32 * wpc, bytes, st>>>(gs_cast<T>(a), o);
        return cudaGetLastError();
    }
    static cudaError_t loop_step(const GsAny& a, const Opts& o, const LoopArgs& la, int wpc, cudaStream_t st) {
        const size_t bytes = smem_bytes(o.N) * wpc;
        static size_t prepared = 0;
        if (prepared != bytes) {
            cudaError_t e = prep(k_loop_step<M, T>, bytes);
            if (e != cudaSuccess) return e;
            prepared = bytes;
        }
        k_loop_step<M, T><<<(a.B + wpc - 1) / wpc, 32 * wpc, bytes, st>>>(gs_cast<T>(a), o, la);
        return cudaGetLastError();
    }
    static ModelOps make(int kind) {
        return ModelOps{M::name(), M::NX, M::NU, M::NP, M::NBLK, M::NXB, M::NUB, kind, M::JAC_CONST ? 1 : 0, (int)sizeof(T),
                        SmLayout<M>::ROWS, &smem_bytes, &solve, &loop_step};
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
