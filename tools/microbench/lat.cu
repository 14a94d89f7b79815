// Dependent-issue latencies of the instructions on the solver's critical path (one warp, one SM), in SM cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void k(double* out, long long* cyc, double a, double b, int lanes) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) & 1023);   // pointer-chase table (as doubles)
    __syncthreads();
    double x = a + threadIdx.x; int idx = threadIdx.x;
    if ((int)threadIdx.x >= lanes) { out[threadIdx.x] = 0; return; }
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = fma(x, b, a);
        if (OP == 1) x = x + b;
        if (OP == 2) x = x * b;
        if (OP == 3) x = a / x + b;                  // full-precision division on the chain
        if (OP == 4) x = rsqrt(x) + a;
        if (OP == 5) { idx = (int)sm[idx & 1023]; }   // LDS + F2I on the chain
        if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1) + b;
        if (OP == 7) x = sqrt(x) + a;
        if (OP == 8) x = (x > b) ? x - b : x + a;    // compare + select + add
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + idx; if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int lanes) {
    double* o; long long* c; cudaMalloc(&o, 8 * 32); cudaMalloc(&c, 8);
    k<OP><<<1, 32>>>(o, c, 1.0000001, 0.9999999, lanes); k<OP><<<1, 32>>>(o, c, 1.0000001, 0.9999999, lanes);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-28s lanes %2d : %.1f cycles per dependent op\n", name, lanes, (double)h / N);
    cudaFree(o); cudaFree(c);
}
int main() {
    for (int lanes : {32, 2}) {
        run<0>("DFMA", lanes); run<1>("DADD", lanes); run<2>("DMUL", lanes); run<3>("a/x + b (div)", lanes);
        run<4>("rsqrt + add", lanes); run<5>("LDS -> F2I chase", lanes); run<6>("SHFL + DADD", lanes); run<7>("sqrt + add", lanes);
        run<8>("DSETP+select+DADD", lanes);
    }
    return 0;
}
