// Micro-benchmarks that size the splat design (atomics / warp-reduce / conversion throughput on B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int ITERS = 4096;

// shared-memory atomics: `distinct` = number of distinct addresses hit by one warp instruction
__global__ void k_atoms(int distinct, int spread, int* out) {
    __shared__ int sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int idx = warp * 256 + (lane % distinct) * spread;
    int v = lane + 1;
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) { atomicAdd(&sm[(idx + (i & 7) * 32) & 4095], v); }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}
__global__ void k_redux(int* out) {
    int v = threadIdx.x, acc = 0;
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) { acc += __reduce_add_sync(0xffffffffu, (v ^ acc) * (i | 1)); }
    if (acc == 12345) out[0] = acc;
}
__global__ void k_shfl(int* out) {
    int v = threadIdx.x, acc = 0;
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) { acc += __shfl_xor_sync(0xffffffffu, (v ^ acc) * (i | 1), 16); }
    if (acc == 12345) out[0] = acc;
}
__global__ void k_f2i(float* in, int* out) {
    float v = in[threadIdx.x]; int acc = 0;
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) { acc += __float2int_rz(v * (float)(i ^ acc)); }
    if (acc == 12345) out[0] = acc;
}
__global__ void k_match(int* out) {
    int v = threadIdx.x >> 3, acc = 0;
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) { acc += __match_any_sync(0xffffffffu, (v ^ (acc & 1)) + (i & 3)); }
    if (acc == 12345) out[0] = acc;
}
// global 64-bit reductions: each warp instruction hits `distinct` addresses out of a pool of `cells` cells,
// pool shared by all CTAs (contention across SMs like the voxel grid).
__global__ void k_redg(unsigned long long* grid, int cells, int distinct, int iters) {
    const int lane = threadIdx.x & 31;
    unsigned h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int i = 0; i < iters; ++i) {
        h = h * 1664525u + 1013904223u;
        unsigned cell = ((h >> 8) % cells) + (lane % distinct);
        atomicAdd(&grid[(size_t)(cell % cells) * 4 + (lane & 3)], (unsigned long long)(lane + 1));
    }
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    int* out; float* in; CK(cudaMalloc(&out, 1 << 20)); CK(cudaMalloc(&in, 4096)); CK(cudaMemset(in, 0, 4096));
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount; const double ghz = 1.965;
    const int blocks = sms * 2, threads = 512;   // 32 warps per SM
    const double winstr_per_sm = (double)ITERS * (threads / 32) * 2;
    for (int distinct : {32, 16, 8, 4, 2, 1}) {
        for (int spread : {1, 33}) {
            float ms = timeit([&] { k_atoms<<<blocks, threads>>>(distinct, spread, out); });
            printf("ATOMS.ADD distinct=%2d spread=%2d : %.2f cycles per warp-instr per SM\n", distinct, spread, ms * 1e-3 * ghz * 1e9 / winstr_per_sm);
        }
    }
    { float ms = timeit([&] { k_redux<<<blocks, threads>>>(out); }); printf("REDUX.SUM : %.2f cycles per warp-instr per SM\n", ms * 1e-3 * ghz * 1e9 / winstr_per_sm); }
    { float ms = timeit([&] { k_shfl<<<blocks, threads>>>(out); }); printf("SHFL.BFLY : %.2f cycles per warp-instr per SM\n", ms * 1e-3 * ghz * 1e9 / winstr_per_sm); }
    { float ms = timeit([&] { k_f2i<<<blocks, threads>>>(in, out); }); printf("F2I+FMUL+IADD : %.2f cycles per warp-iter per SM\n", ms * 1e-3 * ghz * 1e9 / winstr_per_sm); }
    { float ms = timeit([&] { k_match<<<blocks, threads>>>(out); }); printf("MATCH.ANY : %.2f cycles per warp-instr per SM\n", ms * 1e-3 * ghz * 1e9 / winstr_per_sm); }
    unsigned long long* grid; CK(cudaMalloc(&grid, (size_t)262144 * 32)); CK(cudaMemset(grid, 0, (size_t)262144 * 32));
    for (int cells : {262144, 40000, 4000}) {
        for (int distinct : {32, 8, 2, 1}) {
            const int iters = 512;
            float ms = timeit([&] { k_redg<<<sms * 8, 256>>>(grid, cells, distinct, iters); });
            double lane_ops = (double)sms * 8 * 256 * iters;
            printf("REDG.64 pool=%6d cells, %2d addr/warp-instr: %.1f G lane-atomics/s (%.2f cycles per warp-instr per SM)\n", cells, distinct,
                   lane_ops / (ms * 1e-3) / 1e9, ms * 1e-3 * ghz * 1e9 / ((double)8 * 8 * iters));
        }
    }
    return 0;
}
