// Probe 2: which TMA forms work on this box?  variant 6: .L2::cache_hint form, 7: 1-D cp.async.bulk, 8: libcu++ cuda::device::experimental API
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
namespace cde = cuda::device::experimental;
using barrier_t = cuda::barrier<cuda::thread_scope_block>;

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k_hint(const __grid_constant__ CUtensorMap map, float* out, int c0, int c1) {
    __shared__ __align__(128) float tile[64];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 256;" :: "r"(s32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                     :: "r"(s32(tile)), "l"(&map), "r"(s32(&bar)), "r"(c0), "r"(c1), "l"(0x1000000000000000ull) : "memory");
    }
    unsigned ok = 0;
    for (unsigned spin = 0; !ok && spin < (1u << 22); ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    if (threadIdx.x < 64) out[threadIdx.x] = ok ? tile[threadIdx.x] : -1.f;
}

__global__ void k_bulk1d(const float* src, float* out) {
    __shared__ __align__(128) float tile[64];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 256;" :: "r"(s32(&bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];"
                     :: "r"(s32(tile)), "l"(src), "r"(s32(&bar)) : "memory");
    }
    unsigned ok = 0;
    for (unsigned spin = 0; !ok && spin < (1u << 22); ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    if (threadIdx.x < 64) out[threadIdx.x] = ok ? tile[threadIdx.x] : -1.f;
}

__global__ void k_libcu(const __grid_constant__ CUtensorMap map, float* out, int c0, int c1) {
    __shared__ alignas(128) float tile[64];
    #pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier_t bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier_t::arrival_token tok;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(tile, &map, c0, c1, bar);
        tok = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile));
    } else tok = bar.arrive();
    bar.wait(std::move(tok));
    if (threadIdx.x < 64) out[threadIdx.x] = tile[threadIdx.x];
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 6; printf("== variant %d\n", variant);
    const int nx = 96, ny = 64;
    std::vector<float> h((size_t)nx * ny);
    for (int y = 0; y < ny; ++y) for (int x = 0; x < nx; ++x) h[x + (size_t)nx * y] = x + 100.f * y;
    float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    float* o; cudaMalloc(&o, 64 * 4); cudaMemset(o, 0, 64 * 4);
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map; memset(&map, 0, sizeof map);
    const cuuint64_t gd[2] = { nx, ny }, gs[1] = { (cuuint64_t)nx * 4 }; const cuuint32_t b[2] = { 8, 8 }, es[2] = { 1, 1 };
    CUresult r = ((EncodeTiled)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gd, gs, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d; map qwords: %016llx %016llx %016llx %016llx\n", (int)r, (unsigned long long)map.opaque[0], (unsigned long long)map.opaque[1], (unsigned long long)map.opaque[2], (unsigned long long)map.opaque[3]);
    if (variant == 6) k_hint<<<1, 128>>>(map, o, 16, 5);
    if (variant == 7) k_bulk1d<<<1, 128>>>(d + 16 + 5 * nx, o);
    if (variant == 8) k_libcu<<<1, 128>>>(map, o, 16, 5);
    cudaError_t e = cudaDeviceSynchronize(); printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> g(64); cudaMemcpy(g.data(), o, 256, cudaMemcpyDeviceToHost);
    printf("first: %g %g ... [9]=%g (want 516 517, [8]= %s)\n", g[0], g[1], g[8], variant == 7 ? "524" : "616");
    return 0;
}
