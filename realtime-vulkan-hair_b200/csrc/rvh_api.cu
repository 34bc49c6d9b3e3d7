// rvh_api.cu -- the extern "C" layer of librvh.so (see include/rvh.h for the contract and
// the reference call site each entry point replaces).  No CPU fallback: every compute
// entry point launches sm_100a kernels on the context's stream or fails.
#include "../../include/rvh.h"
#include "rvh_kernels.cuh"
#include "rvh_host_math.h"

#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace rvh;

namespace {

thread_local std::string g_create_error;

// ---- NCCL, loaded lazily so single-GPU users need no NCCL at all ---------------------
struct NcclId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
        GetUniqueId = (int (*)(NcclId*))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (int (*)(NcclComm*, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllReduce");
        AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllGather");
        CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { err = "libnccl lacks a required symbol"; return false; }
        return true;
    }
};
NcclApi g_nccl;
constexpr int kNcclInt64 = 4, kNcclSum = 0, kNcclChar = 0;

enum { EV_K1 = 0, EV_K2, EV_AR, EV_CLEAR, EV_SPLAT, EV_FINALIZE, EV_COUNT };

}  // namespace

struct rvh_ctx {
    rvh_config cfg;
    int S = 0, S_pad = 0, N = 0, V = 1;
    cudaStream_t stream = nullptr;
    float* planes = nullptr;              // [6][N][S_pad]
    float* corr = nullptr;                // [3][N][S_pad] (RVH_KEEP_CORRECTION)
    unsigned long long* grid = nullptr;   // [G^3][4] int64 accumulators
    float4* fgrid = nullptr;              // [G^3] float cells for the gather (k_grid_finalize)
    int k1_blocks = 0, k1_launch_blocks = 0;   // CTAs of k_ftl_step for all strands / of the launch being issued (chunked launches)
    uint4* k1_clear = nullptr; unsigned k1_clear_n = 0;    // grid clear fused into k_ftl_step (set per step)
    size_t grid_bytes = 0;
    int* perm = nullptr;                  // internal -> external strand index (Morton order)
    bool perm_active = false;             // the planes currently hold the strands in `perm` order (false: external order)
    bool needs_resort = false;            // rvh_step_host's pipeline left the strands in the caller's order: Morton-sort them before the next resident step
    void* aos_dev = nullptr;              // Strand[S] staging / interop target
    size_t aos_bytes = 0;
    void* interop_aos = nullptr;          // imported VkBuffer memory (rvh_import_strands_fd), or a caller-owned device buffer (test hook)
    cudaExternalMemory_t interop_mem = nullptr;
    uint32_t* interop_indirect = nullptr; // imported StrandDrawIndirect buffer (rvh_import_indirect_fd), or caller-owned (test hook)
    cudaExternalMemory_t interop_indirect_mem = nullptr;
    cudaExternalSemaphore_t interop_sem = nullptr;   // signalled after the step's writes into the imported buffers (rvh_import_semaphore_fd)
    void* sort_tmp = nullptr; size_t sort_tmp_bytes = 0;
    unsigned* sort_keys = nullptr; unsigned* sort_keys_out = nullptr; int* sort_ids = nullptr;
    StepParams P;
    bool uploaded = false, colliders_set = false;
    unsigned long long* scene_timing = nullptr;                        // RVH_SCENE_TIMING=1: phase time stamps of CTA 0, printed after every launch (tuning)
    unsigned* scene_bar = nullptr; unsigned scene_bar_base = 0u;       // k_scene_step's grid barrier: monotonic device counter, its value before the next launch
    int wave_steps = 1;                   // RVH_WAVE_STEPS=0 (tuning / A-B): grid-less small scenes take k_ftl_step<MULTI> instead of the step wavefront k_ftl_wave
    int scene_occupancy[2][2] = { { 0, 0 }, { 0, 0 } };
    int scene_ctas_per_sm = 2;            // RVH_SCENE_CTAS env (tuning; 0 = off): CTAs per SM of the persistent small-scene kernel k_scene_step
    int splat_target_warps = 32768;       // RVH_SPLAT_WARPS env (tuning): the splat splits rows over blockIdx.y until it has this many warps
    // rvh_step_n fast paths for small scenes: several steps per launch (grid off), CUDA-graph replay of the step (grid on, wind off)
    cudaGraphExec_t step_graph = nullptr; float step_graph_dt = 0.f; unsigned step_graph_version = 0, state_version = 1; int step_graph_launches = 0;
    // rvh_step_host pipeline: copy streams + per-chunk events
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> pipe_ev;
    bool gather_pending = false;          // fgrid holds a finalized grid whose gather has not been applied to the velocities yet
    // head SDF (extension)
    float* sdf_dev = nullptr;             // [nz][ny][nxp] node values
    int sdf_dim[3] = { 0, 0, 0 }, sdf_nxp = 0;
    int sdf_mode = 0;                     // 0 = no volume, 1 = plain loads, 2 = TMA-staged tiles
    CUtensorMap sdf_map;                  // 3-D tiled map of the volume, box 8x4x4 nodes (zeroed when unused)
    float* bake_tris = nullptr;
    // collider candidate mask (StepParams::cmask)
    unsigned char* cmask_dev = nullptr;
    std::vector<unsigned char> cmask_host;
    std::vector<float> cmask_ell;         // the ellipsoid floats the mask was built from
    // guide -> render strand expansion (k_expand_strands)
    ExpandTables* exp_tab = nullptr; int exp_I = 0, exp_D = 0;
    float4* exp_pw = nullptr; float4* exp_tu = nullptr; size_t exp_cap = 0;   // vertices allocated
    float sdf_cell = 0.f;
    // multi-GPU
    int rank = 0, nranks = 1;
    NcclComm comm = nullptr;
    // fused peer-memory grid exchange (k_grid_exchange): peers' buffers mapped through CUDA IPC
    bool p2p = false;
    ExchangePeers peers;
    unsigned* xflags = nullptr;           // [2][kMaxRanks] epochs + block counter, zero-initialised
    void* ipc_open[3 * kMaxRanks] = {};   // mappings to close
    unsigned epoch = 0;
    int num_sms = 148;
    long long exchange_timeout_cycles = 0;   // RVH_EXCHANGE_TIMEOUT_MS (default 5000) in SM clocks: bound of every wait on a peer in k_grid_exchange
    bool grid_reduced = true;             // the int64 accumulators hold the all-rank sum (NCCL path, or 1 rank)
    // timing
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    float last_ms = 0.f;
    unsigned prof_step = 0;               // profiling mode 2: k_ftl_step launches seen (events on every 4th)
    int profiling = 0;                    // 0 off, 1 every kernel, 2 only k_ftl_step (the roofline kernel: 2 events per step instead of 10)
    bool prof_open = false;
    std::vector<cudaEvent_t> pev;         // pairs, recycled
    std::vector<int> pev_kind; size_t pev_used = 0;
    float prof_ms[EV_COUNT] = { 0, 0, 0, 0, 0, 0 };
    int prof_n[EV_COUNT] = { 0, 0, 0, 0, 0, 0 };
    long long launches = 0;
    std::string err;
};

namespace {

int fail(rvh_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
// A failed runtime call also latches its code as the thread's "last error"; it is cleared here so that a recoverable failure
// (e.g. an import with a bad handle) is not reported again by the cudaGetLastError() check after the next kernel launch.
#define CU(call)                                                                            \
    do { cudaError_t e_ = (call);                                                           \
         if (e_ != cudaSuccess) {                                                           \
             cudaGetLastError();                                                            \
             return fail(ctx, RVH_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } \
    } while (0)

void prof_begin(rvh_ctx* c, int kind) {
    // mode 2: the roofline kernel only, and only every 4th step -- an event record between two kernels costs ~3 us of launch overlap, and on
    // sharded runs every rank's step waits for the slowest rank's (2 x B200: 13 us per step with events around k_ftl_step in every step)
    c->prof_open = c->profiling == 1 || (c->profiling == 2 && kind == EV_K1 && (c->prof_step++ & 3u) == 0u);
    if (!c->prof_open) return;
    if (c->pev_used + 2 > c->pev.size()) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        c->pev.push_back(a); c->pev.push_back(b); c->pev_kind.push_back(kind);
    }
    c->pev_kind[c->pev_used / 2] = kind;
    cudaEventRecord(c->pev[c->pev_used], c->stream);
}
void prof_end(rvh_ctx* c) {
    if (!c->prof_open) return;
    c->prof_open = false;
    cudaEventRecord(c->pev[c->pev_used + 1], c->stream);
    c->pev_used += 2;
}
void prof_collect(rvh_ctx* c) {   // caller has synchronised the stream
    for (size_t i = 0; i + 2 <= c->pev_used; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->pev[i], c->pev[i + 1]) == cudaSuccess) {
            c->prof_ms[c->pev_kind[i / 2]] += ms; c->prof_n[c->pev_kind[i / 2]] += 1;
        }
    }
    c->pev_used = 0;
}

// Collider candidate mask (rvh_kernels.cuh gather_pack): one byte per coarse box of 2x2x2 grid cells, bit j set when ellipsoid
// j (collider j+1) can contain a point of the box.  With q(p) = A p + t the ellipsoid's unit-space map, every p of a box with
// centre c and half-diagonal r has |q(p)| >= |q(c)| - ||A||_F r, so the bit is cleared only where that bound exceeds 1 (plus a
// margin for the rounding of the cell index).  Rebuilt only when an ellipsoid changes; the sphere moves every frame
// (Scene.cpp:110-136) and is tested analytically, outside the mask.
int update_collider_mask(rvh_ctx* ctx, const float* colliders, int n) {
    StepParams& P = ctx->P;
    const int G = P.G, D = G / 2, ne = n > 0 ? n - 1 : 0;
    const bool usable = (ctx->cfg.flags & RVH_GRID_ON) && !(ctx->cfg.flags & RVH_SDF_ON) && (G % 2 == 0) && ne > 0 && ne <= 8 && !std::getenv("RVH_NO_CMASK");
    if (!usable) { P.cmask = nullptr; P.cmask_dim = 0; return RVH_OK; }
    const std::vector<float> ell(colliders + 48, colliders + 48 * (size_t)n);
    if (ctx->cmask_dev && ell == ctx->cmask_ell) return RVH_OK;
    ctx->cmask_host.assign((size_t)D * D * D, 0);
    const float h2 = 2.0f * P.h, r = 1.7320508f * P.h * 1.01f + 1e-4f;
    for (int j = 0; j < ne; ++j) {
        const float* I = colliders + 48 * (size_t)(j + 1) + 16;           // Collider::inv, column-major
        double fro = 0.0;
        for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) fro += (double)I[c * 4 + rr] * I[c * 4 + rr];
        const float bound = 1.0f + (float)std::sqrt(fro) * r + 1e-3f;
        for (int z = 0; z < D; ++z) for (int y = 0; y < D; ++y) for (int x = 0; x < D; ++x) {
            const float cx = P.origin[0] + h2 * ((float)x + 0.5f), cy = P.origin[1] + h2 * ((float)y + 0.5f), cz = P.origin[2] + h2 * ((float)z + 0.5f);
            const float qx = I[0] * cx + I[4] * cy + I[8] * cz + I[12], qy = I[1] * cx + I[5] * cy + I[9] * cz + I[13], qz = I[2] * cx + I[6] * cy + I[10] * cz + I[14];
            if (std::sqrt(qx * qx + qy * qy + qz * qz) <= bound) ctx->cmask_host[(size_t)x + (size_t)D * (y + (size_t)D * z)] |= (unsigned char)(1u << j);
        }
    }
    CU(cudaSetDevice(ctx->cfg.device));
    if (!ctx->cmask_dev) CU(cudaMalloc(&ctx->cmask_dev, ctx->cmask_host.size()));
    CU(cudaMemcpyAsync(ctx->cmask_dev, ctx->cmask_host.data(), ctx->cmask_host.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->cmask_ell = ell;
    P.cmask = ctx->cmask_dev; P.cmask_dim = D;
    return RVH_OK;
}

template <int V, bool WIND, int NELL>
void launch_k1(rvh_ctx* c, int gather) {
    // gather: 0 = none pending, 1 = friction, 2 = friction + repulsion (extension: V <= 2 only, see create_impl)
    if (gather == 2) {
        if constexpr (V <= 2) k_ftl_step<V, WIND, NELL, 2><<<c->k1_launch_blocks, kBlock, 0, c->stream>>>(c->P, c->planes, c->corr, c->fgrid, c->sdf_map, c->k1_clear, c->k1_clear_n);
    } else if (gather == 1) k_ftl_step<V, WIND, NELL, 1><<<c->k1_launch_blocks, kBlock, 0, c->stream>>>(c->P, c->planes, c->corr, c->fgrid, c->sdf_map, c->k1_clear, c->k1_clear_n);
    else if constexpr (NELL < 100) k_ftl_step<V, WIND, NELL, 0><<<c->k1_launch_blocks, kBlock, 0, c->stream>>>(c->P, c->planes, c->corr, c->fgrid, c->sdf_map, c->k1_clear, c->k1_clear_n);
}
template <int V>
void launch_k1_v(rvh_ctx* c, bool wind, int gather) {
    if (c->cfg.flags & RVH_SDF_ON) {      // head SDF in place of the ellipsoids: -3 = TMA-staged tiles, -2 = plain loads
        if constexpr (V <= 2) {
            if (c->sdf_mode == 2) { if (wind) launch_k1<V, true, -3>(c, gather); else launch_k1<V, false, -3>(c, gather); }
            else                  { if (wind) launch_k1<V, true, -2>(c, gather); else launch_k1<V, false, -2>(c, gather); }
        }
        return;
    }
    const bool five = c->P.n_ell == 5;     // the reference scene's collider count gets the unrolled kernel
    // collider candidate mask: pays when the kernel is throughput-bound (several waves of CTAs); with few CTAs the kernel is
    // latency-bound and the mask lookup in front of the ellipsoid tests only lengthens the chain (C3: 0.133 -> 0.149 ms)
    if (five && gather && c->P.cmask && c->k1_blocks >= 4 * 148) {
        if (wind) launch_k1<V, true, 105>(c, gather); else launch_k1<V, false, 105>(c, gather);
        return;
    }
    if (wind) { if (five) launch_k1<V, true, 5>(c, gather); else launch_k1<V, true, -1>(c, gather); }
    else      { if (five) launch_k1<V, false, 5>(c, gather); else launch_k1<V, false, -1>(c, gather); }
}

// k_ftl_step over the CTAs [cta0, cta0 + nctas) (a CTA covers kBlock * V strands)
void launch_ftl(rvh_ctx* ctx, bool wind, int fused_gather, int cta0, int nctas) {
    ctx->P.cta0 = cta0; ctx->k1_launch_blocks = nctas;
    if (ctx->V == 2) launch_k1_v<2>(ctx, wind, fused_gather); else launch_k1_v<1>(ctx, wind, fused_gather);
    ctx->P.cta0 = 0;
    ctx->launches += 1;
}

// k_grid_splat over the strands [strand0, strand0 + nstrands) (multiples of 128)
void launch_splat(rvh_ctx* ctx, int strand0, int nstrands) {
    // enough warps to fill 148 SMs several times over: split the rows when there are few strands
    const int warps = nstrands / 32, rows = ctx->N - 1;
    int chunks = std::max(1, std::min(rows, (ctx->splat_target_warps + warps - 1) / warps));
    const int rpc = (rows + chunks - 1) / chunks;
    chunks = (rows + rpc - 1) / rpc;
    ctx->P.strand0 = strand0;
    if (ctx->P.scale > 0.f && ctx->P.scale < 8388608.0f)
        k_grid_splat<true><<<dim3(nstrands / kSplatThreads, chunks), kSplatThreads, 0, ctx->stream>>>(ctx->P, ctx->planes, ctx->grid, rpc);
    else
        k_grid_splat<false><<<dim3(nstrands / kSplatThreads, chunks), kSplatThreads, 0, ctx->stream>>>(ctx->P, ctx->planes, ctx->grid, rpc);
    ctx->P.strand0 = 0;
    ctx->launches += 1;
}

int launch_gather(rvh_ctx* ctx) {
    prof_begin(ctx, EV_K2);
    const size_t total = (size_t)(ctx->N - 1) * (ctx->S_pad / 2);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 64);
    if (ctx->cfg.flags & RVH_REPULSION_ON) k_grid_gather<true><<<blocks, 256, 0, ctx->stream>>>(ctx->P, ctx->planes, ctx->fgrid);
    else                                   k_grid_gather<false><<<blocks, 256, 0, ctx->stream>>>(ctx->P, ctx->planes, ctx->fgrid);
    prof_end(ctx);
    ctx->launches += 1;
    CU(cudaGetLastError());
    ctx->gather_pending = false;
    return RVH_OK;
}

// Apply a deferred gather before anything reads velocities back.
int flush_gather(rvh_ctx* ctx) { return ctx->gather_pending ? launch_gather(ctx) : RVH_OK; }

// After the stream has been synchronised: did a k_grid_exchange give up waiting for a peer?
int check_exchange_error(rvh_ctx* ctx) {
    if (ctx->nranks <= 1 || !ctx->p2p || !ctx->xflags) return RVH_OK;
    unsigned w = 0;
    CU(cudaMemcpy(&w, ctx->xflags + kXErrWord, sizeof w, cudaMemcpyDeviceToHost));
    if (w == 0) return RVH_OK;
    return fail(ctx, RVH_ERR_NCCL, "grid exchange timed out waiting for rank " + std::to_string(w - 1) + " (a rank died or did not call rvh_step in lockstep); the state of this context is undefined");
}

int allreduce_grid(rvh_ctx* ctx) {
    if (ctx->grid_reduced) return RVH_OK;
    prof_begin(ctx, EV_AR);
    int r = g_nccl.AllReduce(ctx->grid, ctx->grid, ctx->grid_bytes / 8, kNcclInt64, kNcclSum, ctx->comm, ctx->stream);
    prof_end(ctx);
    if (r != 0) return fail(ctx, RVH_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
    ctx->grid_reduced = true;
    return RVH_OK;
}

// the time-only wind scalars of one step (compute.comp:151-152): evaluated on the host, once per step
void wind_scalars(int wind_mode, float total_time, float& s2T, float& T3, float& amp) {
    s2T = 2.0f * std::sin(total_time * 2.0f);
    T3 = (float)std::fmod((double)(total_time * 3.0f), 6.283185307179586);   // sin(5z + 3T): keep the argument bounded
    amp = (wind_mode == 2) ? 7.0f * wind_fbm(total_time) : 10.0f;
}

// Grid off: `n` steps (<= 32) in ONE launch of k_ftl_step<..., MULTI> (strands are independent; see the kernel).
int launch_multi_step(rvh_ctx* ctx, int n, float dt, float& t) {      // t advances by n float additions of dt, exactly as n calls of rvh_step_n(1) would
    StepParams& P = ctx->P;
    const int flags = ctx->cfg.flags;
    const bool wind = flags & (RVH_WIND_A | RVH_WIND_B);
    P.dt = dt; P.inv_dt = 1.0f / dt; P.dt2 = dt * dt; P.vel_scale = P.damping / dt;
    P.wind_mode = (flags & RVH_WIND_B) ? 2 : ((flags & RVH_WIND_A) ? 1 : 0);
    P.multi_steps = n;
    for (int k = 0; k < n; ++k, t += dt) {
        float s2T = 0.f, T3 = 0.f, amp = 0.f;
        if (wind) wind_scalars(P.wind_mode, t, s2T, T3, amp);
        P.wind_tab[3 * k] = amp * s2T; P.wind_tab[3 * k + 1] = T3; P.wind_tab[3 * k + 2] = amp;
    }
    P.cta0 = 0;
    const bool five = P.n_ell == 5;       // the reference scene's collider count: unrolled tests (the kernel is latency-bound here: the unrolled tests overlap)
    // Small scene (its k_ftl_step CTAs would leave the schedulers nearly empty): wavefront over the steps, W lanes per strand (k_ftl_wave)
    const int wave_w = (ctx->V != 1 || n < 2 || ctx->N < 3 || !ctx->wave_steps) ? 0 : (ctx->k1_blocks * 8 <= 6 * ctx->num_sms ? 8 : (ctx->k1_blocks * 4 <= 6 * ctx->num_sms ? 4 : 0));
    if (wave_w) {
        const int blocks = ctx->S_pad * wave_w / kBlock;
#define RVH_WAVE(W_, NE_, L_) k_ftl_wave<W_, NE_, L_><<<blocks, kBlock, 0, ctx->stream>>>(P, ctx->planes, ctx->corr)
        if (wave_w == 8) {
            if (five) { if (wind) RVH_WAVE(true, 5, 8); else RVH_WAVE(false, 5, 8); }
            else      { if (wind) RVH_WAVE(true, -1, 8); else RVH_WAVE(false, -1, 8); }
        } else {
            if (five) { if (wind) RVH_WAVE(true, 5, 4); else RVH_WAVE(false, 5, 4); }
            else      { if (wind) RVH_WAVE(true, -1, 4); else RVH_WAVE(false, -1, 4); }
        }
#undef RVH_WAVE
        P.multi_steps = 1;
        ctx->launches += 1;
        CU(cudaGetLastError());
        return RVH_OK;
    }
#define RVH_MULTI(V_, W_, NE_) k_ftl_step<V_, W_, NE_, 0, true><<<ctx->k1_blocks, kBlock, 0, ctx->stream>>>(P, ctx->planes, ctx->corr, ctx->fgrid, ctx->sdf_map, nullptr, 0u)
    if (ctx->V == 2) {
        if (five) { if (wind) RVH_MULTI(2, true, 5); else RVH_MULTI(2, false, 5); }
        else      { if (wind) RVH_MULTI(2, true, -1); else RVH_MULTI(2, false, -1); }
    } else {
        if (five) { if (wind) RVH_MULTI(1, true, 5); else RVH_MULTI(1, false, 5); }
        else      { if (wind) RVH_MULTI(1, true, -1); else RVH_MULTI(1, false, -1); }
    }
#undef RVH_MULTI
    P.multi_steps = 1;
    ctx->launches += 1;
    CU(cudaGetLastError());
    return RVH_OK;
}

// Small scene, grid on: `n` steps (<= 32) in ONE cooperative launch of k_scene_step (see the kernel).
bool scene_step_eligible(const rvh_ctx* ctx) {
    const int flags = ctx->cfg.flags;
    return ctx->scene_ctas_per_sm > 0 && (flags & RVH_GRID_ON) && !(flags & (RVH_SDF_ON | RVH_REPULSION_ON)) && ctx->nranks == 1 &&
           ctx->uploaded && ctx->colliders_set && !ctx->needs_resort && ctx->profiling == 0 &&
           !ctx->interop_aos && !ctx->interop_indirect && !ctx->interop_sem &&
           ctx->V <= 2 && ctx->k1_blocks * 2 <= ctx->num_sms && ctx->P.scale > 0.f && ctx->P.scale < 8388608.0f;
}

template <int V, bool WIND, int NELL>
int launch_scene_kernel(rvh_ctx* ctx, int n, int blocks, int splat_bx, int splat_items, int rpc) {
    void* fn = (void*)k_scene_step<V, WIND, NELL>;
    if (blocks == 0) {                                                  // query: co-resident CTAs per SM
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scene_step<V, WIND, NELL>, kBlock, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
        return per_sm;
    }
    int k1 = ctx->k1_blocks;
    void* args[] = { (void*)&ctx->P, (void*)&ctx->planes, (void*)&ctx->corr, (void*)&ctx->grid, (void*)&n, (void*)&k1,
                     (void*)&splat_bx, (void*)&splat_items, (void*)&rpc, (void*)&ctx->scene_bar, (void*)&ctx->scene_bar_base, (void*)&ctx->scene_timing };
    return (int)cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(kBlock), args, 0, ctx->stream);
}

int scene_kernel_dispatch(rvh_ctx* ctx, bool wind, int n, int blocks, int splat_bx, int splat_items, int rpc) {
    const bool five = ctx->P.n_ell == 5;
#define RVH_SCENE(V_, W_, NE_) launch_scene_kernel<V_, W_, NE_>(ctx, n, blocks, splat_bx, splat_items, rpc)
    if (ctx->V == 2) {
        if (five) return wind ? RVH_SCENE(2, true, 5) : RVH_SCENE(2, false, 5);
        return wind ? RVH_SCENE(2, true, -1) : RVH_SCENE(2, false, -1);
    }
    if (five) return wind ? RVH_SCENE(1, true, 5) : RVH_SCENE(1, false, 5);
    return wind ? RVH_SCENE(1, true, -1) : RVH_SCENE(1, false, -1);
#undef RVH_SCENE
}

// Returns kSceneFallback (and switches the path off for this context) when the device refuses the cooperative launch: the caller
// then takes the launch-per-kernel step; nothing has been stepped and `t_io` is untouched in that case.
constexpr int kSceneFallback = 1;
int launch_scene_steps(rvh_ctx* ctx, int n, float dt, float& t_io) {   // t_io advances exactly as n calls of rvh_step would move it
    { int r = flush_gather(ctx); if (r) return r; }                     // an ordinary step left its gather to its successor: apply it now
    StepParams& P = ctx->P;
    const int flags = ctx->cfg.flags;
    const bool wind = flags & (RVH_WIND_A | RVH_WIND_B);
    P.dt = dt; P.inv_dt = 1.0f / dt; P.dt2 = dt * dt; P.vel_scale = P.damping / dt;
    P.wind_mode = (flags & RVH_WIND_B) ? 2 : ((flags & RVH_WIND_A) ? 1 : 0);
    float t = t_io;
    for (int k = 0; k < n; ++k, t += dt) {
        float s2T = 0.f, T3 = 0.f, amp = 0.f;
        if (wind) wind_scalars(P.wind_mode, t, s2T, T3, amp);
        P.wind_tab[3 * k] = amp * s2T; P.wind_tab[3 * k + 1] = T3; P.wind_tab[3 * k + 2] = amp;
    }
    P.cta0 = 0; P.strand0 = 0;
    int& occ = ctx->scene_occupancy[wind ? 1 : 0][P.n_ell == 5 ? 1 : 0];  // co-resident CTAs per SM of this kernel variant: asked once (the per-frame call counts its host microseconds)
    if (occ == 0) occ = scene_kernel_dispatch(ctx, wind, n, 0, 0, 0, 0);
    const int per_sm = std::min(ctx->scene_ctas_per_sm, occ);
    if (per_sm < 1) { ctx->scene_ctas_per_sm = 0; return kSceneFallback; }
    const int max_blocks = per_sm * ctx->num_sms;
    // The launch is sized to the scene, not to the machine: every barrier costs one atomic per CTA.  Enough CTAs for the FTL chains plus
    // a clearing crew beside them, for one splat item (128 strands x a chunk of rows) per CTA, and for one gather point per thread.
    const int splat_bx = ctx->S_pad / kSplatThreads, rows = ctx->N - 1;
    const int points = rows * ctx->S_pad;
    int blocks = std::max(ctx->k1_blocks + std::max(16, ctx->k1_blocks), (points + kBlock - 1) / kBlock);
    blocks = std::max(blocks, std::min(splat_bx * rows, 2 * ctx->num_sms));
    blocks = std::min(blocks, max_blocks);
    int chunks = std::max(1, std::min(rows, (blocks + splat_bx - 1) / splat_bx));
    const int rpc = (rows + chunks - 1) / chunks;
    chunks = (rows + rpc - 1) / rpc;
    if (!ctx->scene_timing && std::getenv("RVH_SCENE_TIMING")) CU(cudaMalloc(&ctx->scene_timing, (1 + 6 * 32) * sizeof(unsigned long long)));
    if (!ctx->scene_bar) {
        CU(cudaMalloc(&ctx->scene_bar, sizeof(unsigned)));
        CU(cudaMemsetAsync(ctx->scene_bar, 0, sizeof(unsigned), ctx->stream));
        ctx->scene_bar_base = 0u;
    }
    const int e = scene_kernel_dispatch(ctx, wind, n, blocks, splat_bx, splat_bx * chunks, rpc);
    if (e != 0) { cudaGetLastError(); ctx->scene_ctas_per_sm = 0; return kSceneFallback; }   // e.g. no cooperative launch under this driver configuration
    ctx->scene_bar_base += (3u * (unsigned)n - 1u) * (unsigned)blocks;  // the counter is monotonic (compared modulo 2^32); no barrier after the last gather
    if (ctx->scene_timing) {                                            // tuning aid: mean duration of every phase and barrier over the launch's steps
        std::vector<unsigned long long> ts(1 + 6 * (size_t)n);
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaMemcpy(ts.data(), ctx->scene_timing, ts.size() * sizeof(ts[0]), cudaMemcpyDeviceToHost));
        double acc[6] = { 0, 0, 0, 0, 0, 0 };
        for (int k = 0; k < n; ++k) for (int j = 0; j < 6; ++j) acc[j] += (double)(ts[1 + 6 * k + j] - ts[6 * k + j]);
        std::fprintf(stderr, "[k_scene_step] %d CTAs, %d steps, CTA 0 mean ns: ftl|clear %.0f  barrier %.0f  splat %.0f  barrier %.0f  gather %.0f  barrier %.0f  = %.0f per step\n",
                     blocks, n, acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n, acc[5] / n, (double)(ts[6 * (size_t)n] - ts[0]) / n);
    }
    ctx->launches += 1;
    ctx->gather_pending = false;                                        // applied inside the launch; the float grid was not produced
    t_io = t;
    return RVH_OK;
}

// The step's grid is complete on this rank: make the float grid the gather reads (compute.comp:276-286), summed over the ranks.
int finalize_grid(rvh_ctx* ctx) {
    const int cells = ctx->P.G * ctx->P.G * ctx->P.G;
    if (ctx->nranks > 1 && ctx->p2p && !ctx->grid_reduced) {
        // reduce-scatter + finalize + all-gather in one kernel over NVLink peer memory (every rank calls in lockstep)
        prof_begin(ctx, EV_AR);
        ctx->epoch += 1;
        const int per = (cells + ctx->nranks - 1) / ctx->nranks;
        // one cell per thread when the slice allows it (all blocks co-resident: <= 4 per SM): every thread then pays ONE NVLink round trip,
        // not one per loop iteration (2 ranks: 131,072 cells per slice used to take 3-4 trips through 148 blocks)
        k_grid_exchange<<<std::min((per + 255) / 256, 4 * ctx->num_sms), 256, 0, ctx->stream>>>(ctx->peers, ctx->rank, ctx->nranks, cells, ctx->P.int32_wrap, ctx->epoch, ctx->exchange_timeout_cycles);
        prof_end(ctx);
    } else {
        prof_begin(ctx, EV_FINALIZE);
        k_grid_finalize<<<(cells + 255) / 256, 256, 0, ctx->stream>>>((const long long*)ctx->grid, ctx->fgrid, cells, ctx->P.int32_wrap);
        prof_end(ctx);
    }
    ctx->launches += 1;
    CU(cudaGetLastError());
    ctx->gather_pending = true;
    return RVH_OK;
}

}  // namespace
extern "C" {
static int unpack_from_staging(rvh_ctx* ctx);
static int pack_to_staging(rvh_ctx* ctx);
}
namespace {

// phases: bit 0 = integrate + FTL (+ splat, all-reduce), bit 1 = grid finalize + gather.
// lazy: leave the gather to the next step's k_ftl_step (steady-state stepping); otherwise run it now.
int do_step(rvh_ctx* ctx, float dt, float total_time, int phases, bool lazy) {
    if (!ctx->uploaded) return fail(ctx, RVH_ERR_STATE, "rvh_step before rvh_upload_strands_aos");
    if (!ctx->colliders_set) return fail(ctx, RVH_ERR_STATE, "rvh_step before rvh_set_colliders");
    if (!(dt > 0.f)) return fail(ctx, RVH_ERR_INVALID, "dt must be > 0");
    if (ctx->needs_resort) {
        // the last state arrived through rvh_step_host's chunk pipeline, in the caller's strand order: restore the Morton order the
        // splat's warp aggregation lives on (device-side: pack to the staging buffer, sort the roots, unpack) before stepping on
        ctx->needs_resort = false;
        { int r = pack_to_staging(ctx); if (r) return r; }
        { int r = unpack_from_staging(ctx); if (r) return r; }
    }
    StepParams& P = ctx->P;
    P.dt = dt; P.inv_dt = 1.0f / dt; P.dt2 = dt * dt; P.vel_scale = P.damping / dt;
    const int flags = ctx->cfg.flags;
    const bool grid = flags & RVH_GRID_ON, wind = flags & (RVH_WIND_A | RVH_WIND_B);
    P.wind_mode = (flags & RVH_WIND_B) ? 2 : ((flags & RVH_WIND_A) ? 1 : 0);
    if (wind) wind_scalars(P.wind_mode, total_time, P.wind_s2T, P.wind_T3, P.wind_amp);
    if ((flags & RVH_SDF_ON) && !ctx->sdf_dev)
        return fail(ctx, RVH_ERR_STATE, "RVH_SDF_ON but no head SDF: call rvh_set_head_sdf or rvh_bake_head_sdf_* first");
    if (phases & 1) {
        // Renderer.cpp:2063.  Wide launches clear the grid from inside k_ftl_step (<= 4 stores of 16 bytes per thread)
        const size_t clear_n = ctx->grid_bytes / 16;
        const bool fused_clear = grid && (size_t)ctx->k1_blocks * kBlock * 4 >= clear_n && clear_n < ((size_t)1 << 32) && !std::getenv("RVH_NO_FUSED_CLEAR");
        ctx->k1_clear = fused_clear ? reinterpret_cast<uint4*>(ctx->grid) : nullptr;
        ctx->k1_clear_n = fused_clear ? (unsigned)clear_n : 0u;
        if (grid && !fused_clear) {
            prof_begin(ctx, EV_CLEAR);
            CU(cudaMemsetAsync(ctx->grid, 0, ctx->grid_bytes, ctx->stream));
            prof_end(ctx);
        }
        prof_begin(ctx, EV_K1);
        const int fused_gather = ctx->gather_pending ? ((flags & RVH_REPULSION_ON) ? 2 : 1) : 0;
        launch_ftl(ctx, wind, fused_gather, 0, ctx->k1_blocks);
        prof_end(ctx);
        ctx->gather_pending = false;
        CU(cudaGetLastError());
        if (grid) {
            prof_begin(ctx, EV_SPLAT);
            launch_splat(ctx, 0, ctx->S_pad);
            prof_end(ctx);
            CU(cudaGetLastError());
        }
        if (grid && ctx->nranks > 1) {
            ctx->grid_reduced = false;
            if (!ctx->p2p) { int r = allreduce_grid(ctx); if (r) return r; }
        }
    }
    if ((phases & 2) && grid) {
        { int r = finalize_grid(ctx); if (r) return r; }
        // (Late round 2: running the stand-alone gather right away on latency-bound mid-size scenes, so that k_ftl_step's chain loses
        // the gather, was measured on C3 100K x 64: k_ftl_step 0.133 -> 0.084 ms, but k_grid_gather costs 0.059 ms: no gain.)
        if (!lazy || ctx->interop_aos) { int r = launch_gather(ctx); if (r) return r; }
    }
    if (ctx->interop_indirect) {                                        // compute.comp:126-130,302: the draw's vertexCount = strands
        k_write_indirect<<<1, 32, 0, ctx->stream>>>(ctx->interop_indirect, (uint32_t)ctx->S);
        ctx->launches += 1;
        CU(cudaGetLastError());
    }
    if (ctx->interop_aos) {
        const int tiles = (ctx->S + kTile - 1) / kTile;
        const size_t sm = (size_t)9 * ctx->N * (kTile + 1) * sizeof(float);
        CU(cudaFuncSetAttribute(k_pack_aos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));   // per function and device, not per context: set per launch
        k_pack_aos<<<tiles, 256, sm, ctx->stream>>>((float4*)ctx->interop_aos, ctx->planes, ctx->corr, ctx->perm_active ? ctx->perm : nullptr, ctx->S, ctx->S_pad, ctx->N, 0, 7);
        ctx->launches += 1;
        CU(cudaGetLastError());
    }
    if (ctx->interop_sem) {                                             // the graphics submit waits on this (VK_KHR_external_semaphore)
        cudaExternalSemaphoreSignalParams sp; std::memset(&sp, 0, sizeof sp);
        CU(cudaSignalExternalSemaphoresAsync(&ctx->interop_sem, &sp, 1, ctx->stream));
    }
    return RVH_OK;
}

// Map every peer's accumulators, float grid and flag block into this process (CUDA IPC) for k_grid_exchange.
// All ranks take the same decision: the handles carry an "ok" byte and p2p is used only if every rank could
// export AND open everything.  RVH_GRID_EXCHANGE=nccl forces the NCCL path.
void setup_peer_exchange(rvh_ctx* c) {
    struct Blob { cudaIpcMemHandle_t h[3]; int ok; int pad[15]; };
    static_assert(sizeof(Blob) == 256, "blob layout");
    const int R = c->nranks;
    c->p2p = false;
    if (R > kMaxRanks || !g_nccl.AllGather) return;
    const char* env = std::getenv("RVH_GRID_EXCHANGE");
    Blob mine; std::memset(&mine, 0, sizeof mine);
    mine.ok = !(env && std::strcmp(env, "nccl") == 0);
    if (cudaMalloc(&c->xflags, (2 * kMaxRanks + 16) * sizeof(unsigned)) != cudaSuccess) mine.ok = 0;
    else cudaMemsetAsync(c->xflags, 0, (2 * kMaxRanks + 16) * sizeof(unsigned), c->stream);
    if (mine.ok) {
        if (cudaIpcGetMemHandle(&mine.h[0], c->grid) != cudaSuccess || cudaIpcGetMemHandle(&mine.h[1], c->fgrid) != cudaSuccess ||
            cudaIpcGetMemHandle(&mine.h[2], c->xflags) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    }
    Blob* dsend = nullptr; Blob* drecv = nullptr;
    std::vector<Blob> all(R);
    bool gathered = cudaMalloc(&dsend, sizeof(Blob)) == cudaSuccess && cudaMalloc(&drecv, sizeof(Blob) * R) == cudaSuccess &&
                    cudaMemcpyAsync(dsend, &mine, sizeof(Blob), cudaMemcpyHostToDevice, c->stream) == cudaSuccess &&
                    g_nccl.AllGather(dsend, drecv, sizeof(Blob), kNcclChar, c->comm, c->stream) == 0 &&
                    cudaMemcpyAsync(all.data(), drecv, sizeof(Blob) * R, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
                    cudaStreamSynchronize(c->stream) == cudaSuccess;
    int ok = gathered ? 1 : 0;
    for (int r = 0; ok && r < R; ++r) ok = all[r].ok;
    if (ok) {
        for (int r = 0; r < R && ok; ++r) {
            void* ptr[3] = { c->grid, c->fgrid, c->xflags };
            if (r != c->rank) {
                for (int k = 0; k < 3 && ok; ++k) {
                    if (cudaIpcOpenMemHandle(&ptr[k], all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
                    else c->ipc_open[3 * r + k] = ptr[k];
                }
            }
            c->peers.grid[r] = (const long long*)ptr[0]; c->peers.fgrid[r] = (float4*)ptr[1]; c->peers.flags[r] = (unsigned*)ptr[2];
        }
    }
    // second round: everybody must have opened everything
    mine.ok = ok;
    if (gathered && cudaMemcpyAsync(dsend, &mine, sizeof(Blob), cudaMemcpyHostToDevice, c->stream) == cudaSuccess &&
        g_nccl.AllGather(dsend, drecv, sizeof(Blob), kNcclChar, c->comm, c->stream) == 0 &&
        cudaMemcpyAsync(all.data(), drecv, sizeof(Blob) * R, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
        cudaStreamSynchronize(c->stream) == cudaSuccess) {
        for (int r = 0; ok && r < R; ++r) ok = all[r].ok;
    } else ok = 0;
    cudaFree(dsend); cudaFree(drecv);
    c->p2p = ok != 0;
    {
        double ms = 5000.0;
        if (const char* e = std::getenv("RVH_EXCHANGE_TIMEOUT_MS")) ms = std::max(1.0, std::atof(e));
        int khz = 1965000;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->cfg.device);
        c->exchange_timeout_cycles = (long long)(ms * (double)khz);
    }
}

int create_impl(rvh_ctx** out, const rvh_config* cfg, int rank, int nranks, const void* uid) {
    rvh_ctx* ctx = nullptr;   // for CU(): errors go to g_create_error
    if (!out || !cfg) return fail(nullptr, RVH_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->num_strands < 1 || cfg->num_points < 2) return fail(nullptr, RVH_ERR_INVALID, "need num_strands >= 1 and num_points >= 2");
    // the AoS <-> planes transposes stage 9 planes x N points x 33 strands of floats in shared memory (227 KB per CTA on sm_100)
    if ((size_t)9 * cfg->num_points * (kTile + 1) * sizeof(float) > (size_t)227 * 1024)
        return fail(nullptr, RVH_ERR_INVALID, "num_points too large: the Strand[S] pack/unpack kernels need 9*N*33*4 bytes of shared memory (N <= 195)");
    if (cfg->grid_dim < 2 || cfg->grid_dim > 1023) return fail(nullptr, RVH_ERR_INVALID, "grid_dim out of range (2..1023: the splat's cell keys hold 10 bits per axis)");
    if ((size_t)cfg->num_strands * cfg->num_points > ((size_t)1 << 31)) return fail(nullptr, RVH_ERR_INVALID, "S*N too large for one context");
    if ((cfg->flags & RVH_REPULSION_ON) && !(cfg->flags & RVH_GRID_ON)) return fail(nullptr, RVH_ERR_INVALID, "RVH_REPULSION_ON needs RVH_GRID_ON (it reads the same voxel grid)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RVH_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, RVH_ERR_INVALID, "bad device ordinal");
    CU(cudaSetDevice(cfg->device));

    rvh_ctx* c = new rvh_ctx();
    c->cfg = *cfg;
    c->S = cfg->num_strands; c->N = cfg->num_points;
    c->S_pad = ((c->S + kTileStrands - 1) / kTileStrands) * kTileStrands;
    c->rank = rank; c->nranks = nranks;
    if (const char* e = std::getenv("RVH_WAVE_STEPS")) c->wave_steps = std::atoi(e) != 0;
    if (const char* e = std::getenv("RVH_SCENE_CTAS")) c->scene_ctas_per_sm = std::max(0, std::min(4, std::atoi(e)));
    if (const char* e = std::getenv("RVH_SPLAT_WARPS")) c->splat_target_warps = std::max(1, std::atoi(e));
    int V = cfg->strands_per_thread;
    if (V == 4) V = 2;                                             // round 1 shipped a 4-strand kernel: it spilled (79-106 local ops) and never won; the value is still accepted
    // measured on B200 (scripts/probes/spt_sweep.py, grid + wind, ms per step 1 vs 2 strands per thread): 100K x 64 0.235 / 0.258,
    // 150K x 32 0.181 / 0.177, 250K x 32 0.259 / 0.254, 600K x 32 0.528 / 0.493, 1M x 32: 2 by 10 %.  Packs halve the issue slots, which
    // pays once the launch is throughput-bound; a single wave of CTAs is a latency chain and wants twice the warps instead.
    if (V != 1 && V != 2) V = c->S >= 131072 ? 2 : 1;
    c->V = V;
    std::memset(&c->sdf_map, 0, sizeof c->sdf_map);
    ctx = c;
#define CUC(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e2_); rvh_destroy(c); return fail(nullptr, RVH_ERR_CUDA, m_); } } while (0)
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const size_t plane = (size_t)c->N * c->S_pad * sizeof(float);
    CUC(cudaMalloc(&c->planes, 6 * plane));
    CUC(cudaMemsetAsync(c->planes, 0, 6 * plane, c->stream));
    if (cfg->flags & RVH_KEEP_CORRECTION) {
        CUC(cudaMalloc(&c->corr, 3 * plane));
        CUC(cudaMemsetAsync(c->corr, 0, 3 * plane, c->stream));
    }
    const size_t G = cfg->grid_dim;
    c->grid_bytes = G * G * G * 4 * sizeof(long long);
    CUC(cudaMalloc(&c->grid, c->grid_bytes));
    CUC(cudaMemsetAsync(c->grid, 0, c->grid_bytes, c->stream));
    CUC(cudaMalloc(&c->fgrid, G * G * G * sizeof(float4)));
    CUC(cudaMemsetAsync(c->fgrid, 0, G * G * G * sizeof(float4), c->stream));
    c->k1_blocks = ((c->S_pad + V - 1) / V + kBlock - 1) / kBlock;
    c->aos_bytes = (size_t)c->S * 48 * c->N;
    CUC(cudaMalloc(&c->aos_dev, c->aos_bytes));
    CUC(cudaEventCreate(&c->ev_a)); CUC(cudaEventCreate(&c->ev_b));
    CUC(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));

    StepParams& P = c->P;
    std::memset(&P, 0, sizeof P);
    P.S = c->S; P.S_pad = c->S_pad; P.N = c->N;
    P.rest = cfg->rest_length; P.gravity_y = cfg->gravity_y; P.damping = cfg->damping;
    P.vmax = cfg->vmax; P.vmax2 = cfg->vmax * cfg->vmax; P.penalty_k = cfg->penalty_k;
    P.sphere_r = cfg->sphere_radius; P.sphere_r2 = cfg->sphere_radius * cfg->sphere_radius;
    P.G = cfg->grid_dim; P.h = cfg->grid_extent / (float)cfg->grid_dim;    // compute.comp:205
    P.rh = 1.0f / P.h;
    {   // Markstein division needs a normal h whose significand is not all ones (rvh_kernels.cuh, grid_coord)
        uint32_t hb; std::memcpy(&hb, &P.h, 4);
        P.div_fast = (std::isnormal(P.h) && P.h > 1e-20f && P.h < 1e20f && (hb & 0x7fffffu) != 0x7fffffu) ? 1 : 0;
    }
    for (int k = 0; k < 3; ++k) P.origin[k] = cfg->grid_origin[k];
    P.scale = cfg->grid_scale; P.friction = cfg->friction;
    P.repulsion = cfg->repulsion; P.inv_h = 1.0f / P.h;
    P.int32_wrap = (cfg->flags & RVH_GRID_INT32_WRAP) ? 1 : 0;
    {   // k_grid_splat sums the 32 points of a row in int32 registers: a contribution is RN(scale * RN(tw * v)) with tw <= 1, bounded
        // by RN(scale * |v|), and must stay below 2^26 (the density term, scale itself, too).  Faster points -- or every point when
        // grid_scale is that large -- go to the grid one by one with 64-bit conversions (splat_point_direct).
        const float lim = 67108863.0f;
        float v = -1.0f;
        if (P.scale > 0.f && P.scale < lim) {
            v = lim / P.scale;
            while (v > 0.f && P.scale * v > lim) v = std::nextafterf(v, 0.f);
        }
        P.splat_vagg = v;
    }
    P.keep_corr = c->corr ? 1 : 0;
    // sparse hair: with a few thousand strands a warp's 32 Morton neighbours stop sharing cells (RVH_SPLAT_SPARSE: the distinct-cell
    // count from which a row is splatted point by point; 0 = never).  Not for dense scenes: forced on at C3 (100K x 64) the 32 atomics
    // per point take the splat from 0.136 to 0.305 ms.
    P.splat_sparse = c->S < 16384 ? 5 : 0;
    if (const char* e = std::getenv("RVH_SPLAT_SPARSE")) P.splat_sparse = std::max(0, std::atoi(e));

    if (nranks > 1) {
        std::string err;
        if (!g_nccl.load(err)) { rvh_destroy(c); return fail(nullptr, RVH_ERR_NCCL, err); }
        NcclId id; std::memcpy(&id, uid, sizeof id);
        int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
        if (r != 0) { rvh_destroy(c); return fail(nullptr, RVH_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error")); }
        setup_peer_exchange(c);           // falls back to the NCCL all-reduce when peer mapping is not possible
    }
    CUC(cudaStreamSynchronize(c->stream));
#undef CUC
    *out = c;
    return RVH_OK;
}

}  // namespace

extern "C" {

int rvh_abi_version(void) { return 3; }   // 2: rvh_config.repulsion, head SDF entry points; 3: indirect / semaphore imports, debug hooks

void rvh_default_config(rvh_config* cfg, int num_strands, int num_points) {
    std::memset(cfg, 0, sizeof *cfg);
    cfg->device = 0;
    cfg->num_strands = num_strands; cfg->num_points = num_points;
    cfg->rest_length = 2.5f / ((float)num_points - 1.0f);
    cfg->gravity_y = -9.8f; cfg->damping = 0.998f; cfg->vmax = 10.0f; cfg->penalty_k = 1900.0f;
    cfg->sphere_radius = 1.0f;
    cfg->grid_dim = 64; cfg->grid_extent = 7.0f;
    cfg->grid_origin[0] = -3.0f; cfg->grid_origin[1] = -2.0f; cfg->grid_origin[2] = -5.0f;
    cfg->grid_scale = 1000000.0f; cfg->friction = 0.08f;
    cfg->flags = RVH_GRID_ON;
    cfg->strands_per_thread = 0;
    cfg->repulsion = 0.2f;
}

int rvh_create(rvh_ctx** out, const rvh_config* cfg) { return create_impl(out, cfg, 0, 1, nullptr); }

int rvh_nccl_unique_id(void* out128) {
    std::string err;
    if (!out128) return fail(nullptr, RVH_ERR_INVALID, "null argument");
    if (!g_nccl.load(err)) return fail(nullptr, RVH_ERR_NCCL, err);
    NcclId id;
    if (g_nccl.GetUniqueId(&id) != 0) return fail(nullptr, RVH_ERR_NCCL, "ncclGetUniqueId failed");
    std::memcpy(out128, &id, sizeof id);
    return RVH_OK;
}

int rvh_create_sharded(rvh_ctx** out, const rvh_config* cfg, int rank, int nranks, const void* uid) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(nullptr, RVH_ERR_INVALID, "bad rank / nranks");
    if (nranks > 1 && !uid) return fail(nullptr, RVH_ERR_INVALID, "nccl_unique_id required when nranks > 1");
    return create_impl(out, cfg, rank, nranks, uid);
}

int rvh_exchange_mode(rvh_ctx* ctx) { return !ctx || ctx->nranks <= 1 ? 0 : (ctx->p2p ? 2 : 1); }

int rvh_set_colliders(rvh_ctx* ctx, const void* colliders, int n) {
    if (!ctx) return RVH_ERR_INVALID;
    if (n < 0 || n > 1 + kMaxEllipsoids || (n > 0 && !colliders)) return fail(ctx, RVH_ERR_INVALID, "0 <= n <= 8 colliders of 192 bytes");
    const float* c = (const float*)colliders;
    StepParams& P = ctx->P;
    P.has_sphere = n > 0; P.n_ell = n > 0 ? n - 1 : 0;
    if (n > 0) { P.sphere_c[0] = c[12]; P.sphere_c[1] = c[13]; P.sphere_c[2] = c[14]; }   // transform[3].xyz, compute.comp:162
    for (int j = 1; j < n; ++j) {
        const float* X = c + 48 * j; const float* I = X + 16; const float* IT = X + 32;
        Ellipsoid& E = P.ell[j - 1];
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 4; ++k) { E.inv[r * 4 + k] = I[k * 4 + r]; E.xf[r * 4 + k] = X[k * 4 + r]; }
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) E.nt[r * 3 + k] = IT[k * 4 + r];
    }
    ctx->colliders_set = true;
    ctx->state_version += 1;                                       // a captured step holds the old colliders in its kernel parameters
    return update_collider_mask(ctx, c, n);
}

static int unpack_from_staging(rvh_ctx* ctx) {
    const bool reorder = !(ctx->cfg.flags & RVH_KEEP_ORDER) && ctx->S >= 1024;
    if (reorder) {
        if (!ctx->perm) {
            CU(cudaMalloc(&ctx->perm, sizeof(int) * ctx->S));
            CU(cudaMalloc(&ctx->sort_keys, sizeof(unsigned) * ctx->S));
            CU(cudaMalloc(&ctx->sort_keys_out, sizeof(unsigned) * ctx->S));
            CU(cudaMalloc(&ctx->sort_ids, sizeof(int) * ctx->S));
            CU(cub::DeviceRadixSort::SortPairs(nullptr, ctx->sort_tmp_bytes, ctx->sort_keys, ctx->sort_keys_out, ctx->sort_ids, ctx->perm, ctx->S, 0, 30, ctx->stream));
            CU(cudaMalloc(&ctx->sort_tmp, ctx->sort_tmp_bytes));
        }
        k_morton_keys<<<(ctx->S + 255) / 256, 256, 0, ctx->stream>>>((const float4*)ctx->aos_dev, ctx->S, ctx->N, ctx->P.origin[0], ctx->P.origin[1], ctx->P.origin[2],
                                                                     1.0f / ctx->cfg.grid_extent, ctx->sort_keys, ctx->sort_ids);
        CU(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp, ctx->sort_tmp_bytes, ctx->sort_keys, ctx->sort_keys_out, ctx->sort_ids, ctx->perm, ctx->S, 0, 30, ctx->stream));
    }
    const int tiles = ctx->S_pad / kTile;
    const size_t sm = (size_t)6 * ctx->N * (kTile + 1) * sizeof(float);
    CU(cudaFuncSetAttribute(k_unpack_aos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     // per function and device, not per context: set per launch
    k_unpack_aos<<<tiles, 256, sm, ctx->stream>>>((const float4*)ctx->aos_dev, ctx->planes, reorder ? ctx->perm : nullptr, ctx->S, ctx->S_pad, ctx->N, ctx->P.rest, 0);
    CU(cudaGetLastError());
    ctx->launches += reorder ? 3 : 1;
    ctx->uploaded = true;
    ctx->needs_resort = false;
    ctx->perm_active = reorder;
    ctx->state_version += 1;
    ctx->gather_pending = false;          // new state: nothing of the old grid applies to it
    return RVH_OK;
}

static int pack_to_staging(rvh_ctx* ctx) {
    { int r = flush_gather(ctx); if (r) return r; }
    const int tiles = (ctx->S + kTile - 1) / kTile;
    const size_t sm = (size_t)9 * ctx->N * (kTile + 1) * sizeof(float);
    const bool reorder = ctx->perm_active;
    CU(cudaFuncSetAttribute(k_pack_aos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_pack_aos<<<tiles, 256, sm, ctx->stream>>>((float4*)ctx->aos_dev, ctx->planes, ctx->corr, reorder ? ctx->perm : nullptr, ctx->S, ctx->S_pad, ctx->N, 0, 7);
    CU(cudaGetLastError());
    ctx->launches += 1;
    return RVH_OK;
}

int rvh_upload_strands_aos(rvh_ctx* ctx, const void* strands, size_t bytes) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!strands || bytes != ctx->aos_bytes) return fail(ctx, RVH_ERR_INVALID, "strands must be S*48*N bytes");
    CU(cudaSetDevice(ctx->cfg.device));
    // correctionVecs are dead on input (compute.comp:201 writes them before :214 reads them): only curvePoints and
    // curveVels (the first 32*N bytes of every 48*N-byte Strand) cross PCIe
    const size_t pitch = (size_t)48 * ctx->N, width = (size_t)32 * ctx->N;
    CU(cudaMemcpy2DAsync(ctx->aos_dev, pitch, strands, pitch, width, ctx->S, cudaMemcpyHostToDevice, ctx->stream));
    return unpack_from_staging(ctx);
}

int rvh_init_synthetic_head(rvh_ctx* ctx, unsigned long long first_strand, float strand_length, unsigned long long seed) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!ctx->colliders_set || ctx->P.n_ell < 1) return fail(ctx, RVH_ERR_STATE, "rvh_init_synthetic_head needs the colliders (collider 1 = head ellipsoid) set first");
    if (!(strand_length > 0.f)) return fail(ctx, RVH_ERR_INVALID, "strand_length must be > 0");
    CU(cudaSetDevice(ctx->cfg.device));
    const float rest = strand_length / ((float)ctx->N - 1.0f);
    k_synth_head_aos<<<(ctx->S + 255) / 256, 256, 0, ctx->stream>>>((float4*)ctx->aos_dev, ctx->S, ctx->N, first_strand, seed, rest, ctx->P.ell[0]);
    CU(cudaGetLastError());
    ctx->launches += 1;
    return unpack_from_staging(ctx);
}

int rvh_init_from_mesh(rvh_ctx* ctx, const float* tri_pos, const float* tri_nrm, int ntris, unsigned long long first_strand,
                       float strand_length, unsigned long long seed) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!tri_pos || ntris < 1) return fail(ctx, RVH_ERR_INVALID, "mesh needs >= 1 triangle");
    if (!(strand_length > 0.f)) return fail(ctx, RVH_ERR_INVALID, "strand_length must be > 0");
    // area CDF in double, sequential sum, normalised, rounded to float (host twin: scenes.triangle_cdf)
    std::vector<double> acc(ntris);
    double total = 0.0;
    for (int t = 0; t < ntris; ++t) {
        const float* A = tri_pos + 9 * (size_t)t;
        const double e1[3] = { (double)A[3] - A[0], (double)A[4] - A[1], (double)A[5] - A[2] }, e2[3] = { (double)A[6] - A[0], (double)A[7] - A[1], (double)A[8] - A[2] };
        const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
        total += 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz);
        acc[t] = total;
    }
    if (!(total > 0.0)) return fail(ctx, RVH_ERR_INVALID, "mesh has zero area");
    std::vector<float> cdf(ntris);
    for (int t = 0; t < ntris; ++t) cdf[t] = (float)(acc[t] / total);
    cdf[ntris - 1] = 1.0f;
    CU(cudaSetDevice(ctx->cfg.device));
    struct DevBuf { float* p = nullptr; ~DevBuf() { cudaFree(p); } } bpos, bnrm, bcdf;     // released on every return path
    CU(cudaMalloc(&bpos.p, sizeof(float) * 9 * ntris));
    CU(cudaMalloc(&bcdf.p, sizeof(float) * ntris));
    if (tri_nrm) CU(cudaMalloc(&bnrm.p, sizeof(float) * 9 * ntris));
    float *dpos = bpos.p, *dnrm = bnrm.p, *dcdf = bcdf.p;
    cudaError_t e = cudaMemcpyAsync(dpos, tri_pos, sizeof(float) * 9 * ntris, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dcdf, cdf.data(), sizeof(float) * ntris, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && tri_nrm) e = cudaMemcpyAsync(dnrm, tri_nrm, sizeof(float) * 9 * ntris, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const float rest = strand_length / ((float)ctx->N - 1.0f);
        k_mesh_follicles_aos<<<(ctx->S + 255) / 256, 256, 0, ctx->stream>>>((float4*)ctx->aos_dev, ctx->S, ctx->N, first_strand, seed, rest, dpos, dnrm, dcdf, ntris);
        e = cudaGetLastError();
        ctx->launches += 1;
    }
    int r = e == cudaSuccess ? unpack_from_staging(ctx) : fail(ctx, RVH_ERR_CUDA, std::string("rvh_init_from_mesh: ") + cudaGetErrorString(e));
    cudaStreamSynchronize(ctx->stream);                      // the pageable host arrays and the temporaries are released now
    return r;
}

int rvh_download_strands_aos(rvh_ctx* ctx, void* strands, size_t bytes) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!strands || bytes != ctx->aos_bytes) return fail(ctx, RVH_ERR_INVALID, "strands must be S*48*N bytes");
    if (!ctx->uploaded) return fail(ctx, RVH_ERR_STATE, "nothing uploaded");
    CU(cudaSetDevice(ctx->cfg.device));
    int r = pack_to_staging(ctx);
    if (r) return r;
    CU(cudaMemcpyAsync(strands, ctx->aos_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return check_exchange_error(ctx);
}

// ---- head SDF (extension) ---------------------------------------------------------------------------
static int sdf_alloc(rvh_ctx* ctx, const int dim[3], const float origin[3], float cell) {
    if (!dim || !origin) return fail(ctx, RVH_ERR_INVALID, "null argument");
    if (dim[0] < 2 || dim[1] < 2 || dim[2] < 2 || (size_t)dim[0] * dim[1] * dim[2] > ((size_t)1 << 30)) return fail(ctx, RVH_ERR_INVALID, "SDF dims must be >= 2 per axis and <= 2^30 nodes");
    if (!(cell > 0.f)) return fail(ctx, RVH_ERR_INVALID, "SDF cell must be > 0");
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->sdf_dev); ctx->sdf_dev = nullptr; ctx->sdf_mode = 0;
    const int nxp = (dim[0] + 3) & ~3;                               // 16-byte rows (TMA stride rule)
    CU(cudaMalloc(&ctx->sdf_dev, (size_t)nxp * dim[1] * dim[2] * sizeof(float)));
    for (int k = 0; k < 3; ++k) { ctx->sdf_dim[k] = dim[k]; ctx->P.sdf.origin[k] = origin[k]; }
    ctx->sdf_nxp = nxp;
    SdfVolume& V = ctx->P.sdf;
    V.data = ctx->sdf_dev; V.nx = dim[0]; V.ny = dim[1]; V.nz = dim[2]; V.nxp = nxp; V.inv_cell = 1.0f / cell;
    ctx->sdf_cell = cell;
    // TMA descriptor: rank-3 tiled map over the TRUE extents (padding columns are outside the tensor), box 8x4x4 nodes
    ctx->sdf_mode = 1;
    std::memset(&ctx->sdf_map, 0, sizeof ctx->sdf_map);
    const char* env = std::getenv("RVH_SDF_TMA");
    const bool want_tma = ((ctx->cfg.flags & RVH_SDF_TMA) || (env && std::atoi(env))) && ctx->V <= 2 && dim[0] >= kSdfBoxX && dim[1] >= kSdfBoxY && dim[2] >= kSdfBoxZ;
    if (want_tma) {
        typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn && q == cudaDriverEntryPointSuccess) {
            const cuuint64_t gdim[3] = { (cuuint64_t)dim[0], (cuuint64_t)dim[1], (cuuint64_t)dim[2] };
            const cuuint64_t gstr[2] = { (cuuint64_t)nxp * 4, (cuuint64_t)nxp * dim[1] * 4 };
            const cuuint32_t box[3] = { kSdfBoxX, kSdfBoxY, kSdfBoxZ }, estr[3] = { 1, 1, 1 };
            CUresult r = ((EncodeTiled)fn)(&ctx->sdf_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ctx->sdf_dev, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) ctx->sdf_mode = 2;
            else std::memset(&ctx->sdf_map, 0, sizeof ctx->sdf_map);
        } else cudaGetLastError();
    }
    return RVH_OK;
}

int rvh_set_head_sdf(rvh_ctx* ctx, const float* sdf, const int dim[3], const float origin[3], float cell) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!sdf) return fail(ctx, RVH_ERR_INVALID, "null SDF");
    int r = sdf_alloc(ctx, dim, origin, cell);
    if (r) return r;
    const int nx = dim[0], nxp = ctx->sdf_nxp;
    std::vector<float> padded((size_t)nxp * dim[1] * dim[2], kSdfFar);
    for (size_t row = 0; row < (size_t)dim[1] * dim[2]; ++row) std::memcpy(&padded[row * nxp], sdf + row * nx, sizeof(float) * nx);
    CU(cudaMemcpy(ctx->sdf_dev, padded.data(), padded.size() * sizeof(float), cudaMemcpyHostToDevice));
    return RVH_OK;
}

int rvh_bake_head_sdf_from_colliders(rvh_ctx* ctx, const int dim[3], const float origin[3], float cell) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!ctx->colliders_set || ctx->P.n_ell < 1) return fail(ctx, RVH_ERR_STATE, "rvh_bake_head_sdf_from_colliders needs ellipsoid colliders (rvh_set_colliders, n >= 2)");
    int r = sdf_alloc(ctx, dim, origin, cell);
    if (r) return r;
    const size_t total = (size_t)ctx->sdf_nxp * dim[1] * dim[2];
    k_sdf_bake_colliders<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ctx->P, ctx->sdf_dev, dim[0], dim[1], dim[2], ctx->sdf_nxp, origin[0], origin[1], origin[2], cell);
    CU(cudaGetLastError());
    ctx->launches += 1;
    CU(cudaStreamSynchronize(ctx->stream));
    return RVH_OK;
}

int rvh_bake_head_sdf_from_mesh(rvh_ctx* ctx, const float* verts, int nverts, const int* tris, int ntris,
                                const int dim[3], const float origin[3], float cell) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!verts || !tris || nverts < 3 || ntris < 1) return fail(ctx, RVH_ERR_INVALID, "mesh needs >= 3 vertices and >= 1 triangle");
    std::vector<float> t9((size_t)ntris * 9);
    for (int t = 0; t < ntris; ++t)
        for (int k = 0; k < 3; ++k) {
            const int v = tris[3 * t + k];
            if (v < 0 || v >= nverts) return fail(ctx, RVH_ERR_INVALID, "triangle index out of range");
            for (int a = 0; a < 3; ++a) t9[(size_t)t * 9 + 3 * k + a] = verts[3 * (size_t)v + a];
        }
    int r = sdf_alloc(ctx, dim, origin, cell);
    if (r) return r;
    cudaFree(ctx->bake_tris); ctx->bake_tris = nullptr;
    CU(cudaMalloc(&ctx->bake_tris, t9.size() * sizeof(float)));
    CU(cudaMemcpy(ctx->bake_tris, t9.data(), t9.size() * sizeof(float), cudaMemcpyHostToDevice));
    const size_t total = (size_t)ctx->sdf_nxp * dim[1] * dim[2];
    k_sdf_bake_mesh<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ctx->sdf_dev, dim[0], dim[1], dim[2], ctx->sdf_nxp, origin[0], origin[1], origin[2], cell, ctx->bake_tris, ntris);
    CU(cudaGetLastError());
    ctx->launches += 1;
    CU(cudaStreamSynchronize(ctx->stream));
    return RVH_OK;
}

int rvh_download_head_sdf(rvh_ctx* ctx, float* sdf, size_t bytes) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!ctx->sdf_dev) return fail(ctx, RVH_ERR_STATE, "no head SDF set");
    const size_t nx = ctx->sdf_dim[0], rows = (size_t)ctx->sdf_dim[1] * ctx->sdf_dim[2];
    if (!sdf || bytes != nx * rows * sizeof(float)) return fail(ctx, RVH_ERR_INVALID, "SDF download must be nx*ny*nz floats");
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy2D(sdf, nx * sizeof(float), ctx->sdf_dev, (size_t)ctx->sdf_nxp * sizeof(float), nx * sizeof(float), rows, cudaMemcpyDeviceToHost));
    return RVH_OK;
}

int rvh_debug_hit_masks(rvh_ctx* ctx, unsigned char* out, size_t bytes) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!ctx->uploaded || !ctx->colliders_set) return fail(ctx, RVH_ERR_STATE, "rvh_debug_hit_masks needs strands and colliders");
    if (!out || bytes != (size_t)ctx->S * ctx->N) return fail(ctx, RVH_ERR_INVALID, "hit masks are S*N bytes");
    if ((ctx->cfg.flags & RVH_SDF_ON) && !ctx->sdf_dev) return fail(ctx, RVH_ERR_STATE, "RVH_SDF_ON but no head SDF");
    CU(cudaSetDevice(ctx->cfg.device));
    unsigned char* dev = nullptr;
    CU(cudaMalloc(&dev, bytes));
    const bool reorder = ctx->perm_active;
    k_hit_masks<<<148 * 8, 256, 0, ctx->stream>>>(ctx->P, ctx->planes, reorder ? ctx->perm : nullptr, dev, (ctx->cfg.flags & RVH_SDF_ON) ? 1 : 0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dev);
    ctx->launches += 1;
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, RVH_ERR_CUDA, std::string("rvh_debug_hit_masks: ") + cudaGetErrorString(e)); }
    return RVH_OK;
}

int rvh_sdf_mode(rvh_ctx* ctx) { return ctx ? ctx->sdf_mode : 0; }

int rvh_download_collider_mask(rvh_ctx* ctx, unsigned char* out, size_t bytes, int* dim) {
    if (!ctx) return RVH_ERR_INVALID;
    const int D = ctx->P.cmask ? ctx->P.cmask_dim : 0;
    if (dim) *dim = D;
    if (!D) return RVH_OK;                                     // no mask in use (grid off, SDF on, odd grid_dim, no ellipsoids)
    if (!out || bytes != (size_t)D * D * D) return fail(ctx, RVH_ERR_INVALID, "collider mask is (grid_dim/2)^3 bytes");
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaMemcpyAsync(out, ctx->cmask_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return RVH_OK;
}


// ---- guide strand -> render strands (hair.tesc / hair.tese) -------------------------------------------------
static float mixh(float a, float b, float t) { return a + t * (b - a); }

// Everything of hair.tese that depends on the isoline (u) or the division (v) alone; see ExpandTables.
static void build_expand_tables(int N, int I, int D, ExpandTables& T) {
    std::memset(&T, 0, sizeof T);
    for (int j = 0; j <= D; ++j) {
        const float v = (float)j / (float)D;                                   // gl_TessCoord.x of an isoline with D segments
        const float vs = v * (float)(N - 1);                                   // hair.tese:165
        int seg = (int)std::floor(vs);
        if (seg > N - 2) seg = N - 2;                                          // v = 1: the shader would index point N
        T.seg[j] = seg;
        T.t[j] = vs - (float)seg;                                              // :216
        T.width[j] = 0.5f * mixh(mixh(0.05f, 0.3f, v), mixh(0.3f, 0.1f, v), v);   // :228-229 (clumpRadius = 0.5)
        T.strand_width[j] = mixh(0.02f, 0.01f, v);                             // :313-315
        const float two_s2 = 2.0f * std::pow(0.2f, 2.0f);
        T.sd[0][j] = 1.0f;
        T.sd[1][j] = 1.8f * std::exp(-std::pow(v - 0.25f, 2.0f) / two_s2);     // :248
        T.sd[2][j] = 4.5f * std::pow(v, 10.0f);                                // :250
        T.sd[3][j] = 2.5f * std::exp(-std::pow(v - 0.7f, 2.0f) / two_s2);      // :252
        T.sd[4][j] = 4.0f * std::pow(v, 1.3f);                                 // :254
        T.sd[5][j] = 1.8f * std::exp(-std::pow(v - 0.8f, 2.0f) / two_s2);      // :256
        if (v == 0.0f) for (int f = 0; f < 6; ++f) T.sd[f][j] = 1.0f;          // :267-269
    }
    for (int k = 0; k < I; ++k) {
        const float u = (float)k / (float)I;                                   // gl_TessCoord.y: isoline k of I
        T.u[k] = u;
        const float uRad = 2.0f * 3.141592653f * u;                            // :235
        const float c = std::cos(uRad), sn = std::sin(uRad);
        const float len = std::sqrt(c * c + sn * sn);
        T.dirx[k] = c / len; T.dirz[k] = sn / len;                             // :236
        T.wr[k] = std::fabs(expand_hash(u, u * u)) + 0.5f;                     // rand2 + 0.5, :226,229
    }
}

int rvh_expand_strands(rvh_ctx* ctx, int isolines, int divisions, float* pos_width, float* tangent_u, size_t bytes_each, float* ms_out) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!ctx->uploaded) return fail(ctx, RVH_ERR_STATE, "rvh_expand_strands before any strands were uploaded");
    if (isolines < 1 || isolines > kMaxIsolines || divisions < 1 || divisions > kMaxDivisions) return fail(ctx, RVH_ERR_INVALID, "1 <= isolines <= 64, 1 <= divisions <= 256");
    const size_t verts = (size_t)ctx->S * isolines * (divisions + 1);
    if ((pos_width || tangent_u) && bytes_each != verts * 16) return fail(ctx, RVH_ERR_INVALID, "each output must be S*isolines*(divisions+1)*16 bytes");
    CU(cudaSetDevice(ctx->cfg.device));
    if (!ctx->exp_tab || ctx->exp_I != isolines || ctx->exp_D != divisions) {
        ExpandTables T;
        build_expand_tables(ctx->N, isolines, divisions, T);
        if (!ctx->exp_tab) CU(cudaMalloc(&ctx->exp_tab, sizeof(ExpandTables)));
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaMemcpy(ctx->exp_tab, &T, sizeof T, cudaMemcpyHostToDevice));
        ctx->exp_I = isolines; ctx->exp_D = divisions;
    }
    if (ctx->exp_cap < verts) {
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->exp_pw); cudaFree(ctx->exp_tu); ctx->exp_pw = ctx->exp_tu = nullptr; ctx->exp_cap = 0;
        CU(cudaMalloc(&ctx->exp_pw, verts * sizeof(float4)));
        CU(cudaMalloc(&ctx->exp_tu, verts * sizeof(float4)));
        ctx->exp_cap = verts;
    }
    const size_t sm = ((size_t)3 * ctx->N * (kExpandTile + 1) + 10 * (divisions + 1) + 4 * isolines + (size_t)isolines * (divisions + 1) + kExpandTile) * sizeof(float) + (size_t)kExpandTile * isolines;
    CU(cudaFuncSetAttribute(k_expand_strands, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const bool reorder = ctx->perm_active;
    if (ms_out) CU(cudaEventRecord(ctx->ev_a, ctx->stream));
    k_expand_strands<<<(ctx->S + kExpandTile - 1) / kExpandTile, 256, sm, ctx->stream>>>(ctx->planes, reorder ? ctx->perm : nullptr, ctx->exp_tab, ctx->exp_pw, ctx->exp_tu,
                                                                                          ctx->S, ctx->S_pad, ctx->N, isolines, divisions);
    CU(cudaGetLastError());
    ctx->launches += 1;
    if (ms_out) {
        CU(cudaEventRecord(ctx->ev_b, ctx->stream));
        CU(cudaEventSynchronize(ctx->ev_b));
        CU(cudaEventElapsedTime(ms_out, ctx->ev_a, ctx->ev_b));
    }
    if (pos_width) CU(cudaMemcpyAsync(pos_width, ctx->exp_pw, verts * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (tangent_u) CU(cudaMemcpyAsync(tangent_u, ctx->exp_tu, verts * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (pos_width || tangent_u) CU(cudaStreamSynchronize(ctx->stream));
    return RVH_OK;
}

int rvh_expand_device_buffers(rvh_ctx* ctx, void** pos_width, void** tangent_u, size_t* vertices) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!ctx->exp_pw) return fail(ctx, RVH_ERR_STATE, "rvh_expand_strands has not run");
    if (pos_width) *pos_width = ctx->exp_pw;
    if (tangent_u) *tangent_u = ctx->exp_tu;
    if (vertices) *vertices = (size_t)ctx->S * ctx->exp_I * (ctx->exp_D + 1);
    return RVH_OK;
}

int rvh_import_strands_fd(rvh_ctx* ctx, int fd, size_t bytes) {
    if (!ctx) return RVH_ERR_INVALID;
    if (bytes < ctx->aos_bytes) return fail(ctx, RVH_ERR_INVALID, "imported buffer smaller than Strand[S]");
    CU(cudaSetDevice(ctx->cfg.device));
    if (ctx->interop_mem) {                                         // re-import: release the previous mapping first
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->interop_aos) cudaFree(ctx->interop_aos);
        cudaDestroyExternalMemory(ctx->interop_mem);
        ctx->interop_aos = nullptr; ctx->interop_mem = nullptr;
    }
    cudaExternalMemoryHandleDesc hd; std::memset(&hd, 0, sizeof hd);
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd; hd.handle.fd = fd; hd.size = bytes;
    CU(cudaImportExternalMemory(&ctx->interop_mem, &hd));
    cudaExternalMemoryBufferDesc bd; std::memset(&bd, 0, sizeof bd);
    bd.offset = 0; bd.size = bytes;
    CU(cudaExternalMemoryGetMappedBuffer(&ctx->interop_aos, ctx->interop_mem, &bd));
    return RVH_OK;
}

int rvh_import_indirect_fd(rvh_ctx* ctx, int fd, size_t bytes) {
    if (!ctx) return RVH_ERR_INVALID;
    if (bytes < 16) return fail(ctx, RVH_ERR_INVALID, "imported buffer smaller than StrandDrawIndirect (16 bytes)");
    CU(cudaSetDevice(ctx->cfg.device));
    if (ctx->interop_indirect_mem) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->interop_indirect) cudaFree(ctx->interop_indirect);
        cudaDestroyExternalMemory(ctx->interop_indirect_mem);
        ctx->interop_indirect = nullptr; ctx->interop_indirect_mem = nullptr;
    }
    cudaExternalMemoryHandleDesc hd; std::memset(&hd, 0, sizeof hd);
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd; hd.handle.fd = fd; hd.size = bytes;
    CU(cudaImportExternalMemory(&ctx->interop_indirect_mem, &hd));
    cudaExternalMemoryBufferDesc bd; std::memset(&bd, 0, sizeof bd);
    bd.offset = 0; bd.size = bytes;
    void* p = nullptr;
    CU(cudaExternalMemoryGetMappedBuffer(&p, ctx->interop_indirect_mem, &bd));
    ctx->interop_indirect = (uint32_t*)p;
    return RVH_OK;
}

int rvh_import_semaphore_fd(rvh_ctx* ctx, int fd) {
    if (!ctx) return RVH_ERR_INVALID;
    CU(cudaSetDevice(ctx->cfg.device));
    if (ctx->interop_sem) { CU(cudaStreamSynchronize(ctx->stream)); cudaDestroyExternalSemaphore(ctx->interop_sem); ctx->interop_sem = nullptr; }
    cudaExternalSemaphoreHandleDesc sd; std::memset(&sd, 0, sizeof sd);
    sd.type = cudaExternalSemaphoreHandleTypeOpaqueFd; sd.handle.fd = fd;
    CU(cudaImportExternalSemaphore(&ctx->interop_sem, &sd));
    return RVH_OK;
}

int rvh_debug_set_interop_device_buffers(rvh_ctx* ctx, void* strands_dev, size_t strands_bytes, void* indirect_dev) {
    if (!ctx) return RVH_ERR_INVALID;
    if (ctx->interop_mem || ctx->interop_indirect_mem) return fail(ctx, RVH_ERR_STATE, "imported Vulkan buffers are in use");
    if (strands_dev && strands_bytes < ctx->aos_bytes) return fail(ctx, RVH_ERR_INVALID, "strands buffer smaller than Strand[S]");
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->interop_aos = strands_dev;                                  // caller-owned: never freed here (interop_mem stays null)
    ctx->interop_indirect = (uint32_t*)indirect_dev;
    return RVH_OK;
}

int rvh_step(rvh_ctx* ctx, float dt, float total_time) {
    if (!ctx) return RVH_ERR_INVALID;
    CU(cudaSetDevice(ctx->cfg.device));
    if (dt > 0.f && scene_step_eligible(ctx) && !std::getenv("RVH_NO_FAST_STEP_N")) {   // small scene in its steady state: one launch per step
        float t = total_time;
        const int r = launch_scene_steps(ctx, 1, dt, t);
        if (r != kSceneFallback) return r;
    }
    return do_step(ctx, dt, total_time, 3, true);
}

int rvh_step_phases(rvh_ctx* ctx, float dt, float total_time, int phases) {
    if (!ctx) return RVH_ERR_INVALID;
    CU(cudaSetDevice(ctx->cfg.device));
    return do_step(ctx, dt, total_time, phases, false);
}

int rvh_step_n(rvh_ctx* ctx, int n, float dt, float total_time0, float* ms_out) {
    if (!ctx) return RVH_ERR_INVALID;
    if (n < 1) return fail(ctx, RVH_ERR_INVALID, "n must be >= 1");
    CU(cudaSetDevice(ctx->cfg.device));
    if (ms_out) CU(cudaEventRecord(ctx->ev_a, ctx->stream));
    const int flags = ctx->cfg.flags;
    const bool grid = flags & RVH_GRID_ON, wind = flags & (RVH_WIND_A | RVH_WIND_B);
    const bool plain = ctx->uploaded && ctx->colliders_set && dt > 0.f && ctx->profiling == 0 && !ctx->interop_aos && !(flags & RVH_SDF_ON) &&
                       ctx->nranks == 1 && !std::getenv("RVH_NO_FAST_STEP_N");
    float t = total_time0;
    int done = 0;
    if (plain && !grid && n > 1) {
        // grid off: the strands never interact, several steps ride in one launch (k_ftl_step<..., MULTI>)
        while (done < n) {
            const int m = std::min(32, n - done);
            int r = launch_multi_step(ctx, m, dt, t);
            if (r) return r;
            done += m;
        }
    } else if (plain && grid && n >= 2 && ctx->scene_ctas_per_sm > 0 && ctx->k1_blocks * 2 <= ctx->num_sms && !(flags & RVH_REPULSION_ON) &&
               !ctx->interop_indirect && !ctx->interop_sem) {
        // small scene with the grid on: whole steps (up to 32 of them) inside ONE persistent cooperative launch (k_scene_step)
        if (!scene_step_eligible(ctx)) {                               // a pipelined host step left the caller's strand order: one ordinary step restores the Morton order
            int r = do_step(ctx, dt, t, 3, true);
            if (r) return r;
            done = 1; t += dt;
        }
        while (done < n && scene_step_eligible(ctx)) {
            const int m = std::min(32, n - done);
            int r = launch_scene_steps(ctx, m, dt, t);
            if (r == kSceneFallback) break;                             // the plain loop below takes the remaining steps
            if (r) return r;
            done += m;
        }
    } else if (plain && grid && !wind && n >= 4 && (size_t)ctx->S_pad * ctx->N <= ((size_t)1 << 23)) {
        // small scene with the grid on: the step's 3-4 launches are launch/latency-bound, replay them as one CUDA graph.  Without
        // wind nothing in the step's kernel parameters depends on the time, so one captured step serves every later step.
        int r = do_step(ctx, dt, t, 3, true);                       // reaches the steady state (gather pending) outside the graph
        if (r) return r;
        done = 1; t += dt;
        if (!ctx->step_graph || ctx->step_graph_dt != dt || ctx->step_graph_version != ctx->state_version) {
            if (ctx->step_graph) { cudaGraphExecDestroy(ctx->step_graph); ctx->step_graph = nullptr; }
            cudaGraph_t g = nullptr;
            const long long launches0 = ctx->launches;
            CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            r = do_step(ctx, dt, t, 3, true);
            cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
            const int captured = (int)(ctx->launches - launches0);
            ctx->launches = launches0;                              // captured, not launched
            if (r) { if (g) cudaGraphDestroy(g); return r; }
            if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, RVH_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); }
            e = cudaGraphInstantiate(&ctx->step_graph, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) { cudaGetLastError(); ctx->step_graph = nullptr; return fail(ctx, RVH_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
            ctx->step_graph_dt = dt; ctx->step_graph_version = ctx->state_version;
            ctx->step_graph_launches = captured;
        }
        for (; done < n; ++done, t += dt) {
            CU(cudaGraphLaunch(ctx->step_graph, ctx->stream));
            ctx->launches += ctx->step_graph_launches;
        }
        ctx->gather_pending = true;
    }
    for (; done < n; ++done) {
        int r = do_step(ctx, dt, t, 3, true);
        if (r) return r;
        t += dt;
    }
    if (ms_out) {
        CU(cudaEventRecord(ctx->ev_b, ctx->stream));
        CU(cudaEventSynchronize(ctx->ev_b));
        CU(cudaEventElapsedTime(&ctx->last_ms, ctx->ev_a, ctx->ev_b));
        *ms_out = ctx->last_ms;
        ctx->last_ms /= (float)n;
        prof_collect(ctx);
    }
    return RVH_OK;
}

// Host round trip, pipelined over strand chunks.  PCIe is the bound (2 x 32*N bytes per strand), so the copies of different
// chunks overlap each other and the kernels: H2D(chunk c+1) runs beside unpack + k_ftl_step + splat of chunk c, and because
// the gather only changes velocities, the POSITIONS of chunk c go back to the host right after its k_ftl_step, in the other
// PCIe direction, while later chunks are still arriving.  Only the velocities wait for the complete grid.  The chunks are
// taken in the caller's order (no Morton reordering: the splat is slower on unsorted strands, but it hides under the copies);
// integer grid sums are order-independent, so the result is bit-identical to upload + step + download.
static int step_host_pipelined(rvh_ctx* ctx, void* strands, float dt, float total_time) {
    const int flags = ctx->cfg.flags;
    const bool grid = flags & RVH_GRID_ON, wind = flags & (RVH_WIND_A | RVH_WIND_B);
    StepParams& P = ctx->P;
    P.dt = dt; P.inv_dt = 1.0f / dt; P.dt2 = dt * dt; P.vel_scale = P.damping / dt;
    P.wind_mode = (flags & RVH_WIND_B) ? 2 : ((flags & RVH_WIND_A) ? 1 : 0);
    if (wind) wind_scalars(P.wind_mode, total_time, P.wind_s2T, P.wind_T3, P.wind_amp);
    const int per_cta = kBlock * ctx->V;                                  // strands per k_ftl_step CTA: 128 or 256
    int nchunks = 16;
    if (const char* e = std::getenv("RVH_HOST_CHUNKS")) nchunks = std::max(1, std::atoi(e));
    const int chunk = std::max(256, ((ctx->S_pad + nchunks - 1) / nchunks + 255) / 256 * 256);
    nchunks = (ctx->S_pad + chunk - 1) / chunk;
    if (!ctx->h2d_stream) { CU(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking)); }
    while ((int)ctx->pipe_ev.size() < 3 * nchunks + 1) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->pipe_ev.push_back(e); }
    const size_t pitch = (size_t)48 * ctx->N, third = (size_t)16 * ctx->N;
    char* host = (char*)strands; char* dev = (char*)ctx->aos_dev;
    const size_t sm_un = (size_t)6 * ctx->N * (kTile + 1) * sizeof(float), sm_pk = (size_t)9 * ctx->N * (kTile + 1) * sizeof(float);
    CU(cudaFuncSetAttribute(k_unpack_aos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_un));
    CU(cudaFuncSetAttribute(k_pack_aos, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pk));
    // the copy streams start after whatever the context's stream still has in flight
    CU(cudaEventRecord(ctx->pipe_ev[3 * nchunks], ctx->stream));
    CU(cudaStreamWaitEvent(ctx->h2d_stream, ctx->pipe_ev[3 * nchunks], 0));
    CU(cudaStreamWaitEvent(ctx->d2h_stream, ctx->pipe_ev[3 * nchunks], 0));
    if (grid) CU(cudaMemsetAsync(ctx->grid, 0, ctx->grid_bytes, ctx->stream));      // Renderer.cpp:2063
    ctx->k1_clear = nullptr; ctx->k1_clear_n = 0;
    ctx->perm_active = false; ctx->gather_pending = false; ctx->uploaded = true; ctx->state_version += 1;
    ctx->needs_resort = !(flags & RVH_KEEP_ORDER) && ctx->S >= 1024;
    for (int c = 0; c < nchunks; ++c) {
        const int s0 = c * chunk, ns = std::min(chunk, ctx->S_pad - s0), ne = std::max(0, std::min(ns, ctx->S - s0));
        cudaEvent_t ev_in = ctx->pipe_ev[3 * c], ev_pos = ctx->pipe_ev[3 * c + 1];
        if (ne > 0) CU(cudaMemcpy2DAsync(dev + (size_t)s0 * pitch, pitch, host + (size_t)s0 * pitch, pitch, 2 * third, ne, cudaMemcpyHostToDevice, ctx->h2d_stream));
        CU(cudaEventRecord(ev_in, ctx->h2d_stream));
        CU(cudaStreamWaitEvent(ctx->stream, ev_in, 0));
        k_unpack_aos<<<ns / kTile, 256, sm_un, ctx->stream>>>((const float4*)ctx->aos_dev, ctx->planes, nullptr, ctx->S, ctx->S_pad, ctx->N, ctx->P.rest, s0 / kTile);
        launch_ftl(ctx, wind, 0, s0 / per_cta, (ns + per_cta - 1) / per_cta);
        if (grid) launch_splat(ctx, s0, ns);
        if (ne > 0) k_pack_aos<<<(ne + kTile - 1) / kTile, 256, sm_pk, ctx->stream>>>((float4*)ctx->aos_dev, ctx->planes, ctx->corr, nullptr, ctx->S, ctx->S_pad, ctx->N, s0 / kTile, grid ? 1 : 3);
        ctx->launches += 2;
        CU(cudaGetLastError());
        CU(cudaEventRecord(ev_pos, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->d2h_stream, ev_pos, 0));
        if (ne > 0) CU(cudaMemcpy2DAsync(host + (size_t)s0 * pitch, pitch, dev + (size_t)s0 * pitch, pitch, grid ? third : 2 * third, ne, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    }
    if (grid) {
        if (ctx->nranks > 1) {                                             // every rank calls rvh_step_host in lockstep, like rvh_step
            ctx->grid_reduced = false;
            if (!ctx->p2p) { int r = allreduce_grid(ctx); if (r) return r; }
        }
        { int r = finalize_grid(ctx); if (r) return r; }
        { int r = launch_gather(ctx); if (r) return r; }
        for (int c = 0; c < nchunks; ++c) {
            const int s0 = c * chunk, ns = std::min(chunk, ctx->S_pad - s0), ne = std::max(0, std::min(ns, ctx->S - s0));
            if (ne <= 0) break;
            cudaEvent_t ev_vel = ctx->pipe_ev[3 * c + 2];
            k_pack_aos<<<(ne + kTile - 1) / kTile, 256, sm_pk, ctx->stream>>>((float4*)ctx->aos_dev, ctx->planes, ctx->corr, nullptr, ctx->S, ctx->S_pad, ctx->N, s0 / kTile, 2);
            ctx->launches += 1;
            CU(cudaGetLastError());
            CU(cudaEventRecord(ev_vel, ctx->stream));
            CU(cudaStreamWaitEvent(ctx->d2h_stream, ev_vel, 0));
            CU(cudaMemcpy2DAsync(host + (size_t)s0 * pitch + third, pitch, dev + (size_t)s0 * pitch + third, pitch, third, ne, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        }
    }
    CU(cudaStreamSynchronize(ctx->d2h_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return check_exchange_error(ctx);
}

int rvh_step_host(rvh_ctx* ctx, void* strands, size_t bytes, float dt, float total_time) {
    if (!ctx) return RVH_ERR_INVALID;
    if (!strands || bytes != ctx->aos_bytes) return fail(ctx, RVH_ERR_INVALID, "strands must be S*48*N bytes");
    if (ctx->colliders_set && dt > 0.f && !ctx->corr && !ctx->interop_aos && ctx->S >= 131072 && !(ctx->cfg.flags & RVH_SDF_TMA) &&
        !((ctx->cfg.flags & RVH_SDF_ON) && !ctx->sdf_dev) && !std::getenv("RVH_NO_HOST_PIPE")) {
        CU(cudaSetDevice(ctx->cfg.device));
        return step_host_pipelined(ctx, strands, dt, total_time);
    }
    int r = rvh_upload_strands_aos(ctx, strands, bytes);
    if (r) return r;
    r = do_step(ctx, dt, total_time, 3, true);
    if (r) return r;
    if (ctx->corr) return rvh_download_strands_aos(ctx, strands, bytes);
    // without RVH_KEEP_CORRECTION the correctionVecs third is not produced: bring back curvePoints + curveVels only
    r = pack_to_staging(ctx);
    if (r) return r;
    const size_t pitch = (size_t)48 * ctx->N, width = (size_t)32 * ctx->N;
    CU(cudaMemcpy2DAsync(strands, pitch, ctx->aos_dev, pitch, width, ctx->S, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return RVH_OK;
}

int rvh_download_grid(rvh_ctx* ctx, void* cells, size_t bytes) {
    if (!ctx) return RVH_ERR_INVALID;
    const bool wrap = ctx->cfg.flags & RVH_GRID_INT32_WRAP;
    const size_t want = wrap ? ctx->grid_bytes / 2 : ctx->grid_bytes;
    if (!cells || bytes != want) return fail(ctx, RVH_ERR_INVALID, "grid download size mismatch");
    CU(cudaSetDevice(ctx->cfg.device));
    if (ctx->nranks > 1) { int r = allreduce_grid(ctx); if (r) return r; }   // peer-exchange mode keeps raw per-rank accumulators (collective call)
    if (!wrap) {
        CU(cudaMemcpyAsync(cells, ctx->grid, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    } else {
        std::vector<long long> tmp(ctx->grid_bytes / 8);
        CU(cudaMemcpyAsync(tmp.data(), ctx->grid, ctx->grid_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        int32_t* o = (int32_t*)cells;
        for (size_t i = 0; i < tmp.size(); ++i) o[i] = (int32_t)(uint32_t)(uint64_t)tmp[i];
    }
    return RVH_OK;
}

int rvh_draw_indirect(rvh_ctx* ctx, uint32_t out[4]) {
    if (!ctx || !out) return RVH_ERR_INVALID;
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaStreamSynchronize(ctx->stream));
    // compute.comp:126-130,302 count invocations; the reference leaves 32*ceil(S/32) here because
    // the shader has no bounds guard.  This path reports the strand count Hair::Hair set (Strand.cpp:178-182).
    out[0] = (uint32_t)ctx->S; out[1] = 1; out[2] = 0; out[3] = 0;
    return RVH_OK;
}

int rvh_profile_enable(rvh_ctx* ctx, int on) {
    if (!ctx) return RVH_ERR_INVALID;
    ctx->profiling = on < 0 ? 0 : (on > 2 ? 1 : on);
    return RVH_OK;
}

int rvh_profile_read(rvh_ctx* ctx, float ms[6], int launches[6]) {
    if (!ctx) return RVH_ERR_INVALID;
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaStreamSynchronize(ctx->stream));
    prof_collect(ctx);
    for (int i = 0; i < EV_COUNT; ++i) {
        if (ms) ms[i] = ctx->prof_ms[i];
        if (launches) launches[i] = ctx->prof_n[i];
        ctx->prof_ms[i] = 0.f; ctx->prof_n[i] = 0;
    }
    return RVH_OK;
}

int rvh_sync(rvh_ctx* ctx) {
    if (!ctx) return RVH_ERR_INVALID;
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaStreamSynchronize(ctx->stream));
    return check_exchange_error(ctx);
}

float rvh_last_step_ms(rvh_ctx* ctx) { return ctx ? ctx->last_ms : 0.f; }
long long rvh_kernel_launches(rvh_ctx* ctx) { return ctx ? ctx->launches : 0; }
const char* rvh_last_error(rvh_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void rvh_destroy(rvh_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void* q : c->ipc_open) if (q) cudaIpcCloseMemHandle(q);
    cudaFree(c->xflags);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    if (c->interop_mem) { if (c->interop_aos) cudaFree(c->interop_aos); cudaDestroyExternalMemory(c->interop_mem); }
    if (c->interop_indirect_mem) { if (c->interop_indirect) cudaFree(c->interop_indirect); cudaDestroyExternalMemory(c->interop_indirect_mem); }
    if (c->interop_sem) cudaDestroyExternalSemaphore(c->interop_sem);
    cudaFree(c->cmask_dev);
    cudaFree(c->sdf_dev); cudaFree(c->bake_tris); cudaFree(c->exp_tab); cudaFree(c->exp_pw); cudaFree(c->exp_tu);
    cudaFree(c->scene_bar); cudaFree(c->scene_timing);
    cudaFree(c->planes); cudaFree(c->corr); cudaFree(c->grid); cudaFree(c->fgrid); cudaFree(c->perm); cudaFree(c->aos_dev);
    cudaFree(c->sort_tmp); cudaFree(c->sort_keys); cudaFree(c->sort_keys_out); cudaFree(c->sort_ids);
    if (c->step_graph) cudaGraphExecDestroy(c->step_graph);
    for (cudaEvent_t e : c->pipe_ev) cudaEventDestroy(e);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (cudaEvent_t e : c->pev) cudaEventDestroy(e);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void rvh_collider_build(const float t[3], const float r[3], const float s[3], float out48[48]) { collider_build(t, r, s, out48); }
void rvh_collider_translate(float c48[48], const float tr[3]) { collider_translate(c48, tr); }
float rvh_wind_fbm(float total_time) { return wind_fbm(total_time); }

}  // extern "C"
