// rvh_kernels.cuh -- sm_100a kernels of the guide-strand physics step.
//
// Device-side replacement for src/shaders/compute.comp (file:line below are relative to
// the reference tree).  Layout in HBM is point-major, tile-interleaved SoA ("planes"):
//     planes[i][t][k][j]   i = point on strand (row), t = tile of 256 strands, k in {px,py,pz,vx,vy,vz},
//                          j = strand within the tile
// so that a warp reading point i of 32*V consecutive strands issues fully coalesced 32/64/128-bit
// loads per plane, the six planes of a row sit at compile-time offsets (k KB) from ONE per-thread
// pointer that advances by a uniform stride per row (no per-plane address arithmetic), and the
// root->tip chain of a strand lives in registers.
// The voxel grid is int64 [G^3][4] (vx,vy,vz,density), fixed point x grid_scale; integer
// accumulation makes the result independent of atomics order and of the GPU count.
//
// Kernels per step (grid on):
//   k_ftl_step<V,WIND,NELL,GATHER,MULTI>  [grid clear in the prologue on wide launches;] gather + friction (+ repulsion) of the PREVIOUS
//                                   step's grid fused into the load (GATHER), collider tests (unrolled ellipsoids, optionally
//                                   behind the collider candidate mask, or the head SDF by plain loads / TMA-staged tiles),
//                                   then integrate + FTL + corrected velocity.  MULTI (grid off): up to 32 whole steps per launch
//   k_grid_splat<MAGIC>             two-phase register-accumulating integer splat, one RED.64 per lane and cell (sparse rows: point by point)
//   k_ftl_wave<WIND,NELL,W>         grid off, small scene: wavefront over the steps (W lanes per strand, lane j = step s0+j, rows handed down by shuffles)
//   k_scene_step<V,WIND,NELL>       small scenes: whole steps in ONE persistent cooperative launch (FTL || clear | splat | per-point gather
//                                   from the raw accumulators, grid-wide barriers in between); rvh_step / rvh_step_n use it
//   k_grid_exchange                 sharded runs: pull-reduce + finalize + push over NVLink peer memory (or ncclAllReduce); bounded waits
//   k_grid_finalize                 int64 accumulators -> float4 cells (v/density, density) for the gather
//   k_grid_gather<REP>              stand-alone gather, only when state is read back before the next step
// Around the step: k_unpack_aos / k_pack_aos (Strand[S] AoS <-> planes, by tile ranges for the pipelined host round trip),
// k_write_indirect (imported draw arguments), k_morton_keys, k_synth_head_aos, k_mesh_follicles_aos (GPU scene init),
// k_sdf_bake_colliders / k_sdf_bake_mesh, k_expand_strands (hair.tesc/tese), k_hit_masks (test hook: collider decisions).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>            // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdint.h>

namespace rvh {

constexpr int kMaxEllipsoids = 7;       // colliders 1..7; collider 0 is the sphere
#ifndef RVH_K1_THREADS
#define RVH_K1_THREADS 128
#endif
constexpr int kBlock = RVH_K1_THREADS;
constexpr int kTileStrands = 256;       // strands per layout tile; S_pad is a multiple of it

// element index of (row, plane k of NP, strand s) in a tiled array [N][S_pad/256][NP][256]
__host__ __device__ __forceinline__ size_t tiled_index(int NP, int S_pad, int row, int k, int s) {
    return (((size_t)row * (S_pad / kTileStrands) + s / kTileStrands) * NP + k) * kTileStrands + (s % kTileStrands);
}

struct Ellipsoid {
    float inv[12];   // rows 0..2 of Collider::inv      : q = inv * (p,1)     compute.comp:64-67
    float xf[12];    // rows 0..2 of Collider::transform: on = xf * (nq,1)    compute.comp:76-79
    float nt[9];     // upper 3x3 of Collider::invTrans : n = nt * q          compute.comp:70-73
};

// Head SDF (north-star extension, include/rvh.h): node values [nz][ny][nxp], x fastest, nxp = nx rounded up to 4
// floats so that the rows satisfy TMA's 16-byte stride rule.
struct SdfVolume {
    const float* data;
    int nx, ny, nz, nxp;
    float origin[3], inv_cell;
};

struct StepParams {
    int S, S_pad, N;
    float rest, gravity_y, damping, vmax, vmax2, penalty_k;
    float sphere_r, sphere_r2, sphere_c[3];
    int has_sphere, n_ell;
    Ellipsoid ell[kMaxEllipsoids];
    int G;
    float h, rh, origin[3], scale, friction;   // rh = RN(1/h)
    int div_fast;                               // 1: h is a normal float whose significand is not all ones (see grid_coord)
    float dt, inv_dt, dt2, vel_scale;      // vel_scale = damping / dt
    int wind_mode;                          // 0 off, 1 = variant A (:151), 2 = variant B (:152)
    float wind_s2T, wind_T3, wind_amp;      // 2*sin(2T); 3T mod 2pi; 10 (A) or 7*fbm(sinT,cosT) (B)
    int int32_wrap, keep_corr;
    const unsigned char* cmask;             // collider candidate mask over the grid box at half resolution (see gather_pack); null = none
    int cmask_dim;                          // G / 2
    SdfVolume sdf;                          // RVH_SDF_ON
    float repulsion, inv_h;                 // RVH_REPULSION_ON: v -= repulsion * h*grad(rho)/sum(D) (gather_pack); inv_h = 1/h
    int cta0, strand0;                      // chunked launches (rvh_step_host's pipeline): first CTA of this k_ftl_step launch, first strand of this splat launch
    int multi_steps;                        // k_ftl_step<..., MULTI>: steps per launch (grid off: strands are independent); wind_tab = per-step (amp*s2T, T3, amp)
    float wind_tab[3 * 32];
    int splat_sparse;                       // k_grid_splat: rows whose 32 points fall into at least this many distinct base cells go to the grid point by point (0 = never)
    float splat_vagg;                       // k_grid_splat: points with |v|_inf <= splat_vagg are aggregated in int32 registers (32 contributions of <= 2^26 each)
};

// ---- wind trigonometry ---------------------------------------------------------------------------
// The wind force enters positions multiplied by dt^2 (2.8e-4 at 60 Hz), so MUFU.SIN/COS accuracy is far
// more than the 1e-4*L position tolerance needs: for the bounded arguments used here (|x| < ~200; the host
// reduces the time term mod 2*pi) the absolute error is < 2e-5, i.e. < 1e-8 in position.  2 instructions
// each instead of a 14-instruction Cody-Waite + polynomial evaluation.
__device__ __forceinline__ float sin_bounded(float x) { return __sinf(x); }
__device__ __forceinline__ float cos_bounded(float x) { return __cosf(x); }
__device__ __forceinline__ float rsqrt_fast(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2) ----------------------------------
// sm_100 issues two fp32 FMAs per lane in ONE instruction (fma.rn.f32x2 -> SASS FFMA2) and accepts a
// scalar register broadcast to both halves.  The step is issue-bound, not pipe-bound, so pairing the
// two strands a thread owns halves the issue slots of all straight-line arithmetic.  The code below is
// written once over T = float (one strand per thread) or T = float2 (two strands per pack).
template <class T> struct VecTraits;
template <> struct VecTraits<float>  { static constexpr int n = 1; };
template <> struct VecTraits<float2> { static constexpr int n = 2; };
template <class T> __device__ __forceinline__ T bc(float a);
template <> __device__ __forceinline__ float  bc<float>(float a)  { return a; }
template <> __device__ __forceinline__ float2 bc<float2>(float a) { return make_float2(a, a); }
__device__ __forceinline__ float  vfma(float a, float b, float c)    { return fmaf(a, b, c); }
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float  vmul(float a, float b)   { return a * b; }
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float  vadd(float a, float b)   { return a + b; }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float  vsub(float a, float b)   { return a - b; }
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }   // a - b, one rounding
__device__ __forceinline__ float  el(float a, int)    { return a; }
__device__ __forceinline__ float  el(float2 a, int i) { return i ? a.y : a.x; }
__device__ __forceinline__ void setel(float& a, int, float v)    { a = v; }
__device__ __forceinline__ void setel(float2& a, int i, float v) { if (i) a.y = v; else a.x = v; }
template <class T> __device__ __forceinline__ T vdot3(T ax, T ay, T az, T bx, T by, T bz) { return vfma(ax, bx, vfma(ay, by, vmul(az, bz))); }

// ---- grid coordinates of a point: compute.comp:219-229, 237-239 ---------------------------------
// g = (p - origin) / h must be the IEEE quotient: the integer grid is compared bit for bit with the
// oracle, and floor(g) decides the cell.  x / h is evaluated as Markstein's sequence around the
// host-rounded reciprocal rh = RN(1/h):
//     q0 = x*rh;  q1 = q0 + (x - h*q0)*rh;  q2 = q1 + (x - h*q1)*rh
// q1 is within one ulp, so q2 = RN(x/h) whenever the significand of h is not all ones and nothing
// under/overflows (checked EXHAUSTIVELY over all 2^32 floats x against x/h for six cell sizes,
// zero mismatches for 1e-30 <= |x| <= 1e30; outside that range the weights / the
// no-cell outcome are identical anyway).  Five packed FFMA2/FMUL2 for two divisions instead of two
// MUFU.RCP + FCHK + branch sequences.  div_fast = 0 selects plain division.
template <class T>
__device__ __forceinline__ T grid_coord(const StepParams& P, T p, int axis) {
    const T x = vsub(p, bc<T>(P.origin[axis]));
    if (P.div_fast) {
        const T rh = bc<T>(P.rh), mh = bc<T>(-P.h);
        const T q0 = vmul(x, rh);
        const T q1 = vfma(vfma(mh, q0, x), rh, q0);
        return vfma(vfma(mh, q1, x), rh, q1);
    }
    T g;
#pragma unroll
    for (int i = 0; i < VecTraits<T>::n; ++i) setel(g, i, __fdiv_rn(el(x, i), P.h));
    return g;
}

// The shader's [max(floor,0), min(floor+1,G-1)] cell range is exactly "cells f and f+1, each kept only
// if it lies in [0,G-1]".  Weights clamp(1-|g-cell|,0,1): for cell f, g-f is in [0,1) so the weight is
// 1-(g-f); for cell f+1, g-(f+1) is in [-1,0) so it is 1+(g-(f+1)); same roundings as the shader's
// expression, no abs / clamp needed.  f = floor(g) comes from ONE conversion (F2I.FLOOR saturates far-away points)
// and is clamped to [-2, G] as an integer: no valid cell outside the grid.  A NaN coordinate converts to 0, so the
// callers exclude NaN points explicitly (nan3 below): they touch no cell, as in the oracle.
template <class T> struct AxisCells { T w0, w1; int f[VecTraits<T>::n]; };
__device__ __forceinline__ bool nan3(float x, float y, float z) { const float t = x + y + z; return !(t == t); }

// CVT_FLOOR: f from one F2I.FLOOR + integer clamp instead of FRND.FLOOR + float clamp + F2I (fewer registers: the gather inside
// k_ftl_step is register-bound).  Same f.  (Round 2: the magic-number floor k_grid_splat uses -- FADD.RM with 1.5*2^23 -- was tried
// here too: fewer instructions, but 88 B of spills at 96 registers, k_ftl_step 0.343 -> 0.411 ms; at 128 registers no change.)
template <class T, bool CVT_FLOOR>
__device__ __forceinline__ AxisCells<T> axis_cells(const StepParams& P, T p, int axis) {
    constexpr int n = VecTraits<T>::n;
    AxisCells<T> a;
    const T g = grid_coord<T>(P, p, axis);
    T fl;
#pragma unroll
    for (int i = 0; i < n; ++i) {
        if (CVT_FLOOR) {
            const int f = max(-2, min(__float2int_rd(el(g, i)), P.G));
            setel(fl, i, (float)f);
            a.f[i] = f;
        } else {
            const float f = fminf(fmaxf(floorf(el(g, i)), -2.0f), (float)P.G);
            setel(fl, i, f);
            a.f[i] = (int)f;
        }
    }
    a.w0 = vfma(vsub(g, fl), bc<T>(-1.0f), bc<T>(1.0f));
    a.w1 = vadd(vsub(g, vadd(fl, bc<T>(1.0f))), bc<T>(1.0f));
    return a;
}
__device__ __forceinline__ bool cell_ok(int f, int G) { return (unsigned)f < (unsigned)G; }

// ---- P2 splat: compute.comp:231-252 -----------------------------------------------------------
// Per corner the shader adds int(SCALE * (w * v_k)) and int(SCALE * w), truncated toward zero; the float
// operations below are ordered exactly as the shader orders them, so the integers are the reference's.
// The accumulators are int64 (int32 overflows beyond ~50K strands); integer sums are
// order-independent, which is what makes the warp aggregation below (and the multi-GPU all-reduce) exact.
__device__ __forceinline__ void global_add(unsigned long long* __restrict__ p, long long v) {
    if (v) atomicAdd(p, (unsigned long long)v);                        // RED.E.ADD.64, fire and forget
}

// Slow path: one lane splats ONE point straight into the global grid (8 corners x 4 atomics).
// 64-bit conversions, so |SCALE*w*v| >= 2^31 keeps its value as in the int64 oracle.
__device__ __forceinline__ void splat_point_direct(const StepParams& P, unsigned long long* __restrict__ grid,
                                                   int fx, int fy, int fz, const float (&wx)[2], const float (&wy)[2], const float (&wz)[2],
                                                   float vx, float vy, float vz) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        if (!cell_ok(fx + a, P.G)) continue;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            if (!cell_ok(fy + b, P.G)) continue;
            const float xyw = __fmul_rn(wx[a], wy[b]);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (!cell_ok(fz + c, P.G)) continue;
                const float tw = __fmul_rn(xyw, wz[c]);
                unsigned long long* cell = grid + 4 * (size_t)((fx + a) + ((fy + b) + (fz + c) * P.G) * P.G);
                global_add(cell + 0, __float2ll_rz(__fmul_rn(P.scale, __fmul_rn(tw, vx))));
                global_add(cell + 1, __float2ll_rz(__fmul_rn(P.scale, __fmul_rn(tw, vy))));
                global_add(cell + 2, __float2ll_rz(__fmul_rn(P.scale, __fmul_rn(tw, vz))));
                global_add(cell + 3, __float2ll_rz(__fmul_rn(P.scale, tw)));
            }
        }
    }
}

// ---- P3 gather + friction of one pack: compute.comp:259-297 ---------------------------------------
// fgrid[cell] = (vx, vy, vz) / density as floats, .w = float(density) (all zero where density <= 0), prepared once per step by
// k_grid_finalize, so a corner costs one LDG.128 and three FMAs (one FFMA2 + one FFMA).  Floating point:
// agrees with the shader's  w * (1/d) * v  to a few ulp (tolerance-checked, not bit-compared).
// REP (extension, include/rvh.h RVH_REPULSION_ON): with rho = sum_c w_c D_c the trilinear density over the 8 corner cells
// (D_c = fgrid.w = float(density), 0 where density <= 0) and grad rho its analytic gradient (d w/d g = -1 for cell f, +1 for
// cell f+1, times 1/h), the velocity also gets  v -= repulsion * h * grad(rho) / sum_c D_c  after the friction blend: a push
// down the density gradient whose every component is bounded by `repulsion` whatever the strand count (|h grad rho| <= sum D).
// Oracle twin: oracle.c gather_strand.
// Collider candidates: the gather has each point's grid cell in hand anyway, so it also looks up a host-built byte per
// coarse (2x2x2-cell) box: bit j set <=> ellipsoid j can contain a point of that box (conservative).  `cand` returns the OR
// over the pack (all ones when a point is outside the grid); the caller ORs it over the warp and k_ftl_step then runs only
// the ellipsoid tests that can hit -- typically 1-2 of 5, each worth 12 FFMA2 and as many constant loads.
// COH: weak (coherent) ld.global instead of the non-coherent read-only path.  The persistent k_scene_step reads data that other
// CTAs of the SAME launch wrote before a grid-wide barrier (float grid, planes): ld.global.nc / LDG.CONSTANT may return stale lines there.
// cell = (float(vel) * (1 / float(density))) per component, zero where density <= 0 (compute.comp:276-286)
__device__ __forceinline__ float4 finalize_cell(longlong2 v01, longlong2 v2d, int int32_wrap) {
    long long dens = v2d.y, v0 = v01.x, v1 = v01.y, v2 = v2d.x;
    if (int32_wrap) { dens = (int)dens; v0 = (int)v0; v1 = (int)v1; v2 = (int)v2; }   // the reference's int32 GridCell
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dens > 0) {
        const float rd = __frcp_rn(__ll2float_rn(dens));                 // 1.0 / float(density), compute.comp:283
        o.x = __ll2float_rn(v0) * rd; o.y = __ll2float_rn(v1) * rd; o.z = __ll2float_rn(v2) * rd; o.w = __ll2float_rn(dens);
    }
    return o;
}

// SRC 0: float grid through the read-only path; 1: float grid, coherent; 2: `fg` really is the int64 accumulator grid -- the cell
// is converted on the fly exactly as k_grid_finalize would have (k_scene_step: no finalize phase, no float grid at all).
template <int SRC> __device__ __forceinline__ float4 ld_cell(const float4* fg, int idx, int int32_wrap) {
    if (SRC == 0) return __ldg(fg + idx);
    if (SRC == 1) {
        float4 v;
        asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(fg + idx));
        return v;
    }
    const unsigned long long* c = reinterpret_cast<const unsigned long long*>(fg) + 4 * (size_t)idx;
    longlong2 a, b;
    asm volatile("ld.global.v2.s64 {%0, %1}, [%2];" : "=l"(a.x), "=l"(a.y) : "l"(c));
    asm volatile("ld.global.v2.s64 {%0, %1}, [%2];" : "=l"(b.x), "=l"(b.y) : "l"(c + 2));
    return finalize_cell(a, b, int32_wrap);
}
template <bool COH> __device__ __forceinline__ float ld_plane(const float* p) {
    if (!COH) return __ldg(p);
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <class T, bool REP, bool CMASK = false, int SRC = 0>
__device__ __forceinline__ void gather_pack(const StepParams& P, const float4* __restrict__ fgrid, T px, T py, T pz, T& vx, T& vy, T& vz, unsigned& cand) {
    constexpr int n = VecTraits<T>::n;
    float rho[n], rgx[n], rgy[n], rgz[n];
#pragma unroll
    for (int i = 0; i < n; ++i) { rho[i] = 0.f; rgx[i] = 0.f; rgy[i] = 0.f; rgz[i] = 0.f; }
    const AxisCells<T> X = axis_cells<T, false>(P, px, 0), Y = axis_cells<T, false>(P, py, 1), Z = axis_cells<T, false>(P, pz, 2);
    float2 gxy[n]; float gz[n];
    bool interior = true, isnan_[n];
#pragma unroll
    for (int i = 0; i < n; ++i) {
        gxy[i] = make_float2(0.f, 0.f); gz[i] = 0.f;
        isnan_[i] = nan3(el(px, i), el(py, i), el(pz, i));
        interior = interior && cell_ok(X.f[i], P.G - 1) && cell_ok(Y.f[i], P.G - 1) && cell_ok(Z.f[i], P.G - 1) && !isnan_[i];
    }
    const T xy[4] = { vmul(X.w0, Y.w0), vmul(X.w1, Y.w0), vmul(X.w0, Y.w1), vmul(X.w1, Y.w1) };
    cand = ~0u;
    if (CMASK && interior) {
        cand = 0u;
#pragma unroll
        for (int i = 0; i < n; ++i)
            cand |= __ldg(P.cmask + ((X.f[i] >> 1) + ((Y.f[i] >> 1) + (Z.f[i] >> 1) * P.cmask_dim) * P.cmask_dim));
    }
    if (interior) {                                                     // all 8 corners of every point are cells
        int base[n];
#pragma unroll
        for (int i = 0; i < n; ++i) base[i] = X.f[i] + (Y.f[i] + Z.f[i] * P.G) * P.G;
        const int sy = P.G, sz = P.G * P.G;
#pragma unroll
        for (int cz = 0; cz < 2; ++cz)
#pragma unroll
            for (int ab = 0; ab < 4; ++ab) {
                const T tw = vmul(xy[ab], cz ? Z.w1 : Z.w0);
                const int off = (ab & 1) + (ab >> 1) * sy + cz * sz;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const float4 cell = ld_cell<SRC>(fgrid, base[i] + off, P.int32_wrap);
                    const float w = el(tw, i);
                    gxy[i] = __ffma2_rn(make_float2(w, w), make_float2(cell.x, cell.y), gxy[i]);
                    gz[i] = fmaf(w, cell.z, gz[i]);
                    if (REP) {
                        const float D = cell.w, wz = el(cz ? Z.w1 : Z.w0, i);
                        rho[i] += D;
                        rgx[i] = fmaf((ab & 1) ? D : -D, el((ab >> 1) ? Y.w1 : Y.w0, i) * wz, rgx[i]);
                        rgy[i] = fmaf((ab >> 1) ? D : -D, el((ab & 1) ? X.w1 : X.w0, i) * wz, rgy[i]);
                        rgz[i] = fmaf(cz ? D : -D, el(xy[ab], i), rgz[i]);
                    }
                }
            }
    } else {
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int cz = 0; cz < 2; ++cz)
#pragma unroll
                for (int ab = 0; ab < 4; ++ab) {
                    const int fx = X.f[i] + (ab & 1), fy = Y.f[i] + (ab >> 1), fz = Z.f[i] + cz;
                    if (!isnan_[i] && cell_ok(fx, P.G) && cell_ok(fy, P.G) && cell_ok(fz, P.G)) {
                        const float4 cell = ld_cell<SRC>(fgrid, fx + (fy + fz * P.G) * P.G, P.int32_wrap);
                        const float w = el(xy[ab], i) * el(cz ? Z.w1 : Z.w0, i);
                        gxy[i] = __ffma2_rn(make_float2(w, w), make_float2(cell.x, cell.y), gxy[i]);
                        gz[i] = fmaf(w, cell.z, gz[i]);
                        if (REP) {
                            const float D = cell.w, wz = el(cz ? Z.w1 : Z.w0, i);
                            rho[i] += D;
                            rgx[i] = fmaf((ab & 1) ? D : -D, el((ab >> 1) ? Y.w1 : Y.w0, i) * wz, rgx[i]);
                            rgy[i] = fmaf((ab >> 1) ? D : -D, el((ab & 1) ? X.w1 : X.w0, i) * wz, rgy[i]);
                            rgz[i] = fmaf(cz ? D : -D, el(xy[ab], i), rgz[i]);
                        }
                    }
                }
    }
    T gx, gy, gzz;
#pragma unroll
    for (int i = 0; i < n; ++i) { setel(gx, i, gxy[i].x); setel(gy, i, gxy[i].y); setel(gzz, i, gz[i]); }
    const T fr = bc<T>(P.friction), omf = bc<T>(1.0f - P.friction);     // :296-297
    vx = vfma(fr, gx, vmul(omf, vx)); vy = vfma(fr, gy, vmul(omf, vy)); vz = vfma(fr, gzz, vmul(omf, vz));
    if (REP) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
            if (rho[i] > 0.f) {
                const float k = -P.repulsion / rho[i];                  // rho = sum of the corner densities
                setel(vx, i, fmaf(k, rgx[i], el(vx, i))); setel(vy, i, fmaf(k, rgy[i], el(vy, i))); setel(vz, i, fmaf(k, rgz[i], el(vz, i)));
            }
        }
    }
}

// Optional (RVH_K1_PREFETCH=1, off by default): L1 prefetch of the cells gather_pack is about to read, issued at the top of a
// row, with the gather itself moved behind the collision tests so that their work hides the L2 latency.  Measured on
// B200 at 1M x 32 (k_ftl_step with fused gather): gather first, no prefetch 0.356 ms; prefetch + late gather 0.481 ms at
// 6 CTAs/SM, 0.417 ms at 5 -- the velocities stay live across the collision code and the extra spills cost more than
// the latency they hide.  Unrolling the row loop by 2 (RVH_K1_UNROLL) to ping-pong the prefetch registers: 0.416 ms.
#ifndef RVH_K1_PREFETCH
#define RVH_K1_PREFETCH 0
#endif
template <class T>
__device__ __forceinline__ void gather_prefetch(const StepParams& P, const float4* __restrict__ fgrid, T px, T py, T pz) {
    constexpr int n = VecTraits<T>::n;
    const AxisCells<T> X = axis_cells<T, false>(P, px, 0), Y = axis_cells<T, false>(P, py, 1), Z = axis_cells<T, false>(P, pz, 2);
#pragma unroll
    for (int i = 0; i < n; ++i) {
        if (cell_ok(X.f[i], P.G - 1) && cell_ok(Y.f[i], P.G - 1) && cell_ok(Z.f[i], P.G - 1)) {
            const float4* b = fgrid + (X.f[i] + (Y.f[i] + Z.f[i] * P.G) * P.G);
            const int sy = P.G, sz = P.G * P.G;
            asm volatile("prefetch.global.L1 [%0];" :: "l"(b));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(b + sy));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(b + sz));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(b + sz + sy));
        }
    }
}

// ---- head SDF sampling (north-star extension; oracle twin: oracle.c sdf_cell / sdf_trilinear) -------------
// Lattice coordinate u = (p - origin) * inv_cell per axis; the point is "in the volume" when all 8 nodes of its
// cell exist.  d = trilinear interpolation, x then y then z, each lerp = fma(t, b - a, a): the same operations in
// the same order as the oracle, so the inside/outside decision (d < 0) is bit-identical on both sides.
// One TMA-staged tile = 8 x 4 x 4 nodes (512 B), private to a WARP (see k_ftl_step).  TMA needs the box's global start
// address 16-byte aligned, i.e. the x coordinate a multiple of 4 floats (an unaligned x coordinate faults as "illegal
// instruction": scripts/probes/tma_probe.cu), so the tile's x origin is the row's minimum cell rounded DOWN to 4 and the
// box is 8 wide: reach >= 4 x 3 x 3 cells from the minimum cell of the warp's 32*V neighbouring strands.
constexpr int kSdfBoxX = 8, kSdfBoxY = 4, kSdfBoxZ = 4;
constexpr int kSdfTileFloats = kSdfBoxX * kSdfBoxY * kSdfBoxZ;
constexpr int kSdfTileBytes = kSdfTileFloats * 4;
struct SdfTile {                              // the staged tile as the consumer threads see it
    const float* data;                        // shared memory [4][4][8], x fastest; nullptr = no tile (plain loads)
    int ox, oy, oz;                           // node index of tile element (0,0,0)
};

__device__ __forceinline__ bool sdf_cell(const SdfVolume& V, float x, float y, float z, int& ix, int& iy, int& iz,
                                         float& tx, float& ty, float& tz) {
    const float ux = __fmul_rn(__fsub_rn(x, V.origin[0]), V.inv_cell);
    const float uy = __fmul_rn(__fsub_rn(y, V.origin[1]), V.inv_cell);
    const float uz = __fmul_rn(__fsub_rn(z, V.origin[2]), V.inv_cell);
    const bool in = ux >= 0.f && ux < (float)(V.nx - 1) && uy >= 0.f && uy < (float)(V.ny - 1) && uz >= 0.f && uz < (float)(V.nz - 1);
    ix = (int)ux; iy = (int)uy; iz = (int)uz;                          // u >= 0: truncation = floor (unused when !in)
    tx = __fsub_rn(ux, (float)ix); ty = __fsub_rn(uy, (float)iy); tz = __fsub_rn(uz, (float)iz);
    return in;
}

__device__ __forceinline__ void sdf_corners(const SdfVolume& V, const SdfTile& T, int ix, int iy, int iz, float (&c)[8]) {
    if (T.data) {
        const int lx = ix - T.ox, ly = iy - T.oy, lz = iz - T.oz;
        if ((unsigned)lx < (unsigned)(kSdfBoxX - 1) && (unsigned)ly < (unsigned)(kSdfBoxY - 1) && (unsigned)lz < (unsigned)(kSdfBoxZ - 1)) {
            constexpr int sy = kSdfBoxX, sz = kSdfBoxX * kSdfBoxY;
            const float* t = T.data + lx + sy * ly + sz * lz;
            c[0] = t[0]; c[1] = t[1]; c[2] = t[sy]; c[3] = t[sy + 1];
            c[4] = t[sz]; c[5] = t[sz + 1]; c[6] = t[sz + sy]; c[7] = t[sz + sy + 1];
            return;
        }
    }
    const size_t sy = (size_t)V.nxp, sz = (size_t)V.nxp * V.ny;        // the row's bounding box left the tile: plain loads
    const float* g = V.data + ix + sy * iy + sz * iz;
    c[0] = __ldg(g); c[1] = __ldg(g + 1); c[2] = __ldg(g + sy); c[3] = __ldg(g + sy + 1);
    c[4] = __ldg(g + sz); c[5] = __ldg(g + sz + 1); c[6] = __ldg(g + sz + sy); c[7] = __ldg(g + sz + sy + 1);
}

__device__ __forceinline__ float sdf_trilinear(const float (&c)[8], float tx, float ty, float tz) {
    const float c00 = fmaf(tx, __fsub_rn(c[1], c[0]), c[0]), c10 = fmaf(tx, __fsub_rn(c[3], c[2]), c[2]);
    const float c01 = fmaf(tx, __fsub_rn(c[5], c[4]), c[4]), c11 = fmaf(tx, __fsub_rn(c[7], c[6]), c[6]);
    const float c0 = fmaf(ty, __fsub_rn(c10, c00), c00), c1 = fmaf(ty, __fsub_rn(c11, c01), c01);
    return fmaf(tz, __fsub_rn(c1, c0), c0);
}

// is the point inside the head?  (the test half; the force half is in collision_force)
__device__ __forceinline__ bool sdf_inside(const SdfVolume& V, const SdfTile& T, float x, float y, float z) {
    int ix, iy, iz; float tx, ty, tz;
    if (!sdf_cell(V, x, y, z, ix, iy, iz, tx, ty, tz)) return false;
    float c[8];
    sdf_corners(V, T, ix, iy, iz, c);
    return sdf_trilinear(c, tx, ty, tz) < 0.f;
}

// ---- TMA / mbarrier primitives for the staged SDF tiles (sm_100a PTX) ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a tile that never lands is a bug, and a trap is better than a hung GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    for (unsigned spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---- P1 for one point (or one pack of two): compute.comp:144-201 --------------------------------
template <class T> struct PointOut { T px, py, pz, vx, vy, vz, dx, dy, dz; };

// q = inv * (p,1) of ellipsoid E and |q|^2 (compute.comp:64-67)
template <class T>
__device__ __forceinline__ T ellipsoid_q(const Ellipsoid& E, T cx, T cy, T cz, T& qx, T& qy, T& qz) {
    qx = vfma(bc<T>(E.inv[0]), cx, vfma(bc<T>(E.inv[1]), cy, vfma(bc<T>(E.inv[2]), cz, bc<T>(E.inv[3]))));
    qy = vfma(bc<T>(E.inv[4]), cx, vfma(bc<T>(E.inv[5]), cy, vfma(bc<T>(E.inv[6]), cz, bc<T>(E.inv[7]))));
    qz = vfma(bc<T>(E.inv[8]), cx, vfma(bc<T>(E.inv[9]), cy, vfma(bc<T>(E.inv[10]), cz, bc<T>(E.inv[11]))));
    return vdot3(qx, qy, qz, qx, qy, qz);
}

// Penalty force of the colliders in `hit` on ONE point (the divergent part; compute.comp:160-184).
// SDF: bit 1 of `hit` is the head volume (penalty_k * (-d) * normalize(grad d), include/rvh.h) instead of ellipsoid 1.
template <bool SDF>
__device__ __forceinline__ void collision_force(const StepParams& P, const SdfTile& tile, unsigned hit, float cx, float cy, float cz,
                                                float& ox, float& oy, float& oz) {
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (hit & 1u) {                                                     // sphere, :160-169
        const float dx = cx - P.sphere_c[0], dy = cy - P.sphere_c[1], dz = cz - P.sphere_c[2];
        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        const float rinv = rsqrt_fast(d2);
        const float s = P.penalty_k * (P.sphere_r - d2 * rinv) * rinv;
        ax = s * dx; ay = s * dy; az = s * dz;
    }
    if (SDF && (hit & 2u)) {
        int ix, iy, iz; float tx, ty, tz;
        sdf_cell(P.sdf, cx, cy, cz, ix, iy, iz, tx, ty, tz);
        float c[8];
        sdf_corners(P.sdf, tile, ix, iy, iz, c);
        const float dx00 = __fsub_rn(c[1], c[0]), dx10 = __fsub_rn(c[3], c[2]), dx01 = __fsub_rn(c[5], c[4]), dx11 = __fsub_rn(c[7], c[6]);
        const float c00 = fmaf(tx, dx00, c[0]), c10 = fmaf(tx, dx10, c[2]), c01 = fmaf(tx, dx01, c[4]), c11 = fmaf(tx, dx11, c[6]);
        const float dy0 = __fsub_rn(c10, c00), dy1 = __fsub_rn(c11, c01);
        const float c0 = fmaf(ty, dy0, c00), c1 = fmaf(ty, dy1, c01);
        const float gz = __fsub_rn(c1, c0);
        const float d = fmaf(tz, gz, c0);                              // the same d sdf_inside found < 0
        const float gx0 = fmaf(ty, dx10 - dx00, dx00), gx1 = fmaf(ty, dx11 - dx01, dx01);
        const float gx = fmaf(tz, gx1 - gx0, gx0), gy = fmaf(tz, dy1 - dy0, dy0);   // gradient of the trilinear interpolant (x inv_cell, cancels)
        const float g2 = fmaf(gx, gx, fmaf(gy, gy, gz * gz));
        if (g2 > 0.f) {
            const float sc = P.penalty_k * (-d) * rsqrt_fast(g2);
            ax = fmaf(sc, gx, ax); ay = fmaf(sc, gy, ay); az = fmaf(sc, gz, az);
        }
    }
    unsigned m = SDF ? 0u : hit >> 1;
    while (m) {                                                         // ellipsoids, :170-179
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const Ellipsoid& E = P.ell[j];
        float qx, qy, qz;
        const float q2 = ellipsoid_q<float>(E, cx, cy, cz, qx, qy, qz);
        const float rq = rsqrt_fast(q2);
        const float ux = qx * rq, uy = qy * rq, uz = qz * rq;
        const float sx = fmaf(E.xf[0], ux, fmaf(E.xf[1], uy, fmaf(E.xf[2], uz, E.xf[3])));
        const float sy = fmaf(E.xf[4], ux, fmaf(E.xf[5], uy, fmaf(E.xf[6], uz, E.xf[7])));
        const float sz = fmaf(E.xf[8], ux, fmaf(E.xf[9], uy, fmaf(E.xf[10], uz, E.xf[11])));
        const float ex = cx - sx, ey = cy - sy, ez = cz - sz;
        const float e2 = fmaf(ex, ex, fmaf(ey, ey, ez * ez));
        const float d = e2 > 0.f ? e2 * rsqrt_fast(e2) : 0.f;
        const float nx = fmaf(E.nt[0], qx, fmaf(E.nt[1], qy, E.nt[2] * qz));
        const float ny = fmaf(E.nt[3], qx, fmaf(E.nt[4], qy, E.nt[5] * qz));
        const float nz = fmaf(E.nt[6], qx, fmaf(E.nt[7], qy, E.nt[8] * qz));
        const float s = P.penalty_k * d * rsqrt_fast(fmaf(nx, nx, fmaf(ny, ny, nz * nz)));
        ax = fmaf(s, nx, ax); ay = fmaf(s, ny, ay); az = fmaf(s, nz, az);
    }
    const float ih = rcp_fast((float)__popc(hit));                      // :182-184
    ox = ax * ih; oy = ay * ih; oz = az * ih;
}

// NELL >= 0: number of ellipsoids known at compile time (+100: with the collider candidate mask); NELL == -1: run-time
// count; NELL <= -2: the head SDF replaces the ellipsoids (-2 plain loads, -3 TMA-staged tile in `tile`).
// GATHER != 0: the previous step's gather (+ repulsion when 2) is applied to (vx,vy,vz) right before they are first used,
// i.e. AFTER the collision tests, whose work hides the latency of the cells prefetched at the top.
struct WindNow { float amp_s2T, T3, amp; };   // this step's time-only wind scalars (host-evaluated): wind_amp * 2 sin(2T), 3T mod 2 pi, wind_amp
template <class T, bool WIND, int NELL, int GATHER, bool COH = false>
__device__ __forceinline__ PointOut<T> point_update(const StepParams& P, const WindNow& W, const SdfTile& tile, const float4* __restrict__ fgrid,
                                                    T cx, T cy, T cz, T vx, T vy, T vz, T parx, T pary, T parz) {
    constexpr int n = VecTraits<T>::n;
    // NELL >= 100: NELL - 100 unrolled ellipsoids, tested only where the collider candidate mask allows (needs GATHER)
    constexpr bool CMASK = NELL >= 100 && GATHER != 0 && !RVH_K1_PREFETCH;
    constexpr int NE = NELL >= 100 ? NELL - 100 : NELL;
    unsigned cand = ~0u;                                                // ellipsoids that can hit a point of this warp's row
    if (GATHER && RVH_K1_PREFETCH) gather_prefetch<T>(P, fgrid, cx, cy, cz);
    if (GATHER && !RVH_K1_PREFETCH) {
        gather_pack<T, GATHER == 2, CMASK, COH ? 1 : 0>(P, fgrid, cx, cy, cz, vx, vy, vz, cand);
        if (CMASK) cand = __reduce_or_sync(0xffffffffu, cand);          // warp-uniform: the skipped tests cost no divergence
    }
    T fx = bc<T>(0.0f), fy = bc<T>(P.gravity_y), fz = bc<T>(0.0f);       // :150
    if (WIND) {
        const T a1 = vmul(cy, bc<T>(10.0f));
        const T a2 = vmul(vadd(cy, bc<T>(5.0f)), bc<T>(15.0f));
        T cs, sn;
#pragma unroll
        for (int i = 0; i < n; ++i) { setel(cs, i, cos_bounded(el(a1, i))); setel(sn, i, sin_bounded(el(a2, i))); }
        fx = vfma(bc<T>(W.amp_s2T), vmul(cs, sn), fx);
        if (P.wind_mode == 1) {                                         // :151
            T cl;
#pragma unroll
            for (int i = 0; i < n; ++i) setel(cl, i, fminf(fmaxf(el(cy, i) * 2.0f, 0.2f), 2.0f));
            fz = vfma(bc<T>(-W.amp), cl, fz);
        } else {                                                        // :152
            const T a3 = vfma(cz, bc<T>(5.0f), bc<T>(W.T3));
            T s3;
#pragma unroll
            for (int i = 0; i < n; ++i) setel(s3, i, sin_bounded(el(a3, i)));
            fy = vfma(bc<T>(4.0f * W.amp), s3, fy);
            fz = vfma(bc<T>(-0.6f * W.amp), vadd(cy, bc<T>(3.0f)), fz);
        }
    }

    // collision tests first (one predicate per collider), bodies only for the colliders hit
    unsigned hit[n];
    {
        const T dx = vsub(cx, bc<T>(P.sphere_c[0])), dy = vsub(cy, bc<T>(P.sphere_c[1])), dz = vsub(cz, bc<T>(P.sphere_c[2]));
        const T d2 = vdot3(dx, dy, dz, dx, dy, dz);
#pragma unroll
        for (int i = 0; i < n; ++i) hit[i] = (el(d2, i) < P.sphere_r2) ? (unsigned)P.has_sphere : 0u;   // :162
    }
    if (NE >= 0) {
#pragma unroll
        for (int j = 0; j < (NE >= 0 ? NE : 0); ++j) {
            if (CMASK && !(cand & (1u << j))) continue;
            T qx, qy, qz;
            const T q2 = ellipsoid_q<T>(P.ell[j], cx, cy, cz, qx, qy, qz);
#pragma unroll
            for (int i = 0; i < n; ++i) if (el(q2, i) <= 1.0f) hit[i] |= 2u << j;                       // :66
        }
    } else if (NE == -1) {
        for (int j = 0; j < P.n_ell; ++j) {
            T qx, qy, qz;
            const T q2 = ellipsoid_q<T>(P.ell[j], cx, cy, cz, qx, qy, qz);
#pragma unroll
            for (int i = 0; i < n; ++i) if (el(q2, i) <= 1.0f) hit[i] |= 2u << j;
        }
    } else {
#pragma unroll
        for (int i = 0; i < n; ++i) if (sdf_inside(P.sdf, tile, el(cx, i), el(cy, i), el(cz, i))) hit[i] |= 2u;
    }
    unsigned any_hit = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) any_hit |= hit[i];
    if (any_hit) {
        T ax = bc<T>(0.f), ay = bc<T>(0.f), az = bc<T>(0.f);
#pragma unroll
        for (int i = 0; i < n; ++i) {
            if (hit[i]) {
                float ox, oy, oz;
                collision_force<(NELL <= -2)>(P, tile, hit[i], el(cx, i), el(cy, i), el(cz, i), ox, oy, oz);
                setel(ax, i, ox); setel(ay, i, oy); setel(az, i, oz);
            }
        }
        fx = vadd(fx, ax); fy = vadd(fy, ay); fz = vadd(fz, az);
    }

    if (GATHER && RVH_K1_PREFETCH) { unsigned late; gather_pack<T, GATHER == 2>(P, fgrid, cx, cy, cz, vx, vy, vz, late); }
    PointOut<T> o;
    const T dt = bc<T>(P.dt), dt2 = bc<T>(P.dt2);
    const T prx = vfma(dt2, fx, vfma(dt, vx, cx));                      // :187
    const T pry = vfma(dt2, fy, vfma(dt, vy, cy));
    const T prz = vfma(dt2, fz, vfma(dt, vz, cz));
    const T ddx = vsub(prx, parx), ddy = vsub(pry, pary), ddz = vsub(prz, parz);   // :191-192
    const T l = vdot3(ddx, ddy, ddz, ddx, ddy, ddz);
    T sc;
#pragma unroll
    for (int i = 0; i < n; ++i) setel(sc, i, P.rest * rsqrt_fast(el(l, i)));
    o.px = vfma(sc, ddx, parx); o.py = vfma(sc, ddy, pary); o.pz = vfma(sc, ddz, parz);
    const T vs = bc<T>(P.vel_scale);
    T nvx = vmul(vsub(o.px, cx), vs), nvy = vmul(vsub(o.py, cy), vs), nvz = vmul(vsub(o.pz, cz), vs);   // :195-197
    const T l2 = vdot3(nvx, nvy, nvz, nvx, nvy, nvz);
    bool clampv = false;
#pragma unroll
    for (int i = 0; i < n; ++i) clampv |= el(l2, i) > P.vmax2;
    if (clampv) {                                                       // :198-200
        T s;
#pragma unroll
        for (int i = 0; i < n; ++i) setel(s, i, el(l2, i) > P.vmax2 ? P.vmax * rsqrt_fast(el(l2, i)) : 1.0f);
        nvx = vmul(nvx, s); nvy = vmul(nvy, s); nvz = vmul(nvz, s);
    }
    o.vx = nvx; o.vy = nvy; o.vz = nvz;
    const T dmp = bc<T>(P.damping);
    o.dx = vmul(dmp, vsub(o.px, prx)); o.dy = vmul(dmp, vsub(o.py, pry)); o.dz = vmul(dmp, vsub(o.pz, prz));  // :201
    return o;
}

// ---- K1: integrate + collide + FTL + corrected velocity -------------------------------------
// One thread owns V consecutive strands and walks them root->tip together.  V = 1: scalar; V = 2:
// the two strands form one fp32x2 pack (LDG.64 delivers it directly).  Point i's
// velocity is final only once d_{i+1} is known (compute.comp:213-215), so the velocity store trails the
// position by one point.
template <int V> struct PackOf { using T = float2; static constexpr int n = V / 2; };
template <> struct PackOf<1> { using T = float; static constexpr int n = 1; };

// RVH_K1_STREAM_LOADS: the strand planes are read exactly once per step, so their loads can skip L1 allocation
// (ld.global.nc.L1::no_allocate) and leave the cache to the gather's float grid, which neighbouring strands re-read.
// Measured on B200 at 1M x 32: 0.338 vs 0.341 ms -- within noise, so it stays off.
#ifndef RVH_K1_STREAM_LOADS
#define RVH_K1_STREAM_LOADS 0
#endif
// COHERENT: plain ld.global instead of the non-coherent ld.global.nc (MULTI re-reads its own stores of the previous step)
template <int V, bool COHERENT = false> __device__ __forceinline__ void load_packs(const float* p, typename PackOf<V>::T (&o)[PackOf<V>::n]) {
#if RVH_K1_STREAM_LOADS
    if constexpr (V == 1) { asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(o[0]) : "l"(p)); }
    else if constexpr (V == 2) { asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(o[0].x), "=f"(o[0].y) : "l"(p)); }
    else { asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0].x), "=f"(o[0].y), "=f"(o[1].x), "=f"(o[1].y) : "l"(p)); }
#else
    if constexpr (V == 1) { o[0] = COHERENT ? *p : __ldg(p); }
    else if constexpr (V == 2) { o[0] = COHERENT ? *reinterpret_cast<const float2*>(p) : __ldg(reinterpret_cast<const float2*>(p)); }
    else { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); o[0] = make_float2(t.x, t.y); o[1] = make_float2(t.z, t.w); }
#endif
}
template <int V> __device__ __forceinline__ void store_packs(float* __restrict__ p, const typename PackOf<V>::T (&o)[PackOf<V>::n]) {
    if constexpr (V == 1) { *p = o[0]; }
    else if constexpr (V == 2) { *reinterpret_cast<float2*>(p) = o[0]; }
    else { *reinterpret_cast<float4*>(p) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y); }
}

// GATHER: the previous step's grid gather + friction (compute.comp:257-298) is applied to each velocity
// as it is loaded -- the same positions and the same fgrid the stand-alone k_grid_gather would use -- so
// the per-step K2 pass (re-read p,v, re-write v: 36 B/point) disappears from steady-state stepping.
// GATHER = 2 adds the repulsion extension to it.
//
// NELL == -3 (head SDF through TMA): a warp owns 32*V Morton-neighbouring strands, whose row-i points sit within a
// few SDF cells of each other.  One row AHEAD of its use the warp reduces the minimum lattice cell of the row's points
// (3 REDUX.MIN), lane 0 arms the warp's mbarrier and issues ONE cp.async.bulk.tensor.3d for the 8x4x4-node box at that
// corner (512 B, zero-filled outside the volume) into the warp's double-buffered tile; a row later the lanes wait on the
// mbarrier's phase and read their 8 corners with LDS.  Points whose cell falls outside the box take plain loads.
// Everything is warp-private (tile, mbarrier, box origin in registers): no __syncthreads, the warps of a CTA keep
// drifting apart as in the other variants.  Positions are prefetched TWO rows ahead (velocities one), so that the box
// corner is known a full row before the tile is needed.
#ifndef RVH_K1_MINBLOCKS
#define RVH_K1_MINBLOCKS 6      // <= 85 registers: 6 CTAs/SM measured fastest on B200 (5: 0.365 ms, 6: 0.357 ms, 7: 0.416 ms at 1M x 32)
#endif
#ifndef RVH_K1G_MINBLOCKS
#define RVH_K1G_MINBLOCKS 5     // with the fused gather (1M x 32, all on, collider candidate mask): 5 CTAs/SM (96 registers) 0.343 ms, 6 (80 registers, spills in the row loop) 0.406 ms; without the mask 4: 0.383, 5: 0.363, 6: 0.356 ms
#endif
#ifndef RVH_K1_UNROLL
#define RVH_K1_UNROLL 1
#endif
constexpr int kK1Unroll = RVH_K1_UNROLL;     // row-loop unrolling (2 lets the compiler ping-pong the prefetch registers instead of copying them)
#ifndef RVH_K1X_MINBLOCKS
#define RVH_K1X_MINBLOCKS 4     // extension variants (SDF and/or repulsion): more live state
#endif
struct __align__(128) SdfStageSmem {
    float tile[kBlock / 32][2][kSdfTileFloats];
    unsigned long long bar[kBlock / 32][2];
};

// Whole warp: reduce the minimum in-volume lattice cell of `row`'s points, lane 0 stages the warp's tile for that row.
// Returns the tile origin (node index) in o0..o2, identical in all lanes.
template <class T, int NP>
__device__ __forceinline__ void sdf_stage_row(const StepParams& P, const CUtensorMap* tmap, SdfStageSmem& sm, int row,
                                              const T (&px)[NP], const T (&py)[NP], const T (&pz)[NP], int& o0, int& o1, int& o2) {
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kNone = 0x7fffffff;
    int m0 = kNone, m1 = kNone, m2 = kNone;
#pragma unroll
    for (int u = 0; u < NP; ++u)
#pragma unroll
        for (int e = 0; e < VecTraits<T>::n; ++e) {
            int ix, iy, iz; float tx, ty, tz;
            if (sdf_cell(P.sdf, el(px[u], e), el(py[u], e), el(pz[u], e), ix, iy, iz, tx, ty, tz)) { m0 = min(m0, ix); m1 = min(m1, iy); m2 = min(m2, iz); }
        }
    m0 = __reduce_min_sync(kFull, m0); m1 = __reduce_min_sync(kFull, m1); m2 = __reduce_min_sync(kFull, m2);   // also: every lane is done with the tile this buffer held two rows ago
    o0 = m0 == kNone ? 0 : (m0 & ~3);      // 16-byte aligned box start (see kSdfBoxX)
    o1 = m1 == kNone ? 0 : m1;
    o2 = m2 == kNone ? 0 : m2;
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5, par = row & 1;
        mbar_expect_tx(&sm.bar[w][par], kSdfTileBytes);
        tma_load_3d(sm.tile[w][par], tmap, &sm.bar[w][par], o0, o1, o2);
    }
}

// One thread's strands, root -> tip, `nsteps` times.  MULTI: this step's wind scalars come from P.wind_tab[step0 + step]; COH: the
// planes (and, in point_update, the float grid) are read by coherent loads -- the launch itself wrote them (MULTI / k_scene_step).
template <int V, bool WIND, int NELL, int GATHER, bool MULTI, bool COH>
__device__ __forceinline__ void ftl_walk(const StepParams& P, float* __restrict__ planes, float* __restrict__ corr, const float4* __restrict__ fgrid,
                                         const CUtensorMap* sdf_map, SdfStageSmem& sm, int cta, int nsteps, int step0) {
    using T = typename PackOf<V>::T;
    constexpr int NP = PackOf<V>::n;
    constexpr bool TMA = NELL == -3;
    const int t = (P.cta0 + cta) * blockDim.x + threadIdx.x;
    const int s0 = t * V;
    if (!TMA && s0 >= P.S_pad) return;                            // TMA variant: S_pad is a multiple of kBlock*V (V <= 2), no partial CTA
    const size_t RS = (size_t)P.S_pad * 6;                      // elements per row (all tiles, six planes)
    constexpr int PK = kTileStrands;                              // plane stride inside a tile: compile-time offsets
    float* const base = planes + tiled_index(6, P.S_pad, 0, 0, s0);

    for (int step = 0; step < nsteps; ++step) {
    const WindNow W = MULTI ? WindNow{ P.wind_tab[3 * (step0 + step)], P.wind_tab[3 * (step0 + step) + 1], P.wind_tab[3 * (step0 + step) + 2] } : WindNow{ P.wind_amp * P.wind_s2T, P.wind_T3, P.wind_amp };
    T parx[NP], pary[NP], parz[NP];
    load_packs<V, COH>(base, parx); load_packs<V, COH>(base + PK, pary); load_packs<V, COH>(base + 2 * PK, parz);
    T nx[NP], ny[NP], nz[NP], nvx[NP], nvy[NP], nvz[NP];
    float* nextp = base + RS;                                     // row 1
    load_packs<V, COH>(nextp, nx); load_packs<V, COH>(nextp + PK, ny); load_packs<V, COH>(nextp + 2 * PK, nz);
    load_packs<V, COH>(nextp + 3 * PK, nvx); load_packs<V, COH>(nextp + 4 * PK, nvy); load_packs<V, COH>(nextp + 5 * PK, nvz);
    T n2x[NP], n2y[NP], n2z[NP];                                  // TMA only: positions two rows ahead
    unsigned phase = 0;                                           // TMA only: mbarrier phase bit per buffer
    int bx = 0, by = 0, bz = 0;                                   // TMA only: origin of the tile of the row being consumed
    const int wid = threadIdx.x >> 5;
    if constexpr (TMA) {
        if (P.N > 2) { load_packs<V, COH>(nextp + RS, n2x); load_packs<V, COH>(nextp + RS + PK, n2y); load_packs<V, COH>(nextp + RS + 2 * PK, n2z); }
        if ((threadIdx.x & 31) == 0) {
            mbar_init(&sm.bar[wid][0], 1); mbar_init(&sm.bar[wid][1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        sdf_stage_row<T, NP>(P, sdf_map, sm, 1, nx, ny, nz, bx, by, bz);
    }
    T lvx[NP], lvy[NP], lvz[NP];   // clamped velocity of the previous point, correction pending
#pragma unroll
    for (int u = 0; u < NP; ++u) { lvx[u] = bc<T>(0.f); lvy[u] = bc<T>(0.f); lvz[u] = bc<T>(0.f); }
    const T minus_inv_dt = bc<T>(-P.inv_dt);

    float* prevp = base;
#pragma unroll kK1Unroll
    for (int i = 1; i < P.N; ++i) {
        float* const curp = nextp;
        nextp += RS;
        T cx[NP], cy[NP], cz[NP], vx[NP], vy[NP], vz[NP];
#pragma unroll
        for (int u = 0; u < NP; ++u) { cx[u] = nx[u]; cy[u] = ny[u]; cz[u] = nz[u]; vx[u] = nvx[u]; vy[u] = nvy[u]; vz[u] = nvz[u]; }
        SdfTile tile = { nullptr, 0, 0, 0 };
        if constexpr (!TMA) {
            if (i + 1 < P.N) {
                load_packs<V, COH>(nextp, nx); load_packs<V, COH>(nextp + PK, ny); load_packs<V, COH>(nextp + 2 * PK, nz);
                load_packs<V, COH>(nextp + 3 * PK, nvx); load_packs<V, COH>(nextp + 4 * PK, nvy); load_packs<V, COH>(nextp + 5 * PK, nvz);
            }
        } else {
            if (i + 1 < P.N) {
#pragma unroll
                for (int u = 0; u < NP; ++u) { nx[u] = n2x[u]; ny[u] = n2y[u]; nz[u] = n2z[u]; }
                load_packs<V, COH>(nextp + 3 * PK, nvx); load_packs<V, COH>(nextp + 4 * PK, nvy); load_packs<V, COH>(nextp + 5 * PK, nvz);
                if (i + 2 < P.N) { load_packs<V, COH>(nextp + RS, n2x); load_packs<V, COH>(nextp + RS + PK, n2y); load_packs<V, COH>(nextp + RS + 2 * PK, n2z); }
            }
            const int b = i & 1;
            tile.data = sm.tile[wid][b]; tile.ox = bx; tile.oy = by; tile.oz = bz;
            if (i + 1 < P.N) sdf_stage_row<T, NP>(P, sdf_map, sm, i + 1, nx, ny, nz, bx, by, bz);   // tile of row i+1 flies while row i is computed
            mbar_wait(&sm.bar[wid][b], (phase >> b) & 1u);
            phase ^= 1u << b;
        }
        T fvx[NP], fvy[NP], fvz[NP], odx[NP], ody[NP], odz[NP];
#pragma unroll
        for (int u = 0; u < NP; ++u) {
            const PointOut<T> o = point_update<T, WIND, NELL, GATHER, COH>(P, W, tile, fgrid, cx[u], cy[u], cz[u], vx[u], vy[u], vz[u], parx[u], pary[u], parz[u]);
            parx[u] = o.px; pary[u] = o.py; parz[u] = o.pz;
            odx[u] = o.dx; ody[u] = o.dy; odz[u] = o.dz;
            // finalise point i-1: v_{i-1} -= d_i / dt   (compute.comp:213-215)
            fvx[u] = vfma(o.dx, minus_inv_dt, lvx[u]); fvy[u] = vfma(o.dy, minus_inv_dt, lvy[u]); fvz[u] = vfma(o.dz, minus_inv_dt, lvz[u]);
            lvx[u] = o.vx; lvy[u] = o.vy; lvz[u] = o.vz;
        }
        store_packs<V>(curp, parx); store_packs<V>(curp + PK, pary); store_packs<V>(curp + 2 * PK, parz);
        if (P.keep_corr) {
            float* c0 = corr + tiled_index(3, P.S_pad, i, 0, s0);
            store_packs<V>(c0, odx); store_packs<V>(c0 + PK, ody); store_packs<V>(c0 + 2 * PK, odz);
        }
        if (i > 1) { store_packs<V>(prevp + 3 * PK, fvx); store_packs<V>(prevp + 4 * PK, fvy); store_packs<V>(prevp + 5 * PK, fvz); }
        prevp = curp;
    }
    // last point: no correction term (compute.comp:213 `i != NUM_CURVE_POINTS - 1`)
    store_packs<V>(prevp + 3 * PK, lvx); store_packs<V>(prevp + 4 * PK, lvy); store_packs<V>(prevp + 5 * PK, lvz);
    }   // step
}

// MULTI (grid off only): P.multi_steps steps in ONE launch.  Without the grid the strands never interact, so a thread simply
// walks its strands again (its own stores of step k are its loads of step k+1; the state of a small scene sits in L1/L2);
// the time-only wind scalars of every step come from the host-built P.wind_tab.  A 16K x 32 scene is launch- and
// latency-bound (3.8 us of HBM time per step): this removes the launch, rvh_step_n uses it.
template <int V, bool WIND, int NELL, int GATHER, bool MULTI = false>
__global__ void __launch_bounds__(kBlock, (NELL <= -2 || GATHER == 2) ? RVH_K1X_MINBLOCKS : (GATHER ? RVH_K1G_MINBLOCKS : RVH_K1_MINBLOCKS))
k_ftl_step(const __grid_constant__ StepParams P, float* __restrict__ planes, float* __restrict__ corr,
           const float4* __restrict__ fgrid, const __grid_constant__ CUtensorMap sdf_map, uint4* __restrict__ grid_clear, unsigned grid_clear_n) {
    constexpr bool TMA = NELL == -3;
    static_assert(!MULTI || (GATHER == 0 && !TMA), "several steps per launch need independent strands: no grid, no staged tiles");
    // The step's grid clear (Renderer.cpp:2063) rides here when the launch is wide enough: this kernel never touches the
    // int64 accumulators (it reads the float grid), the splat that fills them comes after it in the stream, and whoever
    // read them last (finalize / exchange / a download) came before it.  Saves the memset launch.
    if (grid_clear)
        for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < grid_clear_n; k += gridDim.x * blockDim.x) grid_clear[k] = make_uint4(0u, 0u, 0u, 0u);
    __shared__ __align__(128) unsigned char sdf_raw[TMA ? sizeof(SdfStageSmem) : 16];
    SdfStageSmem& sm = *reinterpret_cast<SdfStageSmem*>(sdf_raw);
    ftl_walk<V, WIND, NELL, GATHER, MULTI, MULTI>(P, planes, corr, fgrid, &sdf_map, sm, blockIdx.x, MULTI ? P.multi_steps : 1, 0);
}

// ---- grid off, small scene: a WAVEFRONT over the steps -------------------------------------------------------------------
// k_ftl_step<..., MULTI> removes the launches of a small grid-less scene (C2: 16K x 32), but each of its threads still walks 31
// dependent rows per step, and 128 CTAs put one warp on a scheduler: ~0.5 us per row of pure dependent-issue latency, 16 us per step
// on an idle machine.  The root->tip chain cannot be split -- but consecutive STEPS can overlap: row i of step s+1 needs nothing but
// row i of step s (its final velocity, i.e. once step s has passed row i+1: compute.comp:213-215) and its own row i-1.  So W lanes
// share a strand, lane j runs step s0+j, two rows behind lane j-1:
//     tick k:  lane j updates row i = k - 2j (if 1 <= i <= N-1) and finalises the velocity of row i-1 (row N: just publishes the last one)
//     end of tick:  __shfl_up hands (new position of row i, final velocity of row i-1) to lane j+1, which uses the position two ticks
//                   and the velocity one tick later; lane 0 reads the rows from memory (one row prefetched), the last active lane writes them.
// W steps take (N-1) + 2(W-1) + 1 ticks instead of W(N-1), and the launch has W times the warps: the ticks of 3-4 warps per scheduler
// overlap.  Same device function per point (point_update<float, ...>) and the same operations as single steps: bit-identical.
// A warp = 32/W strands x W lanes; between batches of W steps the warp re-reads its own stores (coherent loads after __syncwarp).
template <bool WIND, int NELL, int W>
__global__ void __launch_bounds__(kBlock, 4)
k_ftl_wave(const __grid_constant__ StepParams P, float* __restrict__ planes, float* __restrict__ corr) {
    static_assert(W == 4 || W == 8, "lanes per strand");
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int PK = kTileStrands;
    const int lane = threadIdx.x & 31, j = lane % W;
    const int s = (blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5)) * (32 / W) + lane / W;   // strand of this lane group
    const bool live = s < P.S_pad;                               // whole groups are live or not; dead groups still take part in the shuffles
    const int sc = live ? s : 0;
    const size_t RS = (size_t)P.S_pad * 6;
    float* const base = planes + tiled_index(6, P.S_pad, 0, 0, sc);
    const float rx = base[0], ry = base[PK], rz = base[2 * PK];  // the root never moves
    const float minus_inv_dt = -P.inv_dt;
    const SdfTile no_tile = { nullptr, 0, 0, 0 };
    const int ticks = (P.N - 1) + 2 * (W - 1) + 1;
    for (int s0 = 0; s0 < P.multi_steps; s0 += W) {
        const int nb = min(W, P.multi_steps - s0);               // steps of this batch: lanes j >= nb idle
        const bool mine = live && j < nb, first = j == 0, last = j == nb - 1;
        const int step = min(s0 + j, P.multi_steps - 1);
        const WindNow Wn = { P.wind_tab[3 * step], P.wind_tab[3 * step + 1], P.wind_tab[3 * step + 2] };
        float parx = rx, pary = ry, parz = rz;                    // this step's position of row i-1
        float lvx = 0.f, lvy = 0.f, lvz = 0.f;                    // its clamped velocity, correction pending
        float hx = 0.f, hy = 0.f, hz = 0.f, nx_ = 0.f, ny_ = 0.f, nz_ = 0.f;   // positions handed down by lane j-1: row of the next tick / the one after
        float ivx = 0.f, ivy = 0.f, ivz = 0.f;                    // final velocity handed down for the row of the next tick
        float mx = 0.f, my = 0.f, mz = 0.f, mvx = 0.f, mvy = 0.f, mvz = 0.f;   // lane 0: row prefetched from memory
        const float* rd = base + RS;
        if (first && mine) { mx = rd[0]; my = rd[PK]; mz = rd[2 * PK]; mvx = rd[3 * PK]; mvy = rd[4 * PK]; mvz = rd[5 * PK]; }
        for (int k = 1; k <= ticks; ++k) {
            const int i = k - 2 * j;
            float cx, cy, cz, vx, vy, vz;
            if (first) {
                cx = mx; cy = my; cz = mz; vx = mvx; vy = mvy; vz = mvz;
                rd += RS;
                if (mine && i + 1 < P.N) { mx = rd[0]; my = rd[PK]; mz = rd[2 * PK]; mvx = rd[3 * PK]; mvy = rd[4 * PK]; mvz = rd[5 * PK]; }
            } else { cx = hx; cy = hy; cz = hz; vx = ivx; vy = ivy; vz = ivz; }
            float opx = 0.f, opy = 0.f, opz = 0.f, ofx = lvx, ofy = lvy, ofz = lvz;   // row N: the last velocity has no correction (compute.comp:213)
            if (mine && i >= 1 && i < P.N) {
                const PointOut<float> o = point_update<float, WIND, NELL, 0>(P, Wn, no_tile, nullptr, cx, cy, cz, vx, vy, vz, parx, pary, parz);
                parx = o.px; pary = o.py; parz = o.pz;
                opx = o.px; opy = o.py; opz = o.pz;
                ofx = fmaf(o.dx, minus_inv_dt, lvx); ofy = fmaf(o.dy, minus_inv_dt, lvy); ofz = fmaf(o.dz, minus_inv_dt, lvz);   // v_{i-1} -= d_i / dt
                lvx = o.vx; lvy = o.vy; lvz = o.vz;
                if (last) {
                    float* w = base + (size_t)i * RS;
                    w[0] = o.px; w[PK] = o.py; w[2 * PK] = o.pz;
                    if (P.keep_corr) { float* c0 = corr + tiled_index(3, P.S_pad, i, 0, sc); c0[0] = o.dx; c0[PK] = o.dy; c0[2 * PK] = o.dz; }
                }
            }
            if (mine && last && i >= 2 && i <= P.N) { float* w = base + (size_t)(i - 1) * RS; w[3 * PK] = ofx; w[4 * PK] = ofy; w[5 * PK] = ofz; }
            // hand (position of row i, final velocity of row i-1) to lane j+1
            const float tpx = __shfl_up_sync(kFull, opx, 1, W), tpy = __shfl_up_sync(kFull, opy, 1, W), tpz = __shfl_up_sync(kFull, opz, 1, W);
            ivx = __shfl_up_sync(kFull, ofx, 1, W); ivy = __shfl_up_sync(kFull, ofy, 1, W); ivz = __shfl_up_sync(kFull, ofz, 1, W);
            hx = nx_; hy = ny_; hz = nz_;
            nx_ = tpx; ny_ = tpy; nz_ = tpz;
        }
        __syncwarp();                                             // the next batch's lane 0 reads what this batch's last lane wrote
    }
}

// ---- K_splat: corrected velocities -> voxel grid (compute.comp:231-252) ----------------------------
// The shader issues 32 integer atomics per point (8 corners x {vx,vy,vz,density}).  Here a warp owns 32
// neighbouring strands (Morton-ordered by root, so the 32 points of a row share their base cell or
// split over very few cells) and walks the rows root->tip in two phases per row:
//   A  lane = strand: coalesced 128-byte loads of the row (next row already in flight), exact grid
//      coordinates, per-axis weights and the base-cell key, staged in shared memory (points moving
//      faster than the int32 bound go straight to the grid on their own);
//   B  lane = (corner, slot): the 32 lanes are the 8 corners x 4 slots; in 4 iterations each lane
//      turns two (point, corner) pairs -- one fp32x2 pack -- into their four integers each, float
//      operations ordered exactly as the shader orders them (FMUL2, then F2I.TRUNC), and adds them into
//      the REGISTER accumulators of K0 or K1.  No atomic, no shuffle, no divergence.
// At the end of the row each of the (at most two) cells is finished by a 12-instruction reduce-scatter
// across the 4 slots and ONE warp-wide RED.64 with 32 distinct addresses (8 cells x 32 bytes).  Integer
// sums are exact in any order, so the grid equals the shader's bit for bit, with 32x fewer atomics and
// no shared-memory atomics or bounding boxes.
#ifndef RVH_SPLAT_THREADS
#define RVH_SPLAT_THREADS 128
#endif

constexpr int kSplatThreads = RVH_SPLAT_THREADS;
constexpr int kStageW = 40;                // words between the a=0 and a=1 weight rows: distinct banks for the quarter-warp phases of LDS.128

struct __align__(16) SplatStage {          // one row of one warp
    float w[3][2][kStageW];                // [axis][cell f / f+1][point]
    float v[3][32];                        // [component][point]
    int key[32];                           // base-cell key, -1 = nothing to add
};

// f+1 per axis in [0, G], G <= 1023: 10 bits each, always a non-negative int
__device__ __forceinline__ int splat_key(int f1x, int f1y, int f1z) { return f1x | (f1y << 10) | (f1z << 20); }

// Sum the accumulators of one cell over the 4 slots of each corner (lanes l, l^8, l^16, l^24) so that the lane
// ends up with ONE component of its corner, then add it to the grid: one RED.64 per lane, 32 distinct addresses.
__device__ __forceinline__ void splat_flush_cell(const StepParams& P, unsigned long long* __restrict__ grid, int lane, int key,
                                                 int a0, int a1, int a2, int a3) {
    constexpr unsigned kFull = 0xffffffffu;
    const bool hi = lane & 16, mid = lane & 8;
    int k0 = hi ? a2 : a0, k1 = hi ? a3 : a1;
    k0 += __shfl_xor_sync(kFull, hi ? a0 : a2, 16);
    k1 += __shfl_xor_sync(kFull, hi ? a1 : a3, 16);
    int mine = mid ? k1 : k0;
    mine += __shfl_xor_sync(kFull, mid ? k0 : k1, 8);
    const int comp = (hi ? 2 : 0) + (mid ? 1 : 0);
    const int fx = (key & 1023) - 1 + (lane & 1), fy = ((key >> 10) & 1023) - 1 + ((lane >> 1) & 1), fz = (key >> 20) - 1 + ((lane >> 2) & 1);
    if (cell_ok(fx, P.G) && cell_ok(fy, P.G) && cell_ok(fz, P.G))
        global_add(grid + 4 * (size_t)(fx + (fy + fz * P.G) * P.G) + comp, (long long)mine);
}

// the four integers of one (point, corner) pack: float operations ordered exactly as the shader orders them (compute.comp:241-248)
template <bool MAGIC>
__device__ __forceinline__ void splat_pack_ints(float2 sc2, float2 wx, float2 wy, float2 wz, float2 vx, float2 vy, float2 vz,
                                                int2& ix, int2& iy, int2& iz, int2& id) {
    const float2 tw = __fmul2_rn(__fmul2_rn(wx, wy), wz);
    const float2 cx = __fmul2_rn(sc2, __fmul2_rn(tw, vx));
    const float2 cy = __fmul2_rn(sc2, __fmul2_rn(tw, vy));
    const float2 cz = __fmul2_rn(sc2, __fmul2_rn(tw, vz));
    const float2 cd = __fmul2_rn(sc2, tw);
    ix = make_int2(__float2int_rz(cx.x), __float2int_rz(cx.y));
    iy = make_int2(__float2int_rz(cy.x), __float2int_rz(cy.y));
    iz = make_int2(__float2int_rz(cz.x), __float2int_rz(cz.y));
    if (MAGIC) {                                                   // 0 <= SCALE*w < 2^23: FADD.RZ with 2^23 on the FMA pipe, bias taken off by the caller's constant
        id = make_int2(__float_as_int(__fadd_rz(cd.x, 8388608.0f)) - 0x4B000000, __float_as_int(__fadd_rz(cd.y, 8388608.0f)) - 0x4B000000);
    } else id = make_int2(__float2int_rz(cd.x), __float2int_rz(cd.y));
}

// Rows are independent in the splat (unlike the FTL chain), so blockIdx.y splits them into chunks of
// `rows_per_chunk`: scenes with few strands (C3: 100K x 64 = 3,125 warps walking 63 rows) still fill the machine.
// MAGIC (host: grid_scale < 2^23): the density term 0 <= SCALE*w < 2^23 is truncated by FADD.RZ with 2^23 on the FMA pipe
// instead of F2I on the quarter-rate XU pipe (same integer; splat 0.478 -> 0.460 ms at 1M x 32).
// Round 2: floor(g) as float AND integer from one FADD.RM with 1.5*2^23 (no F2I / I2F / clamps in phase A); LDS.128 in
// phase B (four points per load); rows whose 32 points share one base cell (about half of them) take a loop without
// key loads, key tests or predicated adds (IADD3 sums both points of a pack).
#ifndef RVH_SPLAT_MINBLOCKS
#define RVH_SPLAT_MINBLOCKS 8
#endif
// The rows [1 + by * rows_per_chunk, ...) of the 128 strands of block bx.  COH: coherent loads of the planes (k_scene_step: the
// same launch wrote them).
template <bool MAGIC, bool COH>
__device__ __forceinline__ void splat_rows(const StepParams& P, const float* __restrict__ planes, unsigned long long* __restrict__ grid, int rows_per_chunk,
                                           int bx, int by, SplatStage (&stage)[kSplatThreads / 32][2]) {
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = P.strand0 + bx * kSplatThreads + threadIdx.x;   // S_pad is a multiple of 128: always in bounds
    const bool live = s < P.S;
    if (!__any_sync(kFull, live)) return;
    const int r0 = 1 + by * rows_per_chunk, r1 = min(P.N, r0 + rows_per_chunk);
    if (r0 >= r1) return;
    const int ca = lane & 1, cb = (lane >> 1) & 1, cc = (lane >> 2) & 1, slot = lane >> 3;
    const size_t RS = (size_t)P.S_pad * 6;
    const float* nextp = planes + tiled_index(6, P.S_pad, r0, 0, s);
    float nx[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) nx[k] = ld_plane<COH>(nextp + k * kTileStrands);
    const float2 sc2 = make_float2(P.scale, P.scale);
    for (int r = r0; r < r1; ++r) {
        float c[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) c[k] = nx[k];
        nextp += RS;
        if (r + 1 < r1) {
#pragma unroll
            for (int k = 0; k < 6; ++k) nx[k] = ld_plane<COH>(nextp + k * kTileStrands);
        }
        // ---- phase A: lane = strand -------------------------------------------------------------------
        // t = RM(g + 1.5*2^23) = 1.5*2^23 + floor(g) for |g| < 2^22: fl = t - 1.5*2^23 is floor(g) as a float and
        // bits(t) - 0x4B3FFFFF is floor(g) + 1 as an integer.  bits(t) is monotonic in g, so "0 <= floor(g)+1 <= G" holds for
        // exactly the g in [-1, G) -- far-away and NaN coordinates included -- and some corner is a cell <=> it holds on every axis.
        int f1[3]; float w0[3], w1[3];
        bool touches = live;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float g = grid_coord<float>(P, c[a], a);
            const float t = __fadd_rd(g, 12582912.0f);
            const float fl = __fsub_rn(t, 12582912.0f);
            f1[a] = __float_as_int(t) - 0x4B3FFFFF;
            touches = touches && (unsigned)f1[a] <= (unsigned)P.G;
            w0[a] = fmaf(__fsub_rn(g, fl), -1.0f, 1.0f);
            w1[a] = __fadd_rn(__fsub_rn(g, __fadd_rn(fl, 1.0f)), 1.0f);
        }
        const float vinf = fmaxf(fabsf(c[3]), fmaxf(fabsf(c[4]), fabsf(c[5])));
        bool agg = touches && vinf <= P.splat_vagg;                   // a NaN velocity fails the test
        if (P.splat_sparse) {
            // Sparse hair (few strands per cell: the shipped 900-strand scene puts a row's 32 points into 20-30 cells): the aggregation
            // below would take a pass per two cells (~170 instructions each, 6.4 us per row measured in k_scene_step); 32 atomics per
            // point, all lanes at once, are ~200 instructions for the whole row.  Same integers (64-bit conversions of the same floats).
            const int k0 = agg ? splat_key(f1[0], f1[1], f1[2]) : -1;
            const unsigned same = __match_any_sync(kFull, k0);
            const int distinct = __popc(__ballot_sync(kFull, agg && (__ffs(same) - 1 == lane)));
            if (distinct >= P.splat_sparse) agg = false;               // warp-uniform
        }
        const int key = agg ? splat_key(f1[0], f1[1], f1[2]) : -1;
        if (touches && !agg) {                                          // very fast (or NaN) point: 64-bit path, on its own
            const float wx[2] = { w0[0], w1[0] }, wy[2] = { w0[1], w1[1] }, wz[2] = { w0[2], w1[2] };
            splat_point_direct(P, grid, f1[0] - 1, f1[1] - 1, f1[2] - 1, wx, wy, wz, c[3], c[4], c[5]);
        }
        unsigned rem = __ballot_sync(kFull, agg);
        if (rem == 0u) continue;                                        // nothing of this row lands in the grid (warp-uniform)
        SplatStage& st = stage[warp][r & 1];
        st.w[0][0][lane] = w0[0]; st.w[0][1][lane] = w1[0]; st.w[1][0][lane] = w0[1]; st.w[1][1][lane] = w1[1];
        st.w[2][0][lane] = w0[2]; st.w[2][1][lane] = w1[2];
        st.v[0][lane] = c[3]; st.v[1][lane] = c[4]; st.v[2][lane] = c[5];
        st.key[lane] = key;
        __syncwarp();
        // ---- phase B: lane = (corner, slot); the slot's 8 points in two LDS.128 rounds of two fp32x2 packs -------------
        const float4* wxp = reinterpret_cast<const float4*>(st.w[0][ca]) + slot;
        const float4* wyp = reinterpret_cast<const float4*>(st.w[1][cb]) + slot;
        const float4* wzp = reinterpret_cast<const float4*>(st.w[2][cc]) + slot;
        const float4* vxp = reinterpret_cast<const float4*>(st.v[0]) + slot;
        const float4* vyp = reinterpret_cast<const float4*>(st.v[1]) + slot;
        const float4* vzp = reinterpret_cast<const float4*>(st.v[2]) + slot;
        const int4* kp = reinterpret_cast<const int4*>(st.key) + slot;
        const int K0 = __shfl_sync(kFull, key, __ffs(rem) - 1);
        if (rem == kFull && __all_sync(kFull, key == K0)) {
            // one base cell, every point aggregated (about half of all rows): no key loads, no tests, IADD3 takes both points of a pack
            int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const float4 WX = wxp[4 * it], WY = wyp[4 * it], WZ = wzp[4 * it], VX = vxp[4 * it], VY = vyp[4 * it], VZ = vzp[4 * it];
                int2 ix, iy, iz, id;
                splat_pack_ints<MAGIC>(sc2, make_float2(WX.x, WX.y), make_float2(WY.x, WY.y), make_float2(WZ.x, WZ.y), make_float2(VX.x, VX.y), make_float2(VY.x, VY.y), make_float2(VZ.x, VZ.y), ix, iy, iz, id);
                a0 += ix.x + ix.y; a1 += iy.x + iy.y; a2 += iz.x + iz.y; a3 += id.x + id.y;
                splat_pack_ints<MAGIC>(sc2, make_float2(WX.z, WX.w), make_float2(WY.z, WY.w), make_float2(WZ.z, WZ.w), make_float2(VX.z, VX.w), make_float2(VY.z, VY.w), make_float2(VZ.z, VZ.w), ix, iy, iz, id);
                a0 += ix.x + ix.y; a1 += iy.x + iy.y; a2 += iz.x + iz.y; a3 += id.x + id.y;
            }
            splat_flush_cell(P, grid, lane, K0, a0, a1, a2, a3);
            continue;
        }
        // two cells of the row per pass (a row that straddles more cells takes another pass)
        while (rem) {
            const int Ka = __shfl_sync(kFull, key, __ffs(rem) - 1);
            rem &= ~__ballot_sync(kFull, key == Ka);
            int Kb = -1;
            if (rem) { Kb = __shfl_sync(kFull, key, __ffs(rem) - 1); rem &= ~__ballot_sync(kFull, key == Kb); }
            int a0 = 0, a1 = 0, a2 = 0, a3 = 0;          // cell Ka
            int e0 = 0, e1 = 0, e2 = 0, e3 = 0;          // cell Kb
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const float4 WX = wxp[4 * it], WY = wyp[4 * it], WZ = wzp[4 * it], VX = vxp[4 * it], VY = vyp[4 * it], VZ = vzp[4 * it];
                const int4 KK = kp[4 * it];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    int2 ix, iy, iz, id;
                    // key -1 points may carry garbage weights: they match neither Ka nor Kb (both != -1 when used)
                    if (hf == 0) splat_pack_ints<MAGIC>(sc2, make_float2(WX.x, WX.y), make_float2(WY.x, WY.y), make_float2(WZ.x, WZ.y), make_float2(VX.x, VX.y), make_float2(VY.x, VY.y), make_float2(VZ.x, VZ.y), ix, iy, iz, id);
                    else         splat_pack_ints<MAGIC>(sc2, make_float2(WX.z, WX.w), make_float2(WY.z, WY.w), make_float2(WZ.z, WZ.w), make_float2(VX.z, VX.w), make_float2(VY.z, VY.w), make_float2(VZ.z, VZ.w), ix, iy, iz, id);
                    const int kx = hf ? KK.z : KK.x, ky = hf ? KK.w : KK.y;
                    if (kx == Ka) { a0 += ix.x; a1 += iy.x; a2 += iz.x; a3 += id.x; }
                    if (ky == Ka) { a0 += ix.y; a1 += iy.y; a2 += iz.y; a3 += id.y; }
                    if (Kb != -1) {
                        if (kx == Kb) { e0 += ix.x; e1 += iy.x; e2 += iz.x; e3 += id.x; }
                        if (ky == Kb) { e0 += ix.y; e1 += iy.y; e2 += iz.y; e3 += id.y; }
                    }
                }
            }
            // ---- end of row: accumulators -> grid ---------------------------------------------------------
            splat_flush_cell(P, grid, lane, Ka, a0, a1, a2, a3);
            if (Kb != -1) splat_flush_cell(P, grid, lane, Kb, e0, e1, e2, e3);
        }
    }
}

template <bool MAGIC>
__global__ void __launch_bounds__(kSplatThreads, RVH_SPLAT_MINBLOCKS)
k_grid_splat(const __grid_constant__ StepParams P, const float* __restrict__ planes, unsigned long long* __restrict__ grid, int rows_per_chunk) {
    __shared__ SplatStage stage[kSplatThreads / 32][2];
    splat_rows<MAGIC, false>(P, planes, grid, rows_per_chunk, blockIdx.x, blockIdx.y, stage);
}

// ---- grid finalize: int64 accumulators -> float cells for the gather ---------------------------
__global__ void __launch_bounds__(256)
k_grid_finalize(const long long* __restrict__ grid, float4* __restrict__ fgrid, int cells, int int32_wrap) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cells) return;
    const longlong2* c = reinterpret_cast<const longlong2*>(grid + 4 * (size_t)k);
    fgrid[k] = finalize_cell(c[0], c[1], int32_wrap);
}

// ---- small scenes: the WHOLE step -- several steps -- in one persistent cooperative launch ---------------------------------
// The reference's shipped scene (900 strands x 10 points) and anything of that order is bound by launches and by dependent
// latencies, not by bytes: replayed as a CUDA graph its 4 kernels take 28 us per step, 12 of them the FTL chain with the fused
// gather (an L2 round trip and ~100 dependent instructions per row sit on the root->tip chain).  Here one co-resident grid of
// CTAs walks the phases of a step with grid-wide barriers in between (scene_barrier), and the gather leaves the chain:
//   1  CTAs [0, k1_ctas): integrate + collide + FTL WITHOUT gather (ftl_walk, as k_ftl_step<..., GATHER = 0>);
//      all other CTAs meanwhile clear the int64 accumulators (Renderer.cpp:2063) -- the clear hides under the FTL chain;
//   2  every CTA: splat (splat_rows over (strand block, row chunk) items, as k_grid_splat);
//   3  every CTA: gather + friction, one thread per point, i.e. fully parallel, reading the int64
//      accumulators directly (ld_cell<2> converts a cell as k_grid_finalize would: no finalize pass, no float grid).
// Later phases read what earlier phases of the SAME launch wrote from other SMs (planes, grid): every such load is a coherent
// ld.global (COH / SRC flavours above), never the read-only path; the grid barrier's fence orders them.  The time-only wind scalars
// of each step come from P.wind_tab (host-evaluated, as MULTI).  Entry condition (host): no gather pending.  On exit the
// accumulators hold the last step's grid (rvh_download_grid), the velocities are final (gather applied) and the float grid is stale.
// Grid-wide barrier of a cooperative launch (all CTAs co-resident): one release-add per CTA on a monotonic counter, thread 0 spins
// with acquire loads until the counter reaches this barrier's target.  The bar.sync on either side extends the release / acquire
// to the CTA's other threads (cumulativity); the acquire invalidates the SM's L1, so the weak loads that follow see the other
// CTAs' writes.  The host sizes the launch to the scene (tens of CTAs, not 148 x k): the barrier costs ~1 us there, where
// cooperative_groups' grid.sync() over 296 CTAs was measured at ~6 us (profiles/r02 summary).
__device__ __forceinline__ void scene_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
        unsigned v, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (++spins > (1u << 25)) __trap();                         // ~20 s: the launch is broken (a co-resident CTA never arrives); a trap is better than a hung GPU
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}

template <int V, bool WIND, int NELL>
__global__ void __launch_bounds__(kBlock, 2)
k_scene_step(const __grid_constant__ StepParams P, float* planes, float* corr, unsigned long long* grid,
             int nsteps, int k1_ctas, int splat_bx, int splat_items, int rows_per_chunk, unsigned* bar_counter, unsigned bar_base,
             unsigned long long* timing) {                         // timing (RVH_SCENE_TIMING, tuning): CTA 0's %globaltimer at every phase boundary
    static_assert(kSplatThreads == kBlock, "one block shape for all phases");
    unsigned bar_target = bar_base;
    int tick = 0;
    auto stamp = [&]() {
        if (timing && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); timing[tick++] = t; }
    };
    stamp();
    __shared__ SplatStage stage[kSplatThreads / 32][2];
    __shared__ __align__(128) unsigned char sdf_raw[16];
    SdfStageSmem& sm = *reinterpret_cast<SdfStageSmem*>(sdf_raw);
    const int cells = P.G * P.G * P.G;
    constexpr int plane = kTileStrands;
    const int points = (P.N - 1) * P.S_pad;
    for (int step = 0; step < nsteps; ++step) {
        if ((int)blockIdx.x < k1_ctas) {
            ftl_walk<V, WIND, NELL, 0, true, true>(P, planes, corr, nullptr, nullptr, sm, blockIdx.x, 1, step);
        } else {
            uint4* gc = reinterpret_cast<uint4*>(grid);
            const unsigned n = 2u * (unsigned)cells, stride = (gridDim.x - k1_ctas) * blockDim.x;
            for (unsigned k = (blockIdx.x - k1_ctas) * blockDim.x + threadIdx.x; k < n; k += stride) gc[k] = make_uint4(0u, 0u, 0u, 0u);
        }
        stamp();
        scene_barrier(bar_counter, bar_target += gridDim.x);
        stamp();
        for (int item = blockIdx.x; item < splat_items; item += gridDim.x) {
            splat_rows<true, true>(P, planes, grid, rows_per_chunk, item % splat_bx, item / splat_bx, stage);
            __syncwarp();                                               // the next item reuses this warp's stage buffers
        }
        stamp();
        scene_barrier(bar_counter, bar_target += gridDim.x);
        stamp();
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < points; k += gridDim.x * blockDim.x) {
            const int row = k / P.S_pad, s0 = k - row * P.S_pad;       // one thread per point: the phase is a latency chain, not bytes
            if (s0 >= P.S) continue;
            float* q = planes + tiled_index(6, P.S_pad, row + 1, 0, s0);     // skip the root row
            float px = ld_plane<true>(q), py = ld_plane<true>(q + plane), pz = ld_plane<true>(q + 2 * plane);
            float vx = ld_plane<true>(q + 3 * plane), vy = ld_plane<true>(q + 4 * plane), vz = ld_plane<true>(q + 5 * plane);
            unsigned cand;
            gather_pack<float, false, false, 2>(P, reinterpret_cast<const float4*>(grid), px, py, pz, vx, vy, vz, cand);
            q[3 * plane] = vx; q[4 * plane] = vy; q[5 * plane] = vz;
        }
        stamp();
        if (step + 1 < nsteps) scene_barrier(bar_counter, bar_target += gridDim.x);   // the end of the launch is the last step's barrier
        stamp();
    }
}

// ---- multi-GPU: fused grid exchange over NVLink peer memory ---------------------------------------
// Replaces  ncclAllReduce(int64 grid)  +  k_grid_finalize  by ONE kernel per rank.  Every rank's raw
// accumulators and its float grid are mapped into all peers (CUDA IPC over NVSwitch).  Rank r owns the
// cell slice [r*C/R, (r+1)*C/R): it LOADS that slice of every peer's int64 accumulators straight over
// NVLink (reduce-scatter by pull), sums -- integers, so the result is the same on every rank count --,
// converts to the gather's float4 cell exactly like k_grid_finalize, and STORES the finished cell into
// every peer's fgrid (all-gather by push, 16 B/cell instead of the 32 B/cell an int64 all-reduce moves).
// Per rank and step: 7/8 * 8 MiB pulled + 7/8 * 4 MiB pushed at G = 64, R = 8.
// Two flag barriers in peer memory order it: (1) nobody reads a peer's accumulators before that peer's
// splat kernel is finished (the peer raises its flag from THIS kernel, which is stream-ordered after its
// splat); (2) nobody leaves the kernel before every peer has finished reading its accumulators and writing
// its fgrid, so the next step's clear and gather are safe.  Flags carry the step epoch (monotonic).
constexpr int kMaxRanks = 8;
struct ExchangePeers {
    const long long* grid[kMaxRanks];      // peers' raw int64 accumulators (read)
    float4* fgrid[kMaxRanks];              // peers' float grids (written)
    unsigned* flags[kMaxRanks];            // peers' flag blocks: [2][kMaxRanks] epochs + [1] local block counter + [1] error word (kXErrWord)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// Every wait on a peer is bounded: a rank that died, or never reached this step, must surface as an error through the ABI
// (RVH_ERR_NCCL from the next rvh_sync / read-back), not hang every GPU of the box.  On a timeout the waiting thread records
// 1 + the peer's rank in the error word of its own flag block and carries on (the step's results are then meaningless).
constexpr int kXErrWord = 2 * kMaxRanks + 1;
__device__ __forceinline__ void wait_epoch(const unsigned* flag, unsigned epoch, long long timeout_cycles, unsigned* err, unsigned who) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        if (clock64() - t0 > timeout_cycles) { atomicMax(err, who + 1u); return; }
    }
}

__global__ void __launch_bounds__(256)
k_grid_exchange(const ExchangePeers X, int rank, int nranks, int cells, int int32_wrap, unsigned epoch, long long timeout_cycles) {
    unsigned* my = X.flags[rank];
    // ---- barrier 1: every rank's splat is complete -----------------------------------------------
    if (blockIdx.x == 0 && threadIdx.x < nranks) st_release_sys(X.flags[threadIdx.x] + rank, epoch);
    if (threadIdx.x < nranks) wait_epoch(my + threadIdx.x, epoch, timeout_cycles, my + kXErrWord, threadIdx.x);
    __syncthreads();
    // ---- reduce my slice over all peers, finalize, broadcast ---------------------------------------
    const int per = (cells + nranks - 1) / nranks;
    const int lo = rank * per, hi = min(cells, lo + per);
    for (int k = lo + blockIdx.x * blockDim.x + threadIdx.x; k < hi; k += gridDim.x * blockDim.x) {
        // all peers' loads are issued before the first use: one NVLink round trip per cell, not one per peer
        longlong2 a[kMaxRanks], b[kMaxRanks];
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) {
            if (r < nranks) {
                const longlong2* c = reinterpret_cast<const longlong2*>(X.grid[r] + 4 * (size_t)k);
                a[r] = c[0]; b[r] = c[1];
            }
        }
        long long v0 = 0, v1 = 0, v2 = 0, dens = 0;
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) {
            if (r < nranks) { v0 += a[r].x; v1 += a[r].y; v2 += b[r].x; dens += b[r].y; }
        }
        if (int32_wrap) { dens = (int)dens; v0 = (int)v0; v1 = (int)v1; v2 = (int)v2; }
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dens > 0) {
            const float rd = __frcp_rn(__ll2float_rn(dens));
            o.x = __ll2float_rn(v0) * rd; o.y = __ll2float_rn(v1) * rd; o.z = __ll2float_rn(v2) * rd; o.w = __ll2float_rn(dens);
        }
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) if (r < nranks) X.fgrid[r][k] = o;
    }
    // ---- barrier 2: all ranks done reading accumulators and writing fgrid ------------------------------
    __threadfence_system();
    __syncthreads();
    unsigned* counter = my + 2 * kMaxRanks;
    if (threadIdx.x == 0) atomicAdd(counter, 1u);
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            while (atomicAdd(counter, 0u) < gridDim.x) { }            // every block of this rank has pushed its cells
            *counter = 0u;
            __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x < nranks) {
            st_release_sys(X.flags[threadIdx.x] + kMaxRanks + rank, epoch);
            wait_epoch(my + kMaxRanks + threadIdx.x, epoch, timeout_cycles, my + kXErrWord, threadIdx.x);
        }
    }
}

// ---- K2: stand-alone grid gather + friction ------------------------------------------------------
// Normally the gather of step k rides in k_ftl_step of step k+1; this kernel applies it when the state
// is read back (download / interop pack) or a phase is requested explicitly.  One thread = one pack.
template <bool REP>
__global__ void __launch_bounds__(256)
k_grid_gather(const __grid_constant__ StepParams P, float* __restrict__ planes, const float4* __restrict__ fgrid) {
    constexpr int plane = kTileStrands;
    const size_t half = (size_t)P.S_pad / 2;
    const size_t total = (size_t)(P.N - 1) * half;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const size_t row = k / half, s0 = 2 * (k - row * half);
        if (s0 >= (size_t)P.S) continue;
        float* q = planes + tiled_index(6, P.S_pad, (int)row + 1, 0, (int)s0);   // skip the root row
        const float2 px = *reinterpret_cast<const float2*>(q), py = *reinterpret_cast<const float2*>(q + plane), pz = *reinterpret_cast<const float2*>(q + 2 * plane);
        float2 vx = *reinterpret_cast<const float2*>(q + 3 * plane), vy = *reinterpret_cast<const float2*>(q + 4 * plane), vz = *reinterpret_cast<const float2*>(q + 5 * plane);
        unsigned cand;
        gather_pack<float2, REP>(P, fgrid, px, py, pz, vx, vy, vz, cand);
        *reinterpret_cast<float2*>(q + 3 * plane) = vx; *reinterpret_cast<float2*>(q + 4 * plane) = vy; *reinterpret_cast<float2*>(q + 5 * plane) = vz;
    }
}

// StrandDrawIndirect of the imported indirect-args buffer (Strand.h:53-58): {vertexCount = S, instanceCount = 1, 0, 0}.  The shader
// zeroes vertexCount and lets every invocation add 1 (compute.comp:126-130, 302); the value it ends with is the invocation count.
__global__ void k_write_indirect(uint32_t* __restrict__ out, uint32_t strands) {
    if (threadIdx.x < 4) out[threadIdx.x] = threadIdx.x == 0 ? strands : (threadIdx.x == 1 ? 1u : 0u);
}

// ---- test hook: the collider decisions of compute.comp:158-180 as THIS path takes them ------------------------------------
// One byte per point (external strand order, [S][N]): bit 0 = inside the sphere (:162), bit j = inside ellipsoid j (:66), or,
// with the head SDF, bit 1 = trilinear distance < 0.  Same device functions, same operations as point_update, so the byte is
// the decision k_ftl_step takes for the positions currently in `planes`; tests count how often it differs from the oracle's
// (squared-distance compare, FMA contraction: a point within an ulp of a collider surface can flip).
__global__ void __launch_bounds__(256)
k_hit_masks(const __grid_constant__ StepParams P, const float* __restrict__ planes, const int* __restrict__ perm, unsigned char* __restrict__ out, int sdf_on) {
    const size_t total = (size_t)P.N * P.S;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(k / P.S), s = (int)(k - (size_t)row * P.S);
        const size_t e = perm ? (size_t)perm[s] : (size_t)s;
        unsigned m = 0u;
        if (row > 0) {
            const float cx = planes[tiled_index(6, P.S_pad, row, 0, s)], cy = planes[tiled_index(6, P.S_pad, row, 1, s)], cz = planes[tiled_index(6, P.S_pad, row, 2, s)];
            const float dx = cx - P.sphere_c[0], dy = cy - P.sphere_c[1], dz = cz - P.sphere_c[2];
            if (P.has_sphere && vdot3(dx, dy, dz, dx, dy, dz) < P.sphere_r2) m |= 1u;
            if (sdf_on) {
                const SdfTile none = { nullptr, 0, 0, 0 };
                if (sdf_inside(P.sdf, none, cx, cy, cz)) m |= 2u;
            } else {
                for (int j = 0; j < P.n_ell; ++j) {
                    float qx, qy, qz;
                    if (ellipsoid_q<float>(P.ell[j], cx, cy, cz, qx, qy, qz) <= 1.0f) m |= 2u << j;
                }
            }
        }
        out[e * P.N + row] = (unsigned char)m;
    }
}

// ---- AoS <-> planes (the reference's Strand[S] vertex-buffer layout, Strand.h:11-15) -------
// One block per tile of 32 strands; shared-memory transpose so both sides are coalesced.
// perm[s_internal] = external strand index (nullptr = identity).
constexpr int kTile = 32;

__global__ void __launch_bounds__(256)
k_unpack_aos(const float4* __restrict__ aos, float* __restrict__ planes, const int* __restrict__ perm,
             int S, int S_pad, int N, float rest, int first_tile) {
    extern __shared__ float sm[];            // [6][N][kTile+1]
    const int tile0 = (first_tile + blockIdx.x) * kTile;
    const int q2 = 2 * N;
    for (int k = threadIdx.x; k < kTile * q2; k += blockDim.x) {
        const int sl = k / q2, q = k % q2;
        const int s = tile0 + sl;
        float4 v;
        if (s < S) {
            const size_t e = perm ? (size_t)perm[s] : (size_t)s;
            v = aos[e * 3 * N + q];
        } else {
            // padding strand: straight, at rest spacing, far outside the grid and colliders
            const int j = q < N ? q : q - N;
            v = q < N ? make_float4(1.0e4f + rest * j, 1.0e4f, 1.0e4f, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int j = q < N ? q : q - N;
        const int kb = q < N ? 0 : 3;
        sm[((kb + 0) * N + j) * (kTile + 1) + sl] = v.x;
        sm[((kb + 1) * N + j) * (kTile + 1) + sl] = v.y;
        sm[((kb + 2) * N + j) * (kTile + 1) + sl] = v.z;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 6 * N * kTile; k += blockDim.x) {
        const int sl = k % kTile, r = k / kTile;     // r = kk*N + j
        const int kk = r / N, j = r % N;
        if (tile0 + sl < S_pad) planes[tiled_index(6, S_pad, j, kk, tile0 + sl)] = sm[r * (kTile + 1) + sl];
    }
}

__global__ void __launch_bounds__(256)
k_pack_aos(float4* __restrict__ aos, const float* __restrict__ planes, const float* __restrict__ corr,
           const int* __restrict__ perm, int S, int S_pad, int N, int first_tile, int which) {
    // which: bit 0 curvePoints, bit 1 curveVels, bit 2 correctionVecs -- only these thirds of every Strand are (read and) written
    extern __shared__ float sm[];            // [9][N][kTile+1]
    const int tile0 = (first_tile + blockIdx.x) * kTile;
    const int nk = corr ? 9 : 6;
    for (int k = threadIdx.x; k < nk * N * kTile; k += blockDim.x) {
        const int sl = k % kTile, r = k / kTile;
        const int kk = r / N, j = r % N;
        float v = 0.f;
        if (tile0 + sl < S_pad && ((which >> (kk / 3)) & 1))
            v = kk < 6 ? planes[tiled_index(6, S_pad, j, kk, tile0 + sl)]
                       : corr[tiled_index(3, S_pad, j, kk - 6, tile0 + sl)];
        sm[r * (kTile + 1) + sl] = v;
    }
    __syncthreads();
    const int q3 = 3 * N;
    for (int k = threadIdx.x; k < kTile * q3; k += blockDim.x) {
        const int sl = k / q3, q = k % q3;
        const int s = tile0 + sl;
        if (s >= S) continue;
        const int a = q / N, j = q % N;      // a: 0 curvePoints, 1 curveVels, 2 correctionVecs
        if (!((which >> a) & 1)) continue;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a < 2 || corr) {
            v.x = sm[((3 * a + 0) * N + j) * (kTile + 1) + sl];
            v.y = sm[((3 * a + 1) * N + j) * (kTile + 1) + sl];
            v.z = sm[((3 * a + 2) * N + j) * (kTile + 1) + sl];
        }
        if (a == 0) v.w = 1.0f;              // curvePoints.w = 1 (Strand.cpp:165, compute.comp:196)
        if (a == 2 && j == 0) v = make_float4(0.f, 0.f, 0.f, 0.f);   // correctionVecs[0] is never written
        const size_t e = perm ? (size_t)perm[s] : (size_t)s;
        aos[e * q3 + q] = v;
    }
}

// ---- GPU scene init: the seeded synthetic head (host twin: scenes.synthetic_head) ------
// Roots on the top hemisphere of the head ellipsoid (collider 1), counter-based splitmix64 keyed by the GLOBAL
// strand id (any rank generates exactly its shard, no host traffic), points at exact rest spacing along
// normalize(n + 0.1*(0.05, 5, -2)), velocity (0, 0, -1).  Writes the Strand[S] AoS staging buffer, so the
// normal unpack (+ Morton ordering) path follows.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
k_synth_head_aos(float4* __restrict__ aos, int S, int N, unsigned long long first_strand, unsigned long long seed, float rest,
                 const __grid_constant__ Ellipsoid head) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const unsigned long long base = (seed << 32) + 2ull * (first_strand + (unsigned long long)s);
    const float r0 = (float)(splitmix64(base) >> 40) * (1.0f / 16777216.0f);
    const float r1 = (float)(splitmix64(base + 1ull) >> 40) * (1.0f / 16777216.0f);
    const float uy = r0, rad = sqrtf(fmaxf(1.0f - uy * uy, 0.0f)), phi = 6.2831855f * r1;
    const float ux = rad * cosf(phi), uz = rad * sinf(phi);
    const float rx = head.xf[0] * ux + head.xf[1] * uy + head.xf[2] * uz + head.xf[3];
    const float ry = head.xf[4] * ux + head.xf[5] * uy + head.xf[6] * uz + head.xf[7];
    const float rz = head.xf[8] * ux + head.xf[9] * uy + head.xf[10] * uz + head.xf[11];
    float nx = head.nt[0] * ux + head.nt[1] * uy + head.nt[2] * uz;
    float ny = head.nt[3] * ux + head.nt[4] * uy + head.nt[5] * uz;
    float nz = head.nt[6] * ux + head.nt[7] * uy + head.nt[8] * uz;
    const float inl = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
    float dx = nx * inl + 0.1f * 0.05f, dy = ny * inl + 0.1f * 5.0f, dz = nz * inl + 0.1f * -2.0f;
    const float idl = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    dx *= idl; dy *= idl; dz *= idl;
    float4* cp = aos + (size_t)s * 3 * N;
    for (int j = 0; j < N; ++j) {
        const float t = (float)j * rest;
        cp[j] = make_float4(rx + t * dx, ry + t * dy, rz + t * dz, 1.0f);
        cp[N + j] = make_float4(0.0f, 0.0f, -1.0f, 0.0f);
    }
}

// ---- GPU scene init from a triangle mesh ---------------------------------------------------
// Follicle placement of Strand.cpp:88-110 on the device, with the fix its own TODO asks for ("account for differences in
// triangle area", Strand.cpp:92): the triangle is drawn from the area CDF (binary search) instead of uniformly; the
// barycentric fold (Strand.cpp:100-103) is kept; the normal is the interpolated corner normal (face normal when the mesh
// has none).  Counter-based splitmix64 keyed by the GLOBAL strand id, three draws per strand, so every rank generates
// exactly its shard.  Strands leave the surface along normalize(n + 0.1*(0.05, 5, -2)) at exact rest spacing with
// velocity (0,0,-1) (Strand.cpp:166), like the synthetic head.  Host twin: scenes.mesh_head.
__global__ void __launch_bounds__(256)
k_mesh_follicles_aos(float4* __restrict__ aos, int S, int N, unsigned long long first_strand, unsigned long long seed, float rest,
                     const float* __restrict__ tri_pos, const float* __restrict__ tri_nrm, const float* __restrict__ cdf, int ntris) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const unsigned long long base = (seed << 32) + 3ull * (first_strand + (unsigned long long)s);
    const float r0 = (float)(splitmix64(base) >> 40) * (1.0f / 16777216.0f);
    float u = (float)(splitmix64(base + 1ull) >> 40) * (1.0f / 16777216.0f);
    float v = (float)(splitmix64(base + 2ull) >> 40) * (1.0f / 16777216.0f);
    int lo = 0, hi = ntris - 1;                               // smallest t with cdf[t] > r0
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(cdf + mid) > r0) hi = mid; else lo = mid + 1; }
    if (__fadd_rn(u, v) >= 1.0f) { u = __fsub_rn(1.0f, u); v = __fsub_rn(1.0f, v); }
    const float w = __fsub_rn(__fsub_rn(1.0f, u), v);
    const float* A = tri_pos + 9 * (size_t)lo;
    float root[3], n[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) root[a] = __fadd_rn(__fadd_rn(__fmul_rn(A[a], w), __fmul_rn(A[3 + a], u)), __fmul_rn(A[6 + a], v));
    if (tri_nrm) {
        const float* Nn = tri_nrm + 9 * (size_t)lo;
#pragma unroll
        for (int a = 0; a < 3; ++a) n[a] = __fadd_rn(__fadd_rn(__fmul_rn(Nn[a], w), __fmul_rn(Nn[3 + a], u)), __fmul_rn(Nn[6 + a], v));
    } else {
        const float e1[3] = { A[3] - A[0], A[4] - A[1], A[5] - A[2] }, e2[3] = { A[6] - A[0], A[7] - A[1], A[8] - A[2] };
        n[0] = e1[1] * e2[2] - e1[2] * e2[1]; n[1] = e1[2] * e2[0] - e1[0] * e2[2]; n[2] = e1[0] * e2[1] - e1[1] * e2[0];
    }
    const float inl = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    float dx = n[0] * inl + 0.1f * 0.05f, dy = n[1] * inl + 0.1f * 5.0f, dz = n[2] * inl + 0.1f * -2.0f;
    const float idl = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    dx *= idl; dy *= idl; dz *= idl;
    float4* cp = aos + (size_t)s * 3 * N;
    for (int j = 0; j < N; ++j) {
        const float t = (float)j * rest;
        cp[j] = make_float4(root[0] + t * dx, root[1] + t * dy, root[2] + t * dz, 1.0f);
        cp[N + j] = make_float4(0.0f, 0.0f, -1.0f, 0.0f);
    }
}

// Morton key (10 bits per axis over the grid box) of each strand's root, for spatial ordering.
__device__ __forceinline__ unsigned part1by2(unsigned x) {
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
__global__ void k_morton_keys(const float4* __restrict__ aos, int S, int N, float ox, float oy, float oz,
                              float inv_extent, unsigned* __restrict__ keys, int* __restrict__ ids) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float4 r = aos[(size_t)s * 3 * N];
    const float fx = fminf(fmaxf((r.x - ox) * inv_extent, 0.f), 0.999999f) * 1024.f;
    const float fy = fminf(fmaxf((r.y - oy) * inv_extent, 0.f), 0.999999f) * 1024.f;
    const float fz = fminf(fmaxf((r.z - oz) * inv_extent, 0.f), 0.999999f) * 1024.f;
    keys[s] = part1by2((unsigned)fx) | (part1by2((unsigned)fy) << 1) | (part1by2((unsigned)fz) << 2);
    ids[s] = s;
}


// ---- head SDF bakes (extension; oracle twins: oracle.c orc_sdf_bake_colliders / orc_sdf_bake_mesh) ----------------
// One thread per lattice node, node (i,j,k) at origin + cell*(i,j,k); padded columns i >= nx get a large positive value.
constexpr float kSdfFar = 1.0e9f;

// Signed radial distance to the union of the ellipsoid colliders: for ellipsoid E, q = inv*(p,1), on = T*(q/|q|, 1),
// |on - p| is the penetration depth the shader uses (compute.comp:171-172); negative where |q| <= 1.
__global__ void __launch_bounds__(256)
k_sdf_bake_colliders(const __grid_constant__ StepParams P, float* __restrict__ out, int nx, int ny, int nz, int nxp,
                     float ox, float oy, float oz, float cell) {
    const size_t total = (size_t)nxp * ny * nz;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const int i = (int)(k % nxp), j = (int)((k / nxp) % ny), kk = (int)(k / ((size_t)nxp * ny));
    if (i >= nx) { out[k] = kSdfFar; return; }
    const float px = ox + cell * (float)i, py = oy + cell * (float)j, pz = oz + cell * (float)kk;
    float best = kSdfFar;
    for (int e = 0; e < P.n_ell; ++e) {
        const Ellipsoid& E = P.ell[e];
        float qx, qy, qz;
        const float q2 = ellipsoid_q<float>(E, px, py, pz, qx, qy, qz);
        float ux = 1.f, uy = 0.f, uz = 0.f;
        if (q2 > 0.f) { const float r = 1.0f / sqrtf(q2); ux = qx * r; uy = qy * r; uz = qz * r; }
        const float sx = fmaf(E.xf[0], ux, fmaf(E.xf[1], uy, fmaf(E.xf[2], uz, E.xf[3])));
        const float sy = fmaf(E.xf[4], ux, fmaf(E.xf[5], uy, fmaf(E.xf[6], uz, E.xf[7])));
        const float sz = fmaf(E.xf[8], ux, fmaf(E.xf[9], uy, fmaf(E.xf[10], uz, E.xf[11])));
        const float ex = px - sx, ey = py - sy, ez = pz - sz;
        const float dist = sqrtf(fmaf(ex, ex, fmaf(ey, ey, ez * ez)));
        best = fminf(best, q2 <= 1.0f ? -dist : dist);
    }
    out[k] = best;
}

// Squared distance from p to triangle (a, b, c): closest-point regions of Ericson, Real-Time Collision Detection 5.1.5.
__device__ __forceinline__ float tri_dist2(float px, float py, float pz, const float* t) {
    const float ax = t[0], ay = t[1], az = t[2];
    const float abx = t[3] - ax, aby = t[4] - ay, abz = t[5] - az;
    const float acx = t[6] - ax, acy = t[7] - ay, acz = t[8] - az;
    const float apx = px - ax, apy = py - ay, apz = pz - az;
    const float d1 = abx * apx + aby * apy + abz * apz, d2 = acx * apx + acy * apy + acz * apz;
    float cx, cy, cz;                                    // closest point - a
    if (d1 <= 0.f && d2 <= 0.f) { cx = 0.f; cy = 0.f; cz = 0.f; }
    else {
        const float bpx = apx - abx, bpy = apy - aby, bpz = apz - abz;
        const float d3 = abx * bpx + aby * bpy + abz * bpz, d4 = acx * bpx + acy * bpy + acz * bpz;
        if (d3 >= 0.f && d4 <= d3) { cx = abx; cy = aby; cz = abz; }
        else {
            const float vc = d1 * d4 - d3 * d2;
            if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { const float v = d1 / (d1 - d3); cx = v * abx; cy = v * aby; cz = v * abz; }
            else {
                const float cpx = apx - acx, cpy = apy - acy, cpz = apz - acz;
                const float d5 = abx * cpx + aby * cpy + abz * cpz, d6 = acx * cpx + acy * cpy + acz * cpz;
                if (d6 >= 0.f && d5 <= d6) { cx = acx; cy = acy; cz = acz; }
                else {
                    const float vb = d5 * d2 - d1 * d6;
                    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { const float w = d2 / (d2 - d6); cx = w * acx; cy = w * acy; cz = w * acz; }
                    else {
                        const float va = d3 * d6 - d5 * d4;
                        if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
                            const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
                            cx = abx + w * (acx - abx); cy = aby + w * (acy - aby); cz = abz + w * (acz - abz);
                        } else {
                            const float den = 1.0f / (va + vb + vc);
                            const float v = vb * den, w = vc * den;
                            cx = abx * v + acx * w; cy = aby * v + acy * w; cz = abz * v + acz * w;
                        }
                    }
                }
            }
        }
    }
    const float ex = apx - cx, ey = apy - cy, ez = apz - cz;
    return ex * ex + ey * ey + ez * ez;
}

// Solid angle of the triangle seen from p (Van Oosterom & Strackee 1983): 2*atan2(a.(b x c), |a||b||c| + (a.b)|c| + (b.c)|a| + (c.a)|b|).
__device__ __forceinline__ float tri_solid_angle(float px, float py, float pz, const float* t) {
    const float ax = t[0] - px, ay = t[1] - py, az = t[2] - pz;
    const float bx = t[3] - px, by = t[4] - py, bz = t[5] - pz;
    const float cx = t[6] - px, cy = t[7] - py, cz = t[8] - pz;
    const float la = sqrtf(ax * ax + ay * ay + az * az), lb = sqrtf(bx * bx + by * by + bz * bz), lc = sqrtf(cx * cx + cy * cy + cz * cz);
    const float num = ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
    const float den = la * lb * lc + (ax * bx + ay * by + az * bz) * lc + (bx * cx + by * cy + bz * cz) * la + (cx * ax + cy * ay + cz * az) * lb;
    return 2.0f * atan2f(num, den);
}

// Brute force over all triangles (staged through shared memory 128 at a time): the mannequin has a few thousand
// triangles and the lattice ~1M nodes, i.e. a few G point-triangle tests, tens of milliseconds on a B200, once.
// Sign: |generalized winding number| > 1/2 (Jacobson et al. 2013) -- well defined for the open-necked mesh.
constexpr int kBakeTris = 128;
__global__ void __launch_bounds__(256)
k_sdf_bake_mesh(float* __restrict__ out, int nx, int ny, int nz, int nxp, float ox, float oy, float oz, float cell,
                const float* __restrict__ tri9, int ntris) {
    __shared__ float st[kBakeTris * 9];
    const size_t total = (size_t)nxp * ny * nz;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = k < total;
    const int i = live ? (int)(k % nxp) : 0, j = live ? (int)((k / nxp) % ny) : 0, kk = live ? (int)(k / ((size_t)nxp * ny)) : 0;
    const float px = ox + cell * (float)i, py = oy + cell * (float)j, pz = oz + cell * (float)kk;
    float best2 = 3.0e38f, omega = 0.f;
    for (int t0 = 0; t0 < ntris; t0 += kBakeTris) {
        const int cnt = min(kBakeTris, ntris - t0);
        __syncthreads();
        for (int q = threadIdx.x; q < cnt * 9; q += blockDim.x) st[q] = tri9[(size_t)t0 * 9 + q];
        __syncthreads();
        for (int t = 0; t < cnt; ++t) {
            best2 = fminf(best2, tri_dist2(px, py, pz, st + 9 * t));
            omega += tri_solid_angle(px, py, pz, st + 9 * t);
        }
    }
    if (!live) return;
    const float d = sqrtf(best2);
    out[k] = i >= nx ? kSdfFar : (fabsf(omega) > 6.2831855f ? -d : d);     // |omega / 4pi| > 1/2
}

// ---- guide strand -> render strands: hair.tesc + hair.tese on the tessellator ---------------
// The reference draws every guide strand as ONE patch that the fixed-function tessellator expands into 12 isolines x 42
// segments (hair.tesc:19-20); hair.tese places each generated vertex: Bezier interpolation along the guide (func(),
// hair.tese:34-80) displaced sideways by width * dir (the multi-strand weights computed at :173-214 are dead:
// `pos = singleStrandPos`, :304).  This kernel produces the same vertices as a line-strip buffer, headless:
//     pos_width[e][k][j] = (func(u_k, v_j) + width_j * wr_k * sd_form(e,k)[j] * dir_k,  mix(0.02, 0.01, v_j))     :226-304, 313-315
//     tangent_u[e][k][j] = (normalize(P[seg+1] - P[seg]), u_k)                                                    :267, 275
// e = external strand index, u_k = k / isolines, v_j = j / divisions.  Everything that depends on (k) or (j) alone --
// the fract(sin()) hashes of u, the Gaussian / power profiles of v, segment index and Bezier parameter -- is tabulated
// on the host (ExpandTables); the only per-strand transcendental is the `randomChoice` hash of the root (:245), evaluated
// as float(sin(double)) so that CPU oracle and GPU pick the same deviation profile.  Deviation from the shader: at v = 1
// the shader indexes curve point N (out of bounds, :165-166); here the last vertex is the guide's last point.
constexpr int kMaxIsolines = 64, kMaxDivisions = 256, kExpandTile = 32;
struct ExpandTables {
    int seg[kMaxDivisions + 1];            // floor(v * (N-1)) clamped to N-2
    float t[kMaxDivisions + 1];            // Bezier parameter inside the segment
    float width[kMaxDivisions + 1];        // clumpRadius * mix(mix(.05,.3,v), mix(.3,.1,v), v)          :229
    float strand_width[kMaxDivisions + 1]; // mix(rootWidth, tipWidth, v)                                 :313-315
    float sd[6][kMaxDivisions + 1];        // deviation profiles: 0 none, 1..5 the branches of :247-259
    float u[kMaxIsolines], dirx[kMaxIsolines], dirz[kMaxIsolines], wr[kMaxIsolines];   // u_k, normalize(cos, 0, sin)(2 pi u), rand2 + 0.5
};

// random() of hair.tese:30-32 with the sine taken in double precision (see above)
__host__ __device__ inline float expand_hash(float x, float y) {
#ifdef __CUDA_ARCH__
    const float d = __fadd_rn(__fmul_rn(x, 12.9898f), __fmul_rn(y, 78.233f));
    const float h = __fmul_rn((float)sin((double)d), 43758.5453123f);
#else
    const float a = x * 12.9898f, b = y * 78.233f;
    const float d = a + b;
    const float h = (float)sin((double)d) * 43758.5453123f;
#endif
    return h - floorf(h);
}
__host__ __device__ inline int expand_form(float u, float rx, float ry, float rz) {     // hair.tese:245-259
    const float choice = expand_hash(u, rx) * expand_hash(ry, rz);
    return choice > 0.9f ? 1 : choice > 0.8f ? 2 : choice > 0.7f ? 3 : choice > 0.6f ? 4 : choice > 0.5f ? 5 : 0;
}

__device__ __forceinline__ float mixf(float a, float b, float t) { return fmaf(t, b - a, a); }

__global__ void __launch_bounds__(256)
k_expand_strands(const float* __restrict__ planes, const int* __restrict__ perm, const ExpandTables* __restrict__ tab,
                 float4* __restrict__ pos_width, float4* __restrict__ tangent_u, int S, int S_pad, int N, int I, int D) {
    // shared: guide positions of the tile, the per-(strand, isoline) deviation profile, and every table the vertex loop
    // reads (the loop is issue-bound: no integer division, no global table loads inside it)
    extern __shared__ float esm[];
    const int D1 = D + 1, per = I * D1;
    float* pos = esm;                                        // [3][N][kExpandTile + 1]
    float* tj = pos + 3 * N * (kExpandTile + 1);             // [4][D1]: t, width, strand_width, (int) seg
    float* sd = tj + 4 * D1;                                 // [6][D1]
    float* tk = sd + 6 * D1;                                 // [4][I]: u, dirx, dirz, wr
    int* kj = reinterpret_cast<int*>(tk + 4 * I);            // [per]: k << 16 | j of flat vertex q
    int* eidx = kj + per;                                    // [kExpandTile]: external strand index
    unsigned char* form = reinterpret_cast<unsigned char*>(eidx + kExpandTile);   // [kExpandTile][I]
    const int s0 = blockIdx.x * kExpandTile;
    if (threadIdx.x < kExpandTile) { const int s = s0 + threadIdx.x; eidx[threadIdx.x] = (perm && s < S) ? perm[s] : s; }
    for (int q = threadIdx.x; q < 3 * N * kExpandTile; q += blockDim.x) {
        const int sl = q % kExpandTile, r = q / kExpandTile;      // r = k*N + row
        const int k = r / N, row = r % N;
        pos[r * (kExpandTile + 1) + sl] = (s0 + sl < S_pad) ? planes[tiled_index(6, S_pad, row, k, s0 + sl)] : 0.f;
    }
    for (int j = threadIdx.x; j < D1; j += blockDim.x) {
        tj[j] = tab->t[j]; tj[D1 + j] = tab->width[j]; tj[2 * D1 + j] = tab->strand_width[j]; tj[3 * D1 + j] = __int_as_float(tab->seg[j]);
#pragma unroll
        for (int f = 0; f < 6; ++f) sd[f * D1 + j] = tab->sd[f][j];
    }
    for (int k = threadIdx.x; k < I; k += blockDim.x) { tk[k] = tab->u[k]; tk[I + k] = tab->dirx[k]; tk[2 * I + k] = tab->dirz[k]; tk[3 * I + k] = tab->wr[k]; }
    for (int q = threadIdx.x; q < per; q += blockDim.x) kj[q] = ((q / D1) << 16) | (q % D1);
    __syncthreads();
    for (int q = threadIdx.x; q < kExpandTile * I; q += blockDim.x) {
        const int sl = q / I, k = q % I;
        form[q] = (unsigned char)expand_form(tk[k], pos[(0 * N) * (kExpandTile + 1) + sl], pos[(1 * N) * (kExpandTile + 1) + sl], pos[(2 * N) * (kExpandTile + 1) + sl]);
    }
    __syncthreads();
    constexpr int RS = kExpandTile + 1;
    // flat walk over the tile's kExpandTile * per vertices, (strand, vertex) kept incrementally
    int sl = 0, q = threadIdx.x;
    while (q >= per) { q -= per; ++sl; }
    while (sl < kExpandTile && s0 + sl < S) {
        const int kq = kj[q], k = kq >> 16, j = kq & 0xffff;
        const int seg = __float_as_int(tj[3 * D1 + j]);
        const float t = tj[j];
        float c[3], tg[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float* P = pos + (a * N + seg) * RS + sl;
            const float v1 = P[0], v2 = P[RS];
            const float v0 = seg == 0 ? v1 + (v1 - v2) : P[-RS];
            const float v3 = seg + 1 == N - 1 ? v2 + (v2 - v1) : P[2 * RS];
            const float b1 = v1 + (1.0f / 3.0f) * ((v2 - v0) / 2.0f), b2 = v2 - (1.0f / 3.0f) * ((v3 - v1) / 2.0f);
            const float b01 = mixf(v1, b1, t), b11 = mixf(b1, b2, t), b21 = mixf(b2, v2, t);
            c[a] = mixf(mixf(b01, b11, t), mixf(b11, b21, t), t);
            tg[a] = v2 - v1;
        }
        const float w = tj[D1 + j] * tk[3 * I + k] * sd[form[sl * I + k] * D1 + j];
        const float inv = 1.0f / sqrtf(tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2]);
        const size_t o = (size_t)eidx[sl] * per + q;         // the strand's vertices are contiguous
        pos_width[o] = make_float4(fmaf(w, tk[I + k], c[0]), c[1], fmaf(w, tk[2 * I + k], c[2]), tj[2 * D1 + j]);
        tangent_u[o] = make_float4(tg[0] * inv, tg[1] * inv, tg[2] * inv, tk[k]);
        q += blockDim.x;
        while (q >= per) { q -= per; ++sl; }
    }
}

}  // namespace rvh
