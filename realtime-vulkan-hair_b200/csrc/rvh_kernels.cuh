// rvh_kernels.cuh -- sm_100a kernels of the guide-strand physics step.
//
// Device-side replacement for src/shaders/compute.comp (file:line below are relative to
// the reference tree).  Layout in HBM is point-major SoA ("planes"):
//     planes[k][i][s]   k in {px,py,pz,vx,vy,vz}, i = point on strand, s = strand
// so that a warp reading point i of 32*V consecutive strands issues fully coalesced
// 32/64/128-bit loads per plane, and the root->tip chain of a strand lives in registers.
// The voxel grid is int64 [G^3][4] (vx,vy,vz,density), fixed point x grid_scale; integer
// accumulation makes the result independent of atomics order and of the GPU count.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvh {

constexpr int kMaxEllipsoids = 7;       // colliders 1..7; collider 0 is the sphere
constexpr int kBlock = 128;

struct Ellipsoid {
    float inv[12];   // rows 0..2 of Collider::inv      : q = inv * (p,1)     compute.comp:64-67
    float xf[12];    // rows 0..2 of Collider::transform: on = xf * (nq,1)    compute.comp:76-79
    float nt[9];     // upper 3x3 of Collider::invTrans : n = nt * q          compute.comp:70-73
};

struct StepParams {
    int S, S_pad, N;
    float rest, gravity_y, damping, vmax, vmax2, penalty_k;
    float sphere_r, sphere_r2, sphere_c[3];
    int has_sphere, n_ell;
    Ellipsoid ell[kMaxEllipsoids];
    int G;
    float h, origin[3], scale, friction;
    float dt, inv_dt, dt2, vel_scale;      // vel_scale = damping / dt
    int wind_mode;                          // 0 off, 1 = variant A (:151), 2 = variant B (:152)
    float wind_s2T, wind_T3, wind_amp;      // 2*sin(2T); 3T; 10 (A) or 7*fbm(sinT,cosT) (B)
    int int32_wrap, keep_corr;
};

template <int V> struct VecOf;
template <> struct VecOf<1> { using type = float; };
template <> struct VecOf<2> { using type = float2; };
template <> struct VecOf<4> { using type = float4; };

template <int V> __device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&o)[V]) {
    if constexpr (V == 1) { o[0] = __ldg(p); }
    else if constexpr (V == 2) { float2 t = __ldg(reinterpret_cast<const float2*>(p)); o[0] = t.x; o[1] = t.y; }
    else { float4 t = __ldg(reinterpret_cast<const float4*>(p)); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
}
template <int V> __device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&o)[V]) {
    if constexpr (V == 1) { *p = o[0]; }
    else if constexpr (V == 2) { *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]); }
    else { *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]); }
}

// ---- per-axis cell range of a point: compute.comp:219-229 --------------------------------
// The shader's [max(floor,0), min(floor+1,G-1)] range is exactly "cells f and f+1, each kept
// only if it lies in [0,G-1]"; w0/w1 are clamp(1-|g-cell|,0,1) (compute.comp:237-239).
struct AxisCells { int f; float w0, w1; bool ok0, ok1; };

__device__ __forceinline__ AxisCells axis_cells(float p, float origin, float h, int G) {
    AxisCells a;
    const float g = __fdiv_rn(p - origin, h);
    float fl = floorf(g);
    fl = fminf(fmaxf(fl, -2.0f), (float)G);           // far-away / NaN points touch no cell
    a.f = (int)fl;
    a.w0 = __saturatef(1.0f - fabsf(g - (float)a.f));
    a.w1 = __saturatef(1.0f - fabsf(g - (float)(a.f + 1)));
    a.ok0 = (a.f >= 0) && (a.f <= G - 1);
    a.ok1 = (a.f + 1 >= 0) && (a.f + 1 <= G - 1);
    return a;
}

// ---- direct (slow-path) splat of one corner into the global int64 grid ------------------------
__device__ __forceinline__ void global_add4(unsigned long long* __restrict__ grid, int idx, int c0, int c1, int c2, int c3) {
    unsigned long long* cell = grid + 4 * (size_t)idx;
    if (c0) atomicAdd(cell + 0, (unsigned long long)(long long)c0);
    if (c1) atomicAdd(cell + 1, (unsigned long long)(long long)c1);
    if (c2) atomicAdd(cell + 2, (unsigned long long)(long long)c2);
    if (c3) atomicAdd(cell + 3, (unsigned long long)(long long)c3);
}

// ---- P2 splat of one point: compute.comp:231-252 ------------------------------------------
// Per corner: int(SCALE * (w * v_k)) and int(SCALE * w), truncated toward zero, exactly as the
// shader orders the float operations.  With a CTA-private box (BOX=true) the 32-bit partial sums
// go to shared memory (see k_grid_splat); otherwise straight to the global int64 grid.
struct SplatBox {
    int* bx; int* by; int* bz; int* bd;     // component-major int32 accumulators in shared memory
    int lo[3], n[3];                        // first cell and extent (cells) covered by the box
    int sy, sz;                             // odd strides (bank spreading)
    int* sat;                               // set when a cell hit its capacity
};
// A box cell may hold at most kDensLimit of density: every admitted velocity contribution obeys
// |c_k| <= kVBound * (c_d + 1), so |sum c_k| <= kVBound * (kDensLimit + 65536) < 2^31.
constexpr float kVBound = 15.9f;
constexpr int kDensLimit = (int)(2147483647.0 / 16.0) - 65536;

template <bool BOX>
__device__ __forceinline__ void splat_point(const StepParams& P, unsigned long long* __restrict__ grid, const SplatBox* B,
                                            float px, float py, float pz, float vx, float vy, float vz) {
    const AxisCells X = axis_cells(px, P.origin[0], P.h, P.G);
    const AxisCells Y = axis_cells(py, P.origin[1], P.h, P.G);
    const AxisCells Z = axis_cells(pz, P.origin[2], P.h, P.G);
    bool fast = false;
    int bbase = 0;
    if (BOX) {
        const int rx = X.f - B->lo[0], ry = Y.f - B->lo[1], rz = Z.f - B->lo[2];
        const float vinf = fmaxf(fabsf(vx), fmaxf(fabsf(vy), fabsf(vz)));
        fast = (rx >= 0) & (ry >= 0) & (rz >= 0) & (rx <= B->n[0] - 2) & (ry <= B->n[1] - 2) & (rz <= B->n[2] - 2) & (vinf <= kVBound);
        bbase = rx + ry * B->sy + rz * B->sz;
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        if (!(a ? X.ok1 : X.ok0)) continue;
        const float xw = a ? X.w1 : X.w0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            if (!(b ? Y.ok1 : Y.ok0)) continue;
            const float xyw = __fmul_rn(xw, b ? Y.w1 : Y.w0);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (!(c ? Z.ok1 : Z.ok0)) continue;
                const float tw = __fmul_rn(xyw, c ? Z.w1 : Z.w0);
                // int(SCALE * weightedVelocity.k), int(SCALE * totalWeight): truncation toward zero
                const int c0 = __float2int_rz(__fmul_rn(P.scale, __fmul_rn(tw, vx)));
                const int c1 = __float2int_rz(__fmul_rn(P.scale, __fmul_rn(tw, vy)));
                const int c2 = __float2int_rz(__fmul_rn(P.scale, __fmul_rn(tw, vz)));
                const int c3 = __float2int_rz(__fmul_rn(P.scale, tw));
                const int idx = (X.f + a) + (Y.f + b) * P.G + (Z.f + c) * P.G * P.G;
                if (BOX && fast) {
                    const int cell = bbase + a + b * B->sy + c * B->sz;
                    const int old = atomicAdd(B->bd + cell, c3);
                    if (old + c3 <= kDensLimit) {
                        if (c0) atomicAdd(B->bx + cell, c0);
                        if (c1) atomicAdd(B->by + cell, c1);
                        if (c2) atomicAdd(B->bz + cell, c2);
                    } else {                       // cell full: take the density back, go to the global grid
                        atomicAdd(B->bd + cell, -c3);
                        *B->sat = 1;
                        global_add4(grid, idx, c0, c1, c2, c3);
                    }
                } else {
                    global_add4(grid, idx, c0, c1, c2, c3);
                }
            }
        }
    }
}

// ---- P3 gather of one point: compute.comp:259-297 ------------------------------------------
// fgrid[cell] = (float(vel.x), float(vel.y), float(vel.z), density > 0 ? 1/float(density) : 0),
// prepared once per step by k_grid_finalize.  Written without FMA contraction so the result is
// bit-identical to the C oracle when positions, velocities and grid are.
__device__ __forceinline__ void gather_point(const StepParams& P, const float4* __restrict__ fgrid,
                                             float px, float py, float pz, float& vx, float& vy, float& vz) {
    const AxisCells X = axis_cells(px, P.origin[0], P.h, P.G);
    const AxisCells Y = axis_cells(py, P.origin[1], P.h, P.G);
    const AxisCells Z = axis_cells(pz, P.origin[2], P.h, P.G);
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        if (!(a ? X.ok1 : X.ok0)) continue;
        const float xw = a ? X.w1 : X.w0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            if (!(b ? Y.ok1 : Y.ok0)) continue;
            const float xyw = __fmul_rn(xw, b ? Y.w1 : Y.w0);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (!(c ? Z.ok1 : Z.ok0)) continue;
                const float tw = __fmul_rn(xyw, c ? Z.w1 : Z.w0);
                const int idx = (X.f + a) + (Y.f + b) * P.G + (Z.f + c) * P.G * P.G;
                const float4 cell = __ldg(fgrid + idx);
                if (cell.w > 0.f) {                                     // density > 0, compute.comp:276
                    const float s = __fmul_rn(tw, cell.w);              // totalWeight * (1.0 / float(density))
                    gx = __fadd_rn(gx, __fmul_rn(s, cell.x));
                    gy = __fadd_rn(gy, __fmul_rn(s, cell.y));
                    gz = __fadd_rn(gz, __fmul_rn(s, cell.z));
                }
            }
        }
    }
    const float fr = P.friction, omf = __fsub_rn(1.0f, fr);
    vx = __fadd_rn(__fmul_rn(omf, vx), __fmul_rn(fr, gx));
    vy = __fadd_rn(__fmul_rn(omf, vy), __fmul_rn(fr, gy));
    vz = __fadd_rn(__fmul_rn(omf, vz), __fmul_rn(fr, gz));
}

// ---- wind trigonometry ---------------------------------------------------------------------------
// sin/cos with a two-constant Cody-Waite reduction by pi and degree-9/10 polynomials on
// [-pi/2, pi/2]: absolute error < 2e-7 for |x| < 1e4, which is far inside what the wind force
// needs (it enters positions multiplied by dt^2).  No slow path, no local memory, ~14 instructions.
__device__ __forceinline__ float reduce_pi(float x, unsigned& sign) {
    const float kf = fmaf(x, 0.31830987f, 12582912.0f);     // round(x/pi) in the low mantissa bits
    sign = __float_as_uint(kf) << 31;                          // parity of k
    const float k = kf - 12582912.0f;
    float r = fmaf(k, -3.141592741f, x);
    r = fmaf(k, 8.742277657e-08f, r);
    return r;
}
__device__ __forceinline__ float sin_bounded(float x) {
    unsigned sign;
    const float r = reduce_pi(x, sign), r2 = r * r;
    float p = fmaf(r2, 2.5992781e-06f, -0.00019806201f);
    p = fmaf(r2, p, 0.0083330106f);
    p = fmaf(r2, p, -0.16666657f);
    const float s = fmaf(r * r2, p, r);
    return fminf(fmaxf(__uint_as_float(__float_as_uint(s) ^ sign), -1.0f), 1.0f);
}
__device__ __forceinline__ float cos_bounded(float x) {
    unsigned sign;
    const float r = reduce_pi(x, sign), r2 = r * r;
    float p = fmaf(r2, -2.6073479e-07f, 2.4761655e-05f);
    p = fmaf(r2, p, -0.0013888398f);
    p = fmaf(r2, p, 0.041666642f);
    p = fmaf(r2, p, -0.5f);
    const float c = fmaf(r2, p, 1.0f);
    return fminf(fmaxf(__uint_as_float(__float_as_uint(c) ^ sign), -1.0f), 1.0f);
}
__device__ __forceinline__ float rsqrt_fast(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- P1 for one point: compute.comp:144-201 -------------------------------------------------
struct PointOut { float px, py, pz, vx, vy, vz, dx, dy, dz; };

// q = inv * (p,1) of ellipsoid E and |q|^2 (compute.comp:64-67)
__device__ __forceinline__ float ellipsoid_q(const Ellipsoid& E, float cx, float cy, float cz, float& qx, float& qy, float& qz) {
    qx = fmaf(E.inv[0], cx, fmaf(E.inv[1], cy, fmaf(E.inv[2], cz, E.inv[3])));
    qy = fmaf(E.inv[4], cx, fmaf(E.inv[5], cy, fmaf(E.inv[6], cz, E.inv[7])));
    qz = fmaf(E.inv[8], cx, fmaf(E.inv[9], cy, fmaf(E.inv[10], cz, E.inv[11])));
    return fmaf(qx, qx, fmaf(qy, qy, qz * qz));
}

// NELL >= 0: number of ellipsoids known at compile time (constants become immediate constant-bank
// operands of the FFMAs); NELL < 0: run-time count.
template <bool WIND, int NELL>
__device__ __forceinline__ PointOut point_update(const StepParams& P, float cx, float cy, float cz,
                                                 float vx, float vy, float vz,
                                                 float parx, float pary, float parz) {
    float fx = 0.0f, fy = P.gravity_y, fz = 0.0f;                       // :150
    if (WIND) {
        const float wx = P.wind_s2T * cos_bounded(cy * 10.0f) * sin_bounded((cy + 5.0f) * 15.0f);
        fx = fmaf(P.wind_amp, wx, fx);
        if (P.wind_mode == 1) {                                         // :151
            fz = fmaf(P.wind_amp, -fminf(fmaxf(cy * 2.0f, 0.2f), 2.0f), fz);
        } else {                                                        // :152
            fy = fmaf(P.wind_amp, 4.0f * sin_bounded(fmaf(cz, 5.0f, P.wind_T3)), fy);
            fz = fmaf(P.wind_amp, -0.6f * (cy + 3.0f), fz);
        }
    }

    // collision tests first (one predicate per collider), bodies only for the colliders hit
    const int nell = NELL >= 0 ? NELL : P.n_ell;
    unsigned hit = 0;
    {
        const float dx = cx - P.sphere_c[0], dy = cy - P.sphere_c[1], dz = cz - P.sphere_c[2];
        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        if (d2 < P.sphere_r2) hit = P.has_sphere;                       // :162
    }
    if (NELL >= 0) {
#pragma unroll
        for (int j = 0; j < (NELL >= 0 ? NELL : 0); ++j) {
            float qx, qy, qz;
            if (ellipsoid_q(P.ell[j], cx, cy, cz, qx, qy, qz) <= 1.0f) hit |= 2u << j;   // :66
        }
    } else {
        for (int j = 0; j < nell; ++j) {
            float qx, qy, qz;
            if (ellipsoid_q(P.ell[j], cx, cy, cz, qx, qy, qz) <= 1.0f) hit |= 2u << j;
        }
    }
    if (hit) {
        float ax = 0.f, ay = 0.f, az = 0.f;
        if (hit & 1u) {                                                 // :160-169
            const float dx = cx - P.sphere_c[0], dy = cy - P.sphere_c[1], dz = cz - P.sphere_c[2];
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            const float rinv = rsqrt_fast(d2);
            const float s = P.penalty_k * (P.sphere_r - d2 * rinv) * rinv;
            ax = s * dx; ay = s * dy; az = s * dz;
        }
        unsigned m = hit >> 1;
        while (m) {                                                     // :170-179
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const Ellipsoid& E = P.ell[j];
            float qx, qy, qz;
            const float q2 = ellipsoid_q(E, cx, cy, cz, qx, qy, qz);
            const float rq = rsqrt_fast(q2);
            const float ux = qx * rq, uy = qy * rq, uz = qz * rq;
            const float ox = fmaf(E.xf[0], ux, fmaf(E.xf[1], uy, fmaf(E.xf[2], uz, E.xf[3])));
            const float oy = fmaf(E.xf[4], ux, fmaf(E.xf[5], uy, fmaf(E.xf[6], uz, E.xf[7])));
            const float oz = fmaf(E.xf[8], ux, fmaf(E.xf[9], uy, fmaf(E.xf[10], uz, E.xf[11])));
            const float ex = cx - ox, ey = cy - oy, ez = cz - oz;
            const float d = sqrtf(fmaf(ex, ex, fmaf(ey, ey, ez * ez)));
            const float nx = fmaf(E.nt[0], qx, fmaf(E.nt[1], qy, E.nt[2] * qz));
            const float ny = fmaf(E.nt[3], qx, fmaf(E.nt[4], qy, E.nt[5] * qz));
            const float nz = fmaf(E.nt[6], qx, fmaf(E.nt[7], qy, E.nt[8] * qz));
            const float s = P.penalty_k * d * rsqrt_fast(fmaf(nx, nx, fmaf(ny, ny, nz * nz)));
            ax = fmaf(s, nx, ax); ay = fmaf(s, ny, ay); az = fmaf(s, nz, az);
        }
        const float ih = __frcp_rn((float)__popc(hit));                 // :182-184
        fx = fmaf(ax, ih, fx); fy = fmaf(ay, ih, fy); fz = fmaf(az, ih, fz);
    }

    PointOut o;
    const float prx = fmaf(P.dt2, fx, fmaf(P.dt, vx, cx));              // :187
    const float pry = fmaf(P.dt2, fy, fmaf(P.dt, vy, cy));
    const float prz = fmaf(P.dt2, fz, fmaf(P.dt, vz, cz));
    const float ddx = prx - parx, ddy = pry - pary, ddz = prz - parz;   // :191-192
    const float sc = P.rest * rsqrt_fast(fmaf(ddx, ddx, fmaf(ddy, ddy, ddz * ddz)));
    o.px = fmaf(sc, ddx, parx); o.py = fmaf(sc, ddy, pary); o.pz = fmaf(sc, ddz, parz);
    float nvx = (o.px - cx) * P.vel_scale, nvy = (o.py - cy) * P.vel_scale, nvz = (o.pz - cz) * P.vel_scale;  // :195-197
    const float l2 = fmaf(nvx, nvx, fmaf(nvy, nvy, nvz * nvz));
    if (l2 > P.vmax2) {                                                 // :198-200
        const float s = P.vmax * rsqrt_fast(l2);
        nvx *= s; nvy *= s; nvz *= s;
    }
    o.vx = nvx; o.vy = nvy; o.vz = nvz;
    o.dx = P.damping * (o.px - prx); o.dy = P.damping * (o.py - pry); o.dz = P.damping * (o.pz - prz);  // :201
    return o;
}

// ---- K1: integrate + collide + FTL + corrected velocity -------------------------------------
// One thread owns V consecutive strands and walks them root->tip together (V independent
// dependency chains per thread).  Point i's velocity is final only once d_{i+1} is known
// (compute.comp:213-215), so the velocity store trails the position by one point.
// With BBOX the block also reduces the bounding box of its strands' new positions, which sizes
// the shared-memory box of the splat kernel that handles the same strands.
template <int V, bool WIND, int NELL, bool BBOX>
__global__ void __launch_bounds__(kBlock)
k_ftl_step(const __grid_constant__ StepParams P, float* __restrict__ planes, float* __restrict__ corr,
           float* __restrict__ bbox) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s0 = t * V;
    const bool active = s0 < P.S_pad;
    float bmin[3] = { 3.0e38f, 3.0e38f, 3.0e38f }, bmax[3] = { -3.0e38f, -3.0e38f, -3.0e38f };
    if (active) {
        const size_t plane = (size_t)P.N * P.S_pad;
        float* const ppx = planes + s0;
        float* const ppy = ppx + plane;
        float* const ppz = ppy + plane;
        float* const pvx = ppz + plane;
        float* const pvy = pvx + plane;
        float* const pvz = pvy + plane;

        float parx[V], pary[V], parz[V];
        load_vec<V>(ppx, parx); load_vec<V>(ppy, pary); load_vec<V>(ppz, parz);
        float nx[V], ny[V], nz[V], nvx[V], nvy[V], nvz[V];
        {
            const size_t o = P.S_pad;
            load_vec<V>(ppx + o, nx); load_vec<V>(ppy + o, ny); load_vec<V>(ppz + o, nz);
            load_vec<V>(pvx + o, nvx); load_vec<V>(pvy + o, nvy); load_vec<V>(pvz + o, nvz);
        }
        float lvx[V], lvy[V], lvz[V];   // clamped velocity of the previous point, correction pending
#pragma unroll
        for (int u = 0; u < V; ++u) { lvx[u] = 0.f; lvy[u] = 0.f; lvz[u] = 0.f; }

        size_t oi = P.S_pad;
        for (int i = 1; i < P.N; ++i, oi += P.S_pad) {
            float cx[V], cy[V], cz[V], vx[V], vy[V], vz[V];
#pragma unroll
            for (int u = 0; u < V; ++u) { cx[u] = nx[u]; cy[u] = ny[u]; cz[u] = nz[u]; vx[u] = nvx[u]; vy[u] = nvy[u]; vz[u] = nvz[u]; }
            if (i + 1 < P.N) {
                const size_t o = oi + P.S_pad;
                load_vec<V>(ppx + o, nx); load_vec<V>(ppy + o, ny); load_vec<V>(ppz + o, nz);
                load_vec<V>(pvx + o, nvx); load_vec<V>(pvy + o, nvy); load_vec<V>(pvz + o, nvz);
            }
            float opx[V], opy[V], opz[V], fvx[V], fvy[V], fvz[V], odx[V], ody[V], odz[V];
#pragma unroll
            for (int u = 0; u < V; ++u) {
                const PointOut o = point_update<WIND, NELL>(P, cx[u], cy[u], cz[u], vx[u], vy[u], vz[u], parx[u], pary[u], parz[u]);
                opx[u] = o.px; opy[u] = o.py; opz[u] = o.pz;
                odx[u] = o.dx; ody[u] = o.dy; odz[u] = o.dz;
                // finalise point i-1: v_{i-1} -= d_i / dt   (compute.comp:213-215)
                fvx[u] = fmaf(-o.dx, P.inv_dt, lvx[u]); fvy[u] = fmaf(-o.dy, P.inv_dt, lvy[u]); fvz[u] = fmaf(-o.dz, P.inv_dt, lvz[u]);
                lvx[u] = o.vx; lvy[u] = o.vy; lvz[u] = o.vz;
                if (BBOX && s0 + u < P.S) {
                    bmin[0] = fminf(bmin[0], o.px); bmin[1] = fminf(bmin[1], o.py); bmin[2] = fminf(bmin[2], o.pz);
                    bmax[0] = fmaxf(bmax[0], o.px); bmax[1] = fmaxf(bmax[1], o.py); bmax[2] = fmaxf(bmax[2], o.pz);
                }
            }
            store_vec<V>(ppx + oi, opx); store_vec<V>(ppy + oi, opy); store_vec<V>(ppz + oi, opz);
            if (P.keep_corr) {
                float* c0 = corr + s0 + oi;
                store_vec<V>(c0, odx); store_vec<V>(c0 + plane, ody); store_vec<V>(c0 + 2 * plane, odz);
            }
            if (i > 1) {
                const size_t om = oi - P.S_pad;
                store_vec<V>(pvx + om, fvx); store_vec<V>(pvy + om, fvy); store_vec<V>(pvz + om, fvz);
            }
#pragma unroll
            for (int u = 0; u < V; ++u) { parx[u] = opx[u]; pary[u] = opy[u]; parz[u] = opz[u]; }
        }
        // last point: no correction term (compute.comp:213 `i != NUM_CURVE_POINTS - 1`)
        const size_t ol = (size_t)(P.N - 1) * P.S_pad;
        store_vec<V>(pvx + ol, lvx); store_vec<V>(pvy + ol, lvy); store_vec<V>(pvz + ol, lvz);
    }
    if (BBOX) {
        __shared__ float red[6][kBlock / 32];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                bmin[k] = fminf(bmin[k], __shfl_xor_sync(0xffffffffu, bmin[k], d));
                bmax[k] = fmaxf(bmax[k], __shfl_xor_sync(0xffffffffu, bmax[k], d));
            }
        }
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { red[k][w] = bmin[k]; red[3 + k][w] = bmax[k]; }
        }
        __syncthreads();
        if (threadIdx.x < 6) {
            float v = red[threadIdx.x][0];
            for (int q = 1; q < kBlock / 32; ++q) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][q]) : fmaxf(v, red[threadIdx.x][q]);
            bbox[6 * blockIdx.x + threadIdx.x] = v;
        }
    }
}

// ---- K_splat: corrected velocities -> voxel grid ------------------------------------------------
// Block b owns the same kBlock*V strands as block b of k_ftl_step.  Points of one ROW of neighbouring
// strands fall into the same one or two cells (hair is dense), so a lane-per-strand splat would
// serialise every shared-memory atomic ~13-fold.  Instead a tile of 32 strands x all rows is staged
// through shared memory (coalesced 128-byte row segments in, transposed out) and each warp walks ONE
// strand with lane = row: the 32 lanes then hit ~23 different cells and the atomics run conflict-light.
// Partial sums live in a CTA-private int32 box sized from the block's bounding box (written by
// k_ftl_step) and are flushed once with 64-bit global reductions.
constexpr int kSplatThreads = 256;
constexpr int kBoxCells = 3072;             // 48 KB of int32 x 4 components
constexpr int kTileRowsMax = 63;            // N <= 64

__device__ __forceinline__ void box_flush(const StepParams& P, const SplatBox& B, unsigned long long* __restrict__ grid, int ncell_padded) {
    for (int k = threadIdx.x; k < ncell_padded; k += blockDim.x) {
        const int d = B.bd[k], x = B.bx[k], y = B.by[k], z = B.bz[k];
        if ((d | x | y | z) != 0) {
            const int cz = k / B.sz, rem = k - cz * B.sz;
            const int cy = rem / B.sy, cx = rem - cy * B.sy;
            const int idx = (B.lo[0] + cx) + (B.lo[1] + cy) * P.G + (B.lo[2] + cz) * P.G * P.G;
            global_add4(grid, idx, x, y, z, d);
            B.bd[k] = 0; B.bx[k] = 0; B.by[k] = 0; B.bz[k] = 0;
        }
    }
}

__global__ void __launch_bounds__(kSplatThreads)
k_grid_splat(const __grid_constant__ StepParams P, const float* __restrict__ planes, const float* __restrict__ bbox,
             unsigned long long* __restrict__ grid, int strands_per_block) {
    extern __shared__ int smem[];
    __shared__ int s_sat;
    int* box = smem;                                        // [4][kBoxCells]
    float* tile = reinterpret_cast<float*>(smem + 4 * kBoxCells);   // [6][rows][33]
    const int rows = P.N - 1;
    const int s_begin = blockIdx.x * strands_per_block;
    if (s_begin >= P.S) return;
    const int s_end = min(s_begin + strands_per_block, P.S);

    // box geometry from the block's bounding box
    SplatBox B;
    B.bx = box; B.by = box + kBoxCells; B.bz = box + 2 * kBoxCells; B.bd = box + 3 * kBoxCells;
    B.sat = &s_sat;
    {
        const float* bb = bbox + 6 * blockIdx.x;
        int hi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const AxisCells lo = axis_cells(bb[k], P.origin[k], P.h, P.G), up = axis_cells(bb[3 + k], P.origin[k], P.h, P.G);
            B.lo[k] = max(lo.f, 0);
            hi[k] = min(up.f + 1, P.G - 1);
            B.n[k] = max(hi[k] - B.lo[k] + 1, 2);
        }
        // shrink until it fits (points outside the box take the global path)
        int n0 = B.n[0], n1 = B.n[1], n2 = B.n[2];
        while (true) {
            const int sy = n0 | 1, sz = (sy * n1) | 1;
            if ((long long)sz * n2 <= kBoxCells) { B.sy = sy; B.sz = sz; break; }
            if (n0 >= n1 && n0 >= n2) n0 = (n0 + 1) / 2;
            else if (n1 >= n2) n1 = (n1 + 1) / 2;
            else n2 = (n2 + 1) / 2;
        }
        B.n[0] = n0; B.n[1] = n1; B.n[2] = n2;
    }
    const int ncell = B.sz * B.n[2];
    for (int k = threadIdx.x; k < 4 * kBoxCells; k += blockDim.x) box[k] = 0;
    if (threadIdx.x == 0) s_sat = 0;
    __syncthreads();

    const size_t plane = (size_t)P.N * P.S_pad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kSplatThreads / 32;
    const int tstride = rows * 33;
    for (int t0 = s_begin; t0 < s_end; t0 += 32) {
        // stage rows 1..N-1 of strands t0..t0+31: lane = strand (coalesced), warp strides rows
        for (int r = warp; r < rows; r += nwarps) {
            const size_t g = (size_t)(r + 1) * P.S_pad + t0 + lane;
#pragma unroll
            for (int k = 0; k < 6; ++k) tile[k * tstride + r * 33 + lane] = __ldg(planes + k * plane + g);
        }
        __syncthreads();
        // splat: lane = row (short strands: several strands share a warp pass), warps stride strands
        const int nst = min(32, s_end - t0);
        int lps = 32;                                   // lanes per strand
        while ((lps >> 1) >= rows && lps > 1) lps >>= 1;
        const int spp = 32 / lps;                       // strands per warp pass
        const int sub = lane / lps, rl = lane % lps;
        for (int rb = 0; rb < rows; rb += lps) {
            const int r = rb + rl;
            for (int sl = warp * spp + sub; sl < nst; sl += nwarps * spp) {
                if (r < rows) {
                    const int o = r * 33 + sl;
                    splat_point<true>(P, grid, &B, tile[o], tile[tstride + o], tile[2 * tstride + o],
                                      tile[3 * tstride + o], tile[4 * tstride + o], tile[5 * tstride + o]);
                }
            }
        }
        if (__syncthreads_or(s_sat)) {           // some cell reached its int32-safe capacity: empty the box
            box_flush(P, B, grid, ncell);
            if (threadIdx.x == 0) s_sat = 0;
            __syncthreads();
        }
    }
    box_flush(P, B, grid, ncell);
}

// ---- grid finalize: int64 accumulators -> float cells for the gather ---------------------------
__global__ void __launch_bounds__(256)
k_grid_finalize(const long long* __restrict__ grid, float4* __restrict__ fgrid, int cells, int int32_wrap) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cells) return;
    const longlong2* c = reinterpret_cast<const longlong2*>(grid + 4 * (size_t)k);
    const longlong2 v01 = c[0], v2d = c[1];
    long long dens = v2d.y, v0 = v01.x, v1 = v01.y, v2 = v2d.x;
    if (int32_wrap) { dens = (int)dens; v0 = (int)v0; v1 = (int)v1; v2 = (int)v2; }   // the reference's int32 GridCell
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dens > 0) {
        o.x = __ll2float_rn(v0); o.y = __ll2float_rn(v1); o.z = __ll2float_rn(v2);
        o.w = __frcp_rn(__ll2float_rn(dens));                    // 1.0 / float(density), compute.comp:283
    }
    fgrid[k] = o;
}

// ---- K2: grid gather + friction, one thread per point -------------------------------------
__global__ void __launch_bounds__(256)
k_grid_gather(const __grid_constant__ StepParams P, float* __restrict__ planes, const float4* __restrict__ fgrid) {
    const size_t plane = (size_t)P.N * P.S_pad;
    const size_t total = (size_t)(P.N - 1) * P.S_pad;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(k % P.S_pad);
        if (s >= P.S) continue;
        const size_t o = k + P.S_pad;   // skip the root row
        float vx = planes[3 * plane + o], vy = planes[4 * plane + o], vz = planes[5 * plane + o];
        gather_point(P, fgrid, planes[o], planes[plane + o], planes[2 * plane + o], vx, vy, vz);
        planes[3 * plane + o] = vx; planes[4 * plane + o] = vy; planes[5 * plane + o] = vz;
    }
}

// ---- AoS <-> planes (the reference's Strand[S] vertex-buffer layout, Strand.h:11-15) -------
// One block per tile of 32 strands; shared-memory transpose so both sides are coalesced.
// perm[s_internal] = external strand index (nullptr = identity).
constexpr int kTile = 32;

__global__ void __launch_bounds__(256)
k_unpack_aos(const float4* __restrict__ aos, float* __restrict__ planes, const int* __restrict__ perm,
             int S, int S_pad, int N, float rest) {
    extern __shared__ float sm[];            // [6][N][kTile+1]
    const int tile0 = blockIdx.x * kTile;
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
    const size_t plane = (size_t)N * S_pad;
    for (int k = threadIdx.x; k < 6 * N * kTile; k += blockDim.x) {
        const int sl = k % kTile, r = k / kTile;     // r = kk*N + j
        const int kk = r / N, j = r % N;
        if (tile0 + sl < S_pad) planes[kk * plane + (size_t)j * S_pad + tile0 + sl] = sm[r * (kTile + 1) + sl];
    }
}

__global__ void __launch_bounds__(256)
k_pack_aos(float4* __restrict__ aos, const float* __restrict__ planes, const float* __restrict__ corr,
           const int* __restrict__ perm, int S, int S_pad, int N) {
    extern __shared__ float sm[];            // [9][N][kTile+1]
    const int tile0 = blockIdx.x * kTile;
    const size_t plane = (size_t)N * S_pad;
    const int nk = corr ? 9 : 6;
    for (int k = threadIdx.x; k < nk * N * kTile; k += blockDim.x) {
        const int sl = k % kTile, r = k / kTile;
        const int kk = r / N, j = r % N;
        float v = 0.f;
        if (tile0 + sl < S_pad)
            v = kk < 6 ? planes[kk * plane + (size_t)j * S_pad + tile0 + sl]
                       : corr[(kk - 6) * plane + (size_t)j * S_pad + tile0 + sl];
        sm[r * (kTile + 1) + sl] = v;
    }
    __syncthreads();
    const int q3 = 3 * N;
    for (int k = threadIdx.x; k < kTile * q3; k += blockDim.x) {
        const int sl = k / q3, q = k % q3;
        const int s = tile0 + sl;
        if (s >= S) continue;
        const int a = q / N, j = q % N;      // a: 0 curvePoints, 1 curveVels, 2 correctionVecs
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

}  // namespace rvh
