/*
 * oracle.c -- CPU restatement of the reference compute pass.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h for the scope statement, the reference file:line map and the pin).
 *
 * Build with -ffp-contract=off: every float operation below is written in the order
 * glm 0.9.9.0 / the GLSL text evaluates it, and must not be fused.
 */
#include "oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } v3;

/* ---- glm 0.9.9.0 restated (column-major: m[col*4+row]) ------------------------------ */

static void mat_identity(float* m) { memset(m, 0, 64); m[0] = m[5] = m[10] = m[15] = 1.f; }

/* glm/detail/type_mat4x4.inl:581-599: Result[c] = ((A0*b0 + A1*b1) + A2*b2) + A3*b3 */
static void mat_mul(const float* a, const float* b, float* out) {
    float r[16];
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < 4; ++i) {
            float t = a[0 * 4 + i] * b[c * 4 + 0] + a[1 * 4 + i] * b[c * 4 + 1];
            t = t + a[2 * 4 + i] * b[c * 4 + 2];
            t = t + a[3 * 4 + i] * b[c * 4 + 3];
            r[c * 4 + i] = t;
        }
    memcpy(out, r, 64);
}

/* glm/gtc/matrix_transform.inl:11-16 */
static void mat_translate(const float* m, const float v[3], float* out) {
    float r[16];
    memcpy(r, m, 64);
    for (int i = 0; i < 4; ++i) {
        float t = m[0 * 4 + i] * v[0] + m[1 * 4 + i] * v[1];
        t = t + m[2 * 4 + i] * v[2];
        t = t + m[3 * 4 + i];
        r[3 * 4 + i] = t;
    }
    memcpy(out, r, 64);
}

/* glm/gtc/matrix_transform.inl:19-47 with m = identity (gtx/transform.inl:13-16) */
static void mat_rotate(float angle, const float v[3], float* out) {
    float m[16];
    mat_identity(m);
    const float c = cosf(angle), s = sinf(angle);
    /* normalize(v) = v * (1/sqrt(dot(v,v))) */
    float d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    float inv = 1.0f / sqrtf(d);
    float axis[3] = { v[0] * inv, v[1] * inv, v[2] * inv };
    float omc = 1.0f - c;
    float temp[3] = { omc * axis[0], omc * axis[1], omc * axis[2] };
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = temp[0] * axis[1] + s * axis[2];
    R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1];
    R[2][1] = temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    for (int col = 0; col < 3; ++col)
        for (int i = 0; i < 4; ++i) {
            float t = m[0 * 4 + i] * R[col][0] + m[1 * 4 + i] * R[col][1];
            t = t + m[2 * 4 + i] * R[col][2];
            out[col * 4 + i] = t;
        }
    for (int i = 0; i < 4; ++i) out[3 * 4 + i] = m[3 * 4 + i];
}

/* glm/gtc/matrix_transform.inl:79-87 with m = identity */
static void mat_scale(const float v[3], float* out) {
    float m[16];
    mat_identity(m);
    for (int i = 0; i < 4; ++i) {
        out[0 * 4 + i] = m[0 * 4 + i] * v[0];
        out[1 * 4 + i] = m[1 * 4 + i] * v[1];
        out[2 * 4 + i] = m[2 * 4 + i] * v[2];
        out[3 * 4 + i] = m[3 * 4 + i];
    }
}

/* glm/detail/func_matrix.inl:297-355 */
static void mat_inverse(const float* mm, float* out) {
#define M(c, r) mm[(c) * 4 + (r)]
    float Coef00 = M(2,2) * M(3,3) - M(3,2) * M(2,3);
    float Coef02 = M(1,2) * M(3,3) - M(3,2) * M(1,3);
    float Coef03 = M(1,2) * M(2,3) - M(2,2) * M(1,3);
    float Coef04 = M(2,1) * M(3,3) - M(3,1) * M(2,3);
    float Coef06 = M(1,1) * M(3,3) - M(3,1) * M(1,3);
    float Coef07 = M(1,1) * M(2,3) - M(2,1) * M(1,3);
    float Coef08 = M(2,1) * M(3,2) - M(3,1) * M(2,2);
    float Coef10 = M(1,1) * M(3,2) - M(3,1) * M(1,2);
    float Coef11 = M(1,1) * M(2,2) - M(2,1) * M(1,2);
    float Coef12 = M(2,0) * M(3,3) - M(3,0) * M(2,3);
    float Coef14 = M(1,0) * M(3,3) - M(3,0) * M(1,3);
    float Coef15 = M(1,0) * M(2,3) - M(2,0) * M(1,3);
    float Coef16 = M(2,0) * M(3,2) - M(3,0) * M(2,2);
    float Coef18 = M(1,0) * M(3,2) - M(3,0) * M(1,2);
    float Coef19 = M(1,0) * M(2,2) - M(2,0) * M(1,2);
    float Coef20 = M(2,0) * M(3,1) - M(3,0) * M(2,1);
    float Coef22 = M(1,0) * M(3,1) - M(3,0) * M(1,1);
    float Coef23 = M(1,0) * M(2,1) - M(2,0) * M(1,1);
    float Fac0[4] = { Coef00, Coef00, Coef02, Coef03 };
    float Fac1[4] = { Coef04, Coef04, Coef06, Coef07 };
    float Fac2[4] = { Coef08, Coef08, Coef10, Coef11 };
    float Fac3[4] = { Coef12, Coef12, Coef14, Coef15 };
    float Fac4[4] = { Coef16, Coef16, Coef18, Coef19 };
    float Fac5[4] = { Coef20, Coef20, Coef22, Coef23 };
    float Vec0[4] = { M(1,0), M(0,0), M(0,0), M(0,0) };
    float Vec1[4] = { M(1,1), M(0,1), M(0,1), M(0,1) };
    float Vec2[4] = { M(1,2), M(0,2), M(0,2), M(0,2) };
    float Vec3[4] = { M(1,3), M(0,3), M(0,3), M(0,3) };
    static const float SignA[4] = { +1.f, -1.f, +1.f, -1.f };
    static const float SignB[4] = { -1.f, +1.f, -1.f, +1.f };
    float Inv[16];
    for (int i = 0; i < 4; ++i) {
        float i0 = (Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i]) + Vec3[i] * Fac2[i];
        float i1 = (Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i]) + Vec3[i] * Fac4[i];
        float i2 = (Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i]) + Vec3[i] * Fac5[i];
        float i3 = (Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i]) + Vec2[i] * Fac5[i];
        Inv[0 * 4 + i] = i0 * SignA[i];
        Inv[1 * 4 + i] = i1 * SignB[i];
        Inv[2 * 4 + i] = i2 * SignA[i];
        Inv[3 * 4 + i] = i3 * SignB[i];
    }
    float d0 = M(0,0) * Inv[0 * 4 + 0], d1 = M(0,1) * Inv[1 * 4 + 0];
    float d2 = M(0,2) * Inv[2 * 4 + 0], d3 = M(0,3) * Inv[3 * 4 + 0];
    float Dot1 = (d0 + d1) + (d2 + d3);
    float OneOverDeterminant = 1.0f / Dot1;
    for (int i = 0; i < 16; ++i) out[i] = Inv[i] * OneOverDeterminant;
#undef M
}

static void mat_transpose(const float* m, float* out) {
    float r[16];
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < 4; ++i) r[c * 4 + i] = m[i * 4 + c];
    memcpy(out, r, 64);
}

/* glm/detail/type_mat4x4.inl:512-523: (m0*v0 + m1*v1) + (m2*v2 + m3*v3); xyz only. */
static inline v3 mat_mul_vec_xyz(const float* m, v3 v, float w) {
    v3 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8]  * v.z + m[12] * w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9]  * v.z + m[13] * w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * w);
    return r;
}

static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 sub3(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static inline v3 add3(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static inline v3 scl3(float s, v3 a) { v3 r = { s * a.x, s * a.y, s * a.z }; return r; }
static inline float length3(v3 a) { return sqrtf(dot3(a, a)); }
/* glm distance(p0,p1) = length(p1 - p0) */
static inline float distance3(v3 p0, v3 p1) { return length3(sub3(p1, p0)); }
/* glm normalize(v) = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x) */
static inline v3 normalize3(v3 a) { float s = 1.0f / sqrtf(dot3(a, a)); v3 r = { a.x * s, a.y * s, a.z * s }; return r; }

/* ---- host-side scene helpers -------------------------------------------------------- */

#define ORC_DEG_TO_RAD 0.01745329251 /* Scene.h:12 (a double literal, cast to float) */

void orc_collider_build(const float trans[3], const float rot_deg[3], const float scale[3],
                        float out48[48]) {
    /* Scene.h:28-38 */
    float I[16], T[16], Rx[16], Ry[16], Rz[16], Sc[16], acc[16];
    static const float ax[3] = { 1.f, 0.f, 0.f }, ay[3] = { 0.f, 1.f, 0.f }, az[3] = { 0.f, 0.f, 1.f };
    mat_identity(I);
    mat_translate(I, trans, T);
    mat_rotate(rot_deg[0] * (float)ORC_DEG_TO_RAD, ax, Rx);
    mat_rotate(rot_deg[1] * (float)ORC_DEG_TO_RAD, ay, Ry);
    mat_rotate(rot_deg[2] * (float)ORC_DEG_TO_RAD, az, Rz);
    mat_scale(scale, Sc);
    mat_mul(T, Rz, acc);
    mat_mul(acc, Ry, acc);
    mat_mul(acc, Rx, acc);
    mat_mul(acc, Sc, acc);
    mat_mul(acc, I, acc);
    memcpy(out48, acc, 64);
    mat_inverse(out48, out48 + 16);
    mat_transpose(out48 + 16, out48 + 32);
}

void orc_collider_translate(float c48[48], const float translation[3]) {
    /* Scene.cpp:112-119 */
    float nt[16];
    mat_translate(c48, translation, nt);
    memcpy(c48, nt, 64);
    mat_inverse(c48, c48 + 16);
    mat_transpose(c48 + 16, c48 + 32);
}

void orc_default_colliders(float out[6 * 48]) {
    /* main.cpp:229-237: trans, rot (degrees), scale */
    static const float T[6][3] = { { 2.0f, 0.0f, 1.0f }, { 0.0f, 2.64f, 0.08f }, { 0.0f, 1.35f, -0.288f },
                                   { 0.0f, -0.380f, -0.116f }, { -0.698f, 0.087f, -0.36f }, { 0.698f, 0.087f, -0.36f } };
    static const float R[6][3] = { { 0.f, 0.f, 0.f }, { -38.270f, 0.0f, 0.0f }, { 18.301f, 0.0f, 0.0f },
                                   { -17.260f, 0.0f, 0.0f }, { -20.254f, 13.144f, 34.5f }, { -20.254f, 13.144f, -34.5f } };
    static const float S[6][3] = { { 1.f, 1.f, 1.f }, { 0.817f, 1.158f, 1.01f }, { 0.457f, 1.0f, 0.538f },
                                   { 1.078f, 1.683f, 0.974f }, { 0.721f, 1.0f, 0.724f }, { 0.721f, 1.0f, 0.724f } };
    for (int i = 0; i < 6; ++i) orc_collider_build(T[i], R[i], S[i], out + 48 * i);
}

void orc_default_params(orc_params* p, int num_strands, int num_points) {
    memset(p, 0, sizeof(*p));
    p->num_strands = num_strands;
    p->num_points = num_points;
    p->rest_length = 2.5f / ((float)num_points - 1.0f); /* compute.comp:139-140 */
    p->gravity_y = -9.8f;
    p->damping = 0.998f;
    p->vmax = 10.0f;
    p->penalty_k = 1900.0f;
    p->sphere_radius = 1.0f;
    p->grid_dim = 64;
    p->grid_extent = 7.0f;
    p->grid_origin[0] = -3.0f; p->grid_origin[1] = -2.0f; p->grid_origin[2] = -5.0f;
    p->grid_scale = 1000000.0f;
    p->friction = 0.08f;
    p->flags = ORC_GRID_ON;
    p->num_colliders = 6;
    p->repulsion = 0.2f;
}

void orc_init_strands_reference(int S, int N, const float* roots3, const float* normals3, float* strands) {
    /* Strand.cpp:157-175; `length / (NUM_CURVE_POINTS - 1.0)` is a double division cast to float */
    const float seg = (float)(2.5f / ((double)N - 1.0));
    for (int s = 0; s < S; ++s) {
        float* st = strands + (size_t)s * 12 * N;
        float cp[3] = { roots3[3 * s], roots3[3 * s + 1], roots3[3 * s + 2] };
        for (int j = 0; j < N; ++j) {
            float* P = st + 4 * j; float* V = st + 4 * (N + j); float* D = st + 4 * (2 * N + j);
            P[0] = cp[0]; P[1] = cp[1]; P[2] = cp[2]; P[3] = 1.0f;
            V[0] = 0.f; V[1] = 0.f; V[2] = -1.0f; V[3] = 0.f;
            D[0] = D[1] = D[2] = D[3] = 0.f;
            float dir[3] = { normals3[3 * s], normals3[3 * s + 1], normals3[3 * s + 2] };
            dir[2] -= 2.0f; dir[1] += 5.0f; dir[0] += 0.05f;
            cp[0] += seg * dir[0]; cp[1] += seg * dir[1]; cp[2] += seg * dir[2];
        }
    }
}

/* ---- wind helpers (compute.comp:83-121) ---------------------------------------------- */

static float glsl_fract(float x) { return x - floorf(x); }
static float wind_random(float px, float py) {
    float d = px * 12.9898f + py * 78.233f;
    return glsl_fract(sinf(d) * 43758.5453123f);
}
static float wind_noise(float px, float py) {
    float ix = floorf(px), iy = floorf(py);
    float fx = glsl_fract(px), fy = glsl_fract(py);
    float a = wind_random(ix, iy);
    float b = wind_random(ix + 1.0f, iy + 0.0f);
    float c = wind_random(ix + 0.0f, iy + 1.0f);
    float d = wind_random(ix + 1.0f, iy + 1.0f);
    float ux = fx * fx * (3.0f - 2.0f * fx);
    float uy = fy * fy * (3.0f - 2.0f * fy);
    /* glm mix(x,y,a) = x + a*(y-x) */
    float m = a + ux * (b - a);
    return m + (c - a) * uy * (1.0f - ux) + (d - b) * ux * uy;
}
float orc_fbm_time(float T) {
    float px = sinf(T), py = cosf(T);
    float value = 0.0f, amplitude = 0.5f;
    for (int i = 0; i < 6; ++i) {
        value += amplitude * wind_noise(px, py);
        px *= 2.0f; py *= 2.0f;
        amplitude *= 0.5f;
    }
    return value;
}

/* ---- head SDF (extension; device twin: rvh_kernels.cuh sdf_cell / sdf_trilinear) ------ */

static struct { const float* data; int nx, ny, nz; float origin[3], inv_cell; } g_sdf;

void orc_set_head_sdf(const float* sdf, const int dim[3], const float origin[3], float cell) {
    memset(&g_sdf, 0, sizeof g_sdf);
    if (!sdf) return;
    g_sdf.data = sdf; g_sdf.nx = dim[0]; g_sdf.ny = dim[1]; g_sdf.nz = dim[2];
    for (int k = 0; k < 3; ++k) g_sdf.origin[k] = origin[k];
    g_sdf.inv_cell = 1.0f / cell;
}

/* u = (p - origin) * inv_cell; in the volume when all 8 nodes of the cell exist; lerps are
 * fma(t, b - a, a), x then y then z -- the same operations in the same order as the kernel. */
int orc_sdf_sample(const float p[3], float* d_out, float grad[3]) {
    if (!g_sdf.data) return 0;
    const float ux = (p[0] - g_sdf.origin[0]) * g_sdf.inv_cell;
    const float uy = (p[1] - g_sdf.origin[1]) * g_sdf.inv_cell;
    const float uz = (p[2] - g_sdf.origin[2]) * g_sdf.inv_cell;
    if (!(ux >= 0.f && ux < (float)(g_sdf.nx - 1) && uy >= 0.f && uy < (float)(g_sdf.ny - 1) && uz >= 0.f && uz < (float)(g_sdf.nz - 1))) return 0;
    const int ix = (int)ux, iy = (int)uy, iz = (int)uz;
    const float tx = ux - (float)ix, ty = uy - (float)iy, tz = uz - (float)iz;
    const size_t sy = (size_t)g_sdf.nx, sz = (size_t)g_sdf.nx * g_sdf.ny;
    const float* g = g_sdf.data + ix + sy * iy + sz * iz;
    const float c0 = g[0], c1 = g[1], c2 = g[sy], c3 = g[sy + 1], c4 = g[sz], c5 = g[sz + 1], c6 = g[sz + sy], c7 = g[sz + sy + 1];
    const float dx00 = c1 - c0, dx10 = c3 - c2, dx01 = c5 - c4, dx11 = c7 - c6;
    const float c00 = fmaf(tx, dx00, c0), c10 = fmaf(tx, dx10, c2), c01 = fmaf(tx, dx01, c4), c11 = fmaf(tx, dx11, c6);
    const float dy0 = c10 - c00, dy1 = c11 - c01;
    const float e0 = fmaf(ty, dy0, c00), e1 = fmaf(ty, dy1, c01);
    const float gz = e1 - e0;
    *d_out = fmaf(tz, gz, e0);
    if (grad) {
        const float gx0 = fmaf(ty, dx10 - dx00, dx00), gx1 = fmaf(ty, dx11 - dx01, dx01);
        grad[0] = fmaf(tz, gx1 - gx0, gx0);
        grad[1] = fmaf(tz, dy1 - dy0, dy0);
        grad[2] = gz;
    }
    return 1;
}

/* ---- the compute pass --------------------------------------------------------------- */

static void integrate_strand(const orc_params* p, const float* col, float dt, float T, float fbmT, float* st) {
    const int N = p->num_points;
    float* P = st; float* V = st + 4 * N; float* D = st + 8 * N;
    const float radius = p->rest_length;
    for (int i = 1; i < N; ++i) {
        v3 cur = { P[4 * i], P[4 * i + 1], P[4 * i + 2] };
        v3 vel = { V[4 * i], V[4 * i + 1], V[4 * i + 2] };
        v3 parent = { P[4 * (i - 1)], P[4 * (i - 1) + 1], P[4 * (i - 1) + 2] };

        v3 force = { 0.0f, p->gravity_y, 0.0f };                               /* :150 */
        if (p->flags & ORC_WIND_A) {                                            /* :151 */
            float cl = cur.y * 2.0f; cl = cl < 0.2f ? 0.2f : (cl > 2.0f ? 2.0f : cl);
            v3 w = { 2.0f * sinf(T * 2.0f) * cosf(cur.y * 10.0f) * sinf((cur.y + 5.0f) * 15.0f), 0.0f, -cl };
            force = add3(force, scl3(10.0f, w));
        }
        if (p->flags & ORC_WIND_B) {                                            /* :152 */
            v3 w = { 2.0f * sinf(T * 2.0f) * cosf(cur.y * 10.0f) * sinf((cur.y + 5.0f) * 15.0f),
                     4.0f * sinf(cur.z * 5.0f + T * 3.0f),
                     -0.6f * (cur.y + 3.0f) };
            force = add3(force, scl3(7.0f * fbmT, w));
        }

        int hits = 0;
        v3 added = { 0.f, 0.f, 0.f };
        const int sdf_on = (p->flags & ORC_SDF_ON) != 0;
        if (sdf_on) {                                /* extension: the head volume stands in for colliders 1..n */
            const float pc[3] = { cur.x, cur.y, cur.z };
            float d, g[3];
            if (orc_sdf_sample(pc, &d, g) && d < 0.0f) {
                const float g2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
                if (g2 > 0.0f) {
                    const float sc = p->penalty_k * (-d) / sqrtf(g2);
                    added.x += sc * g[0]; added.y += sc * g[1]; added.z += sc * g[2];
                }
                ++hits;
            }
        }
        for (int j = 0; j < (sdf_on ? (p->num_colliders > 0 ? 1 : 0) : p->num_colliders); ++j) {   /* :158-180 */
            const float* c = col + 48 * j;
            if (j == 0) {
                v3 centre = { c[12], c[13], c[14] };
                const float r = p->sphere_radius;
                if (distance3(cur, centre) < r) {
                    float d = r - distance3(cur, centre);
                    v3 n = normalize3(sub3(cur, centre));
                    added = add3(added, scl3(p->penalty_k * d, n));
                    ++hits;
                }
            } else {
                v3 q = mat_mul_vec_xyz(c + 16, cur, 1.0f);                      /* :64-67 */
                v3 zero = { 0.f, 0.f, 0.f };
                if (distance3(q, zero) <= 1.0f) {
                    v3 on = mat_mul_vec_xyz(c, normalize3(q), 1.0f);            /* :76-79 */
                    float d = distance3(on, cur);
                    v3 nn = normalize3(mat_mul_vec_xyz(c + 32, q, 0.0f));       /* :70-73 */
                    nn = normalize3(nn);                                        /* :174 normalises again */
                    added = add3(added, scl3(p->penalty_k * d, nn));
                    ++hits;
                }
            }
        }
        if (hits > 0) {                                                          /* :182-184 */
            float fh = (float)hits;
            v3 a = { added.x / fh, added.y / fh, added.z / fh };
            force = add3(force, a);
        }

        v3 pred = add3(add3(cur, scl3(dt, vel)), scl3(dt * dt, force));         /* :187 */
        v3 dir = normalize3(sub3(pred, parent));                                /* :191 */
        v3 np = add3(parent, scl3(radius, dir));                                /* :192 */
        v3 nv = sub3(np, cur);                                                  /* :195 */
        nv.x = nv.x / dt; nv.y = nv.y / dt; nv.z = nv.z / dt;
        P[4 * i] = np.x; P[4 * i + 1] = np.y; P[4 * i + 2] = np.z; P[4 * i + 3] = 1.0f;
        float vx = p->damping * nv.x, vy = p->damping * nv.y, vz = p->damping * nv.z, vw = p->damping * 0.0f;
        /* length(vec4): (x*x + y*y) + (z*z + w*w) */
        float l2 = (vx * vx + vy * vy) + (vz * vz + vw * vw);
        if (sqrtf(l2) > p->vmax) {                                              /* :198-200 */
            float s = 1.0f / sqrtf(l2);
            vx = vx * s * p->vmax; vy = vy * s * p->vmax; vz = vz * s * p->vmax; vw = vw * s * p->vmax;
        }
        V[4 * i] = vx; V[4 * i + 1] = vy; V[4 * i + 2] = vz; V[4 * i + 3] = vw;
        D[4 * i] = p->damping * (np.x - pred.x);                                /* :201 */
        D[4 * i + 1] = p->damping * (np.y - pred.y);
        D[4 * i + 2] = p->damping * (np.z - pred.z);
        D[4 * i + 3] = p->damping * 0.0f;
    }
}

/* The collider decisions of integrate_strand (compute.comp:162, :66; SDF mode: d < 0) for every point of every strand, without
 * stepping: out[s*N + i] bit 0 = sphere, bit j = collider j (SDF mode: bit 1 = head volume).  Collision tests read the OLD
 * position of point i (compute.comp:147), so the masks depend on the input state only. */
void orc_hit_masks(const orc_params* p, const float* col, const float* strands, unsigned char* out) {
    const int N = p->num_points;
    const int sdf_on = (p->flags & ORC_SDF_ON) != 0;
    for (int s = 0; s < p->num_strands; ++s) {
        const float* P = strands + (size_t)s * 12 * N;
        out[(size_t)s * N] = 0;
        for (int i = 1; i < N; ++i) {
            v3 cur = { P[4 * i], P[4 * i + 1], P[4 * i + 2] };
            unsigned m = 0;
            if (sdf_on) {
                const float pc[3] = { cur.x, cur.y, cur.z };
                float d, g[3];
                if (orc_sdf_sample(pc, &d, g) && d < 0.0f) m |= 2u;
            }
            for (int j = 0; j < (sdf_on ? (p->num_colliders > 0 ? 1 : 0) : p->num_colliders); ++j) {
                const float* c = col + 48 * j;
                if (j == 0) {
                    v3 centre = { c[12], c[13], c[14] };
                    if (distance3(cur, centre) < p->sphere_radius) m |= 1u;
                } else {
                    v3 q = mat_mul_vec_xyz(c + 16, cur, 1.0f);
                    v3 zero = { 0.f, 0.f, 0.f };
                    if (distance3(q, zero) <= 1.0f) m |= 1u << j;
                }
            }
            out[(size_t)s * N + i] = (unsigned char)m;
        }
    }
}

typedef struct { int lo[3], hi[3]; float g[3]; } cellrange;

static inline cellrange cell_range(const orc_params* p, const float* pos) {
    cellrange c;
    const float h = p->grid_extent / (float)p->grid_dim;                        /* :205 */
    for (int k = 0; k < 3; ++k) {
        float g = (pos[k] - p->grid_origin[k]) / h;                             /* :219-221 */
        float fl = floorf(g);
        /* int(floor(x)): clamp first so the conversion is defined for far-away points */
        if (fl < -2.0f) fl = -2.0f;
        if (fl > (float)p->grid_dim) fl = (float)p->grid_dim;
        if (!(fl == fl)) fl = (float)p->grid_dim; /* NaN -> no cells */
        int f = (int)fl;
        c.g[k] = g;
        c.lo[k] = f > 0 ? f : 0;                                                /* :224-229 */
        c.hi[k] = (f + 1) < (p->grid_dim - 1) ? (f + 1) : (p->grid_dim - 1);
    }
    return c;
}

static inline float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }

static void splat_strand(const orc_params* p, float dt, float* st, int64_t* grid) {
    const int N = p->num_points, G = p->grid_dim;
    float* P = st; float* V = st + 4 * N; float* D = st + 8 * N;
    for (int i = 1; i < N; ++i) {
        if (i != N - 1) {                                                       /* :213-215 */
            V[4 * i]     -= D[4 * (i + 1)]     / dt;
            V[4 * i + 1] -= D[4 * (i + 1) + 1] / dt;
            V[4 * i + 2] -= D[4 * (i + 1) + 2] / dt;
            V[4 * i + 3] -= 0.0f;
        }
        cellrange c = cell_range(p, P + 4 * i);
        for (int a = c.lo[0]; a <= c.hi[0]; ++a)
            for (int b = c.lo[1]; b <= c.hi[1]; ++b)
                for (int cc = c.lo[2]; cc <= c.hi[2]; ++cc) {
                    int index = a + b * G + cc * G * G;                         /* :234 */
                    float xw = clamp01(1.0f - fabsf(c.g[0] - (float)a));
                    float yw = clamp01(1.0f - fabsf(c.g[1] - (float)b));
                    float zw = clamp01(1.0f - fabsf(c.g[2] - (float)cc));
                    float tw = xw * yw * zw;
                    float wx = tw * V[4 * i], wy = tw * V[4 * i + 1], wz = tw * V[4 * i + 2];
                    int64_t* cell = grid + 4 * (size_t)index;                   /* :245-248 */
                    cell[0] += (int64_t)(p->grid_scale * wx);
                    cell[1] += (int64_t)(p->grid_scale * wy);
                    cell[2] += (int64_t)(p->grid_scale * wz);
                    cell[3] += (int64_t)(p->grid_scale * tw);
                }
    }
}

static void gather_strand(const orc_params* p, float* st, const int64_t* grid) {
    const int N = p->num_points, G = p->grid_dim;
    const int wrap = p->flags & ORC_GRID_INT32_WRAP;
    float* P = st; float* V = st + 4 * N;
    for (int i = 1; i < N; ++i) {
        cellrange c = cell_range(p, P + 4 * i);
        v3 gv = { 0.f, 0.f, 0.f };
        float rho = 0.f, rg[3] = { 0.f, 0.f, 0.f };         /* extension: corner-density sum and h * grad(rho) */
        const float fl[3] = { floorf(c.g[0]), floorf(c.g[1]), floorf(c.g[2]) };
        for (int a = c.lo[0]; a <= c.hi[0]; ++a)
            for (int b = c.lo[1]; b <= c.hi[1]; ++b)
                for (int cc = c.lo[2]; cc <= c.hi[2]; ++cc) {
                    int index = a + b * G + cc * G * G;
                    const int64_t* cell = grid + 4 * (size_t)index;
                    int64_t dens = cell[3], v0 = cell[0], v1 = cell[1], v2 = cell[2];
                    if (wrap) { dens = (int32_t)(uint32_t)dens; v0 = (int32_t)(uint32_t)v0; v1 = (int32_t)(uint32_t)v1; v2 = (int32_t)(uint32_t)v2; }
                    if (dens > 0) {                                             /* :276 */
                        float xw = clamp01(1.0f - fabsf(c.g[0] - (float)a));
                        float yw = clamp01(1.0f - fabsf(c.g[1] - (float)b));
                        float zw = clamp01(1.0f - fabsf(c.g[2] - (float)cc));
                        float tw = xw * yw * zw;
                        float s = tw * (1.0f / (float)dens);                    /* :283 */
                        gv.x += s * (float)v0; gv.y += s * (float)v1; gv.z += s * (float)v2;
                        if (p->flags & ORC_REPULSION_ON) {
                            /* d/dg clamp01(1-|g-a|) = -1 for the cell at floor(g), +1 for the next one */
                            const float D = (float)dens;
                            rho += D;
                            rg[0] += ((float)a  > fl[0] ? D : -D) * (yw * zw);
                            rg[1] += ((float)b  > fl[1] ? D : -D) * (xw * zw);
                            rg[2] += ((float)cc > fl[2] ? D : -D) * (xw * yw);
                        }
                    }
                }
        const float fr = p->friction, omf = 1.0f - fr;                          /* :296-297 */
        V[4 * i]     = omf * V[4 * i]     + fr * gv.x;
        V[4 * i + 1] = omf * V[4 * i + 1] + fr * gv.y;
        V[4 * i + 2] = omf * V[4 * i + 2] + fr * gv.z;
        V[4 * i + 3] = 0.0f;
        if ((p->flags & ORC_REPULSION_ON) && rho > 0.0f) {
            const float k = -p->repulsion / rho;            /* rho = sum of the corner densities: |rg| <= rho per component */
            V[4 * i] += k * rg[0]; V[4 * i + 1] += k * rg[1]; V[4 * i + 2] += k * rg[2];
        }
    }
}

void orc_phase_integrate(const orc_params* p, const float* col, float dt, float T, float* strands) {
    const float fbmT = (p->flags & ORC_WIND_B) ? orc_fbm_time(T) : 0.0f;
    const size_t stride = (size_t)12 * p->num_points;
    for (int s = 0; s < p->num_strands; ++s) integrate_strand(p, col, dt, T, fbmT, strands + s * stride);
}

void orc_phase_splat(const orc_params* p, float dt, float* strands, int64_t* grid) {
    const size_t stride = (size_t)12 * p->num_points;
    for (int s = 0; s < p->num_strands; ++s) splat_strand(p, dt, strands + s * stride, grid);
}

void orc_phase_gather(const orc_params* p, float* strands, const int64_t* grid) {
    const size_t stride = (size_t)12 * p->num_points;
    for (int s = 0; s < p->num_strands; ++s) gather_strand(p, strands + s * stride, grid);
}

/* Grid off: the shader always runs P2/P3; with ORC_GRID_ON clear the velocity correction
 * (compute.comp:213-215) is still applied, only the splat/gather are skipped. */
static void correct_strand(const orc_params* p, float dt, float* st) {
    const int N = p->num_points;
    float* V = st + 4 * N; float* D = st + 8 * N;
    for (int i = 1; i < N - 1; ++i) {
        V[4 * i]     -= D[4 * (i + 1)]     / dt;
        V[4 * i + 1] -= D[4 * (i + 1) + 1] / dt;
        V[4 * i + 2] -= D[4 * (i + 1) + 2] / dt;
    }
}

void orc_step(const orc_params* p, const float* col, float dt, float T, float* strands, int64_t* grid) {
    const size_t cells = (size_t)p->grid_dim * p->grid_dim * p->grid_dim;
    const size_t stride = (size_t)12 * p->num_points;
    orc_phase_integrate(p, col, dt, T, strands);
    if (p->flags & ORC_GRID_ON) {
        memset(grid, 0, cells * 4 * sizeof(int64_t));                           /* Renderer.cpp:2063 */
        orc_phase_splat(p, dt, strands, grid);
        orc_phase_gather(p, strands, grid);
    } else {
        for (int s = 0; s < p->num_strands; ++s) correct_strand(p, dt, strands + s * stride);
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_step_parallel(const orc_params* p, const float* col, float dt, float T, float* strands,
                       int64_t* grid, int num_threads) {
#ifdef _OPENMP
    const size_t cells4 = (size_t)p->grid_dim * p->grid_dim * p->grid_dim * 4;
    const size_t stride = (size_t)12 * p->num_points;
    const float fbmT = (p->flags & ORC_WIND_B) ? orc_fbm_time(T) : 0.0f;
    const int S = p->num_strands;
    if (num_threads <= 0) num_threads = omp_get_max_threads();
    if (!(p->flags & ORC_GRID_ON)) {
        #pragma omp parallel for num_threads(num_threads) schedule(static)
        for (int s = 0; s < S; ++s) {
            integrate_strand(p, col, dt, T, fbmT, strands + s * stride);
            correct_strand(p, dt, strands + s * stride);
        }
        return;
    }
    int64_t* priv = (int64_t*)calloc(cells4 * (size_t)num_threads, sizeof(int64_t));
    #pragma omp parallel num_threads(num_threads)
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
        int64_t* mine = priv + cells4 * (size_t)t;
        #pragma omp for schedule(static)
        for (int s = 0; s < S; ++s) {
            integrate_strand(p, col, dt, T, fbmT, strands + s * stride);
            splat_strand(p, dt, strands + s * stride, mine);
        }
        #pragma omp for schedule(static)
        for (size_t k = 0; k < cells4; ++k) {
            int64_t acc = 0;
            for (int j = 0; j < nt; ++j) acc += priv[cells4 * (size_t)j + k];
            grid[k] = acc;
        }
        #pragma omp for schedule(static)
        for (int s = 0; s < S; ++s) gather_strand(p, strands + s * stride, grid);
    }
    free(priv);
#else
    (void)num_threads;
    orc_step(p, col, dt, T, strands, grid);
#endif
}

/* ---- SDF bakes (extension; device twins k_sdf_bake_colliders / k_sdf_bake_mesh) ------------ */

void orc_sdf_bake_colliders(const float* col, int num_colliders, const int dim[3], const float origin[3],
                            float cell, float* out) {
    for (int k = 0; k < dim[2]; ++k)
        for (int j = 0; j < dim[1]; ++j)
            for (int i = 0; i < dim[0]; ++i) {
                v3 p = { origin[0] + cell * (float)i, origin[1] + cell * (float)j, origin[2] + cell * (float)k };
                float best = 1.0e9f;
                for (int e = 1; e < num_colliders; ++e) {
                    const float* c = col + 48 * e;
                    v3 q = mat_mul_vec_xyz(c + 16, p, 1.0f);
                    float q2 = q.x * q.x + q.y * q.y + q.z * q.z;
                    v3 u = { 1.f, 0.f, 0.f };
                    if (q2 > 0.f) { float r = 1.0f / sqrtf(q2); u.x = q.x * r; u.y = q.y * r; u.z = q.z * r; }
                    v3 on = mat_mul_vec_xyz(c, u, 1.0f);
                    float dist = distance3(on, p);
                    float sd = q2 <= 1.0f ? -dist : dist;
                    if (sd < best) best = sd;
                }
                out[i + (size_t)dim[0] * (j + (size_t)dim[1] * k)] = best;
            }
}

static float tri_dist2(v3 p, v3 a, v3 b, v3 c) {       /* Ericson, Real-Time Collision Detection 5.1.5 */
    v3 ab = sub3(b, a), ac = sub3(c, a), ap = sub3(p, a);
    float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    v3 cp_;
    if (d1 <= 0.f && d2 <= 0.f) { cp_.x = cp_.y = cp_.z = 0.f; }
    else {
        v3 bp = sub3(p, b);
        float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
        if (d3 >= 0.f && d4 <= d3) cp_ = ab;
        else {
            float vc = d1 * d4 - d3 * d2;
            if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) cp_ = scl3(d1 / (d1 - d3), ab);
            else {
                v3 cpv = sub3(p, c);
                float d5 = dot3(ab, cpv), d6 = dot3(ac, cpv);
                if (d6 >= 0.f && d5 <= d6) cp_ = ac;
                else {
                    float vb = d5 * d2 - d1 * d6;
                    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) cp_ = scl3(d2 / (d2 - d6), ac);
                    else {
                        float va = d3 * d6 - d5 * d4;
                        if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
                            float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
                            cp_ = add3(ab, scl3(w, sub3(ac, ab)));
                        } else {
                            float den = 1.0f / (va + vb + vc);
                            cp_ = add3(scl3(vb * den, ab), scl3(vc * den, ac));
                        }
                    }
                }
            }
        }
    }
    v3 e = sub3(ap, cp_);
    return dot3(e, e);
}

static float tri_solid_angle(v3 p, v3 A, v3 B, v3 C) {  /* Van Oosterom & Strackee 1983 */
    v3 a = sub3(A, p), b = sub3(B, p), c = sub3(C, p);
    float la = sqrtf(dot3(a, a)), lb = sqrtf(dot3(b, b)), lc = sqrtf(dot3(c, c));
    float num = a.x * (b.y * c.z - b.z * c.y) + a.y * (b.z * c.x - b.x * c.z) + a.z * (b.x * c.y - b.y * c.x);
    float den = la * lb * lc + dot3(a, b) * lc + dot3(b, c) * la + dot3(c, a) * lb;
    return 2.0f * atan2f(num, den);
}

void orc_sdf_bake_mesh(const float* verts, const int* tris, int ntris, const int dim[3], const float origin[3],
                       float cell, float* out) {
    #pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < dim[2]; ++k)
        for (int j = 0; j < dim[1]; ++j)
            for (int i = 0; i < dim[0]; ++i) {
                v3 p = { origin[0] + cell * (float)i, origin[1] + cell * (float)j, origin[2] + cell * (float)k };
                float best2 = 3.0e38f, omega = 0.f;
                for (int t = 0; t < ntris; ++t) {
                    const float* a = verts + 3 * (size_t)tris[3 * t], *b = verts + 3 * (size_t)tris[3 * t + 1], *c = verts + 3 * (size_t)tris[3 * t + 2];
                    v3 A = { a[0], a[1], a[2] }, B = { b[0], b[1], b[2] }, C = { c[0], c[1], c[2] };
                    float d2 = tri_dist2(p, A, B, C);
                    if (d2 < best2) best2 = d2;
                    omega += tri_solid_angle(p, A, B, C);
                }
                float d = sqrtf(best2);
                out[i + (size_t)dim[0] * (j + (size_t)dim[1] * k)] = fabsf(omega) > 6.2831855f ? -d : d;   /* |winding| > 1/2 */
            }
}

/* ---- hair.tese restated (vertex placement only; shading outputs depend on the cameras) ------ */

static float tese_random(float x, float y) {                /* hair.tese:30-32, sine in double */
    const float a = x * 12.9898f, b = y * 78.233f;
    const float d = a + b;
    const float h = (float)sin((double)d) * 43758.5453123f;
    return h - floorf(h);
}
static float tese_mix(float a, float b, float t) { return a + t * (b - a); }

void orc_expand_strands(const float* strands, int S, int N, int isolines, int divisions,
                        float* pos_width, float* tangent_u) {
    const float PI = 3.141592653f;                          /* hair.tese:3 */
    #pragma omp parallel for schedule(static)
    for (int s = 0; s < S; ++s) {
        const float* P = strands + (size_t)s * 12 * N;      /* curvePoints[N] as vec4 */
        for (int k = 0; k < isolines; ++k) {
            const float u = (float)k / (float)isolines;     /* gl_TessCoord.y */
            const float rand2 = fabsf(tese_random(u, u * u));                   /* :226 */
            const float uRad = 2.0f * PI * u;                                   /* :235 */
            const float cx = cosf(uRad), sz = sinf(uRad);
            const float len = sqrtf(cx * cx + sz * sz);
            const float dirx = cx / len, dirz = sz / len;                       /* :236 */
            const float choice = tese_random(u, P[0]) * tese_random(P[1], P[2]); /* :245 */
            for (int j = 0; j <= divisions; ++j) {
                const float v = (float)j / (float)divisions; /* gl_TessCoord.x */
                const float vs = v * (float)(N - 1);
                int seg = (int)floorf(vs);                                       /* :40, :165 */
                if (seg > N - 2) seg = N - 2;               /* v = 1: the shader reads point N */
                const float t = vs - (float)seg;                                /* :65 */
                float c[3], tg[3];
                for (int a = 0; a < 3; ++a) {               /* func(), :34-80 */
                    const float v1 = P[4 * seg + a], v2 = P[4 * (seg + 1) + a];
                    const float v0 = seg == 0 ? v1 + (v1 - v2) : P[4 * (seg - 1) + a];
                    const float v3 = seg + 1 == N - 1 ? v2 + (v2 - v1) : P[4 * (seg + 2) + a];
                    const float b1 = v1 + (1.0f / 3.0f) * ((v2 - v0) / 2.0f);
                    const float b2 = v2 - (1.0f / 3.0f) * ((v3 - v1) / 2.0f);
                    const float b01 = tese_mix(v1, b1, t), b11 = tese_mix(b1, b2, t), b21 = tese_mix(b2, v2, t);
                    const float b02 = tese_mix(b01, b11, t), b12 = tese_mix(b11, b21, t);
                    c[a] = tese_mix(b02, b12, t);
                    tg[a] = v2 - v1;                                            /* :267 */
                }
                float width = 0.5f * tese_mix(tese_mix(0.05f, 0.3f, v), tese_mix(0.3f, 0.1f, v), v) * (rand2 + 0.5f);   /* :228-229 */
                float sd = 1.0f;                                                /* :239-269 */
                const float two_s2 = 2.0f * powf(0.2f, 2.0f);
                if (choice > 0.5f) {
                    if (choice > 0.9f) sd = 1.8f * expf(-powf(v - 0.25f, 2.0f) / two_s2);
                    else if (choice > 0.8f) sd = 4.5f * powf(v, 10.0f);
                    else if (choice > 0.7f) sd = 2.5f * expf(-powf(v - 0.7f, 2.0f) / two_s2);
                    else if (choice > 0.6f) sd = 4.0f * powf(v, 1.3f);
                    else sd = 1.8f * expf(-powf(v - 0.8f, 2.0f) / two_s2);
                }
                if (v == 0.0f) sd = 1.0f;
                const size_t o = (((size_t)s * isolines + k) * (size_t)(divisions + 1) + j) * 4;
                pos_width[o]     = c[0] + width * (dirx * sd);                  /* :278, :304 */
                pos_width[o + 1] = c[1] + width * (0.0f * sd);
                pos_width[o + 2] = c[2] + width * (dirz * sd);
                pos_width[o + 3] = tese_mix(0.02f, 0.01f, v);                   /* :313-315 */
                const float tl = sqrtf(tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2]);
                tangent_u[o] = tg[0] / tl; tangent_u[o + 1] = tg[1] / tl; tangent_u[o + 2] = tg[2] / tl;
                tangent_u[o + 3] = u;
            }
        }
    }
}
