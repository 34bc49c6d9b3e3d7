// rvh_host_math.h -- host-side scene maths of the drop-in layer (no GPU, no glm).
//
// The reference builds its collider matrices with glm (Scene.h:28-38, Scene.cpp:110-120)
// and its wind scalar in the shader (compute.comp:83-121).  These are the product's own
// single-precision implementations of the same formulas; tests compare them with the
// oracle / the reference-compiled glm build.
#pragma once
#include <cmath>
#include <cstring>

namespace rvh {

struct Mat4 {
    float m[16];  // column-major, m[c*4+r]
    static Mat4 identity() { Mat4 r; std::memset(r.m, 0, sizeof r.m); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }
    float& at(int c, int r) { return m[c * 4 + r]; }
    float at(int c, int r) const { return m[c * 4 + r]; }
};

inline Mat4 mul(const Mat4& a, const Mat4& b) {
    Mat4 o;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float acc = a.at(0, r) * b.at(c, 0);
            acc += a.at(1, r) * b.at(c, 1);
            acc += a.at(2, r) * b.at(c, 2);
            acc += a.at(3, r) * b.at(c, 3);
            o.at(c, r) = acc;
        }
    return o;
}

inline Mat4 translated(const Mat4& a, const float v[3]) {   // a * T(v)
    Mat4 o = a;
    for (int r = 0; r < 4; ++r) {
        float acc = a.at(0, r) * v[0];
        acc += a.at(1, r) * v[1];
        acc += a.at(2, r) * v[2];
        acc += a.at(3, r);
        o.at(3, r) = acc;
    }
    return o;
}

inline Mat4 axis_rotation(float angle, int axis) {          // unit axis x/y/z, Rodrigues form
    const float c = std::cos(angle), s = std::sin(angle), t = 1.0f - c;
    float a[3] = { 0.f, 0.f, 0.f };
    a[axis] = 1.f;
    Mat4 o = Mat4::identity();
    o.at(0, 0) = c + t * a[0] * a[0];        o.at(0, 1) = t * a[0] * a[1] + s * a[2]; o.at(0, 2) = t * a[0] * a[2] - s * a[1];
    o.at(1, 0) = t * a[1] * a[0] - s * a[2]; o.at(1, 1) = c + t * a[1] * a[1];        o.at(1, 2) = t * a[1] * a[2] + s * a[0];
    o.at(2, 0) = t * a[2] * a[0] + s * a[1]; o.at(2, 1) = t * a[2] * a[1] - s * a[0]; o.at(2, 2) = c + t * a[2] * a[2];
    return o;
}

inline Mat4 scaling(const float v[3]) {
    Mat4 o = Mat4::identity();
    o.at(0, 0) = v[0]; o.at(1, 1) = v[1]; o.at(2, 2) = v[2];
    return o;
}

// General 4x4 inverse by cofactors (adjugate / determinant), single precision.
inline Mat4 inverse(const Mat4& a) {
    const float* m = a.m;
    // 2x2 sub-determinants of rows 2,3 (column pairs), same grouping glm uses
    auto M = [&](int c, int r) { return m[c * 4 + r]; };
    const float c00 = M(2,2) * M(3,3) - M(3,2) * M(2,3), c02 = M(1,2) * M(3,3) - M(3,2) * M(1,3), c03 = M(1,2) * M(2,3) - M(2,2) * M(1,3);
    const float c04 = M(2,1) * M(3,3) - M(3,1) * M(2,3), c06 = M(1,1) * M(3,3) - M(3,1) * M(1,3), c07 = M(1,1) * M(2,3) - M(2,1) * M(1,3);
    const float c08 = M(2,1) * M(3,2) - M(3,1) * M(2,2), c10 = M(1,1) * M(3,2) - M(3,1) * M(1,2), c11 = M(1,1) * M(2,2) - M(2,1) * M(1,2);
    const float c12 = M(2,0) * M(3,3) - M(3,0) * M(2,3), c14 = M(1,0) * M(3,3) - M(3,0) * M(1,3), c15 = M(1,0) * M(2,3) - M(2,0) * M(1,3);
    const float c16 = M(2,0) * M(3,2) - M(3,0) * M(2,2), c18 = M(1,0) * M(3,2) - M(3,0) * M(1,2), c19 = M(1,0) * M(2,2) - M(2,0) * M(1,2);
    const float c20 = M(2,0) * M(3,1) - M(3,0) * M(2,1), c22 = M(1,0) * M(3,1) - M(3,0) * M(1,1), c23 = M(1,0) * M(2,1) - M(2,0) * M(1,1);
    const float f0[4] = { c00, c00, c02, c03 }, f1[4] = { c04, c04, c06, c07 }, f2[4] = { c08, c08, c10, c11 };
    const float f3[4] = { c12, c12, c14, c15 }, f4[4] = { c16, c16, c18, c19 }, f5[4] = { c20, c20, c22, c23 };
    const float v0[4] = { M(1,0), M(0,0), M(0,0), M(0,0) }, v1[4] = { M(1,1), M(0,1), M(0,1), M(0,1) };
    const float v2[4] = { M(1,2), M(0,2), M(0,2), M(0,2) }, v3[4] = { M(1,3), M(0,3), M(0,3), M(0,3) };
    Mat4 adj;
    for (int r = 0; r < 4; ++r) {
        const float sa = (r & 1) ? -1.f : 1.f, sb = -sa;
        adj.at(0, r) = ((v1[r] * f0[r] - v2[r] * f1[r]) + v3[r] * f2[r]) * sa;
        adj.at(1, r) = ((v0[r] * f0[r] - v2[r] * f3[r]) + v3[r] * f4[r]) * sb;
        adj.at(2, r) = ((v0[r] * f1[r] - v1[r] * f3[r]) + v3[r] * f5[r]) * sa;
        adj.at(3, r) = ((v0[r] * f2[r] - v1[r] * f4[r]) + v2[r] * f5[r]) * sb;
    }
    const float det = (M(0,0) * adj.at(0, 0) + M(0,1) * adj.at(1, 0)) + (M(0,2) * adj.at(2, 0) + M(0,3) * adj.at(3, 0));
    const float inv_det = 1.0f / det;
    Mat4 o;
    for (int i = 0; i < 16; ++i) o.m[i] = adj.m[i] * inv_det;
    return o;
}

inline Mat4 transpose(const Mat4& a) {
    Mat4 o;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) o.at(c, r) = a.at(r, c);
    return o;
}

// Collider(trans, rot_deg, scale): transform = T * Rz * Ry * Rx * S   (Scene.h:28-38)
inline void collider_build(const float t[3], const float rot_deg[3], const float sc[3], float out48[48]) {
    const float d2r = (float)0.01745329251;                                // Scene.h:12
    Mat4 T = translated(Mat4::identity(), t);
    Mat4 X = mul(mul(mul(mul(T, axis_rotation(rot_deg[2] * d2r, 2)), axis_rotation(rot_deg[1] * d2r, 1)),
                     axis_rotation(rot_deg[0] * d2r, 0)), scaling(sc));
    Mat4 I = inverse(X), IT = transpose(I);
    std::memcpy(out48, X.m, 64); std::memcpy(out48 + 16, I.m, 64); std::memcpy(out48 + 32, IT.m, 64);
}

inline void collider_translate(float c48[48], const float tr[3]) {         // Scene.cpp:112-119
    Mat4 X; std::memcpy(X.m, c48, 64);
    Mat4 Y = translated(X, tr), I = inverse(Y), IT = transpose(I);
    std::memcpy(c48, Y.m, 64); std::memcpy(c48 + 16, I.m, 64); std::memcpy(c48 + 32, IT.m, 64);
}

// fbm(vec2(sin T, cos T)) of compute.comp:83-121: 6 octaves of value noise on a
// fract(sin(dot)*43758.5453) hash.  It depends on time only, so the host evaluates it once
// per step and the kernel receives a scalar.
inline float wind_hash(float x, float y) {
    const float v = std::sin(x * 12.9898f + y * 78.233f) * 43758.5453123f;
    return v - std::floor(v);
}
inline float wind_noise(float x, float y) {
    const float ix = std::floor(x), iy = std::floor(y), fx = x - ix, fy = y - iy;
    const float a = wind_hash(ix, iy), b = wind_hash(ix + 1.f, iy), c = wind_hash(ix, iy + 1.f), d = wind_hash(ix + 1.f, iy + 1.f);
    const float ux = fx * fx * (3.f - 2.f * fx), uy = fy * fy * (3.f - 2.f * fy);
    return (a + ux * (b - a)) + (c - a) * uy * (1.f - ux) + (d - b) * ux * uy;
}
inline float wind_fbm(float T) {
    float x = std::sin(T), y = std::cos(T), value = 0.f, amp = 0.5f;
    for (int o = 0; o < 6; ++o) { value += amp * wind_noise(x, y); x *= 2.f; y *= 2.f; amp *= 0.5f; }
    return value;
}

}  // namespace rvh
