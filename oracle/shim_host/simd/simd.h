/* oracle/shim_host/simd/simd.h -- TEST INFRASTRUCTURE.
 * Host C++ stand-in for the slice of Apple's <simd/simd.h> that the HOST branch of the reference's headers uses
 * (RT_Metal/Metal/Common.hh:17-38, AABB.hh:213-253, BVH.hh:30-314), so that the reference's own BVH builder can be
 * compiled here. Strict IEEE fp32, one rounding per operation; simd_float3 is 16 bytes like Apple's. simd_mul(matrix,
 * vector) accumulates column by column, ((c0*x + c1*y) + c2*z) + c3*w, the order of Apple's simd/matrix.h without
 * fusing (Apple's may fuse the multiply-adds; the reference defines no bit pattern there). */
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

typedef unsigned int uint;

struct simd_float2 {
    float x, y;
    simd_float2() : x(0), y(0) {}
    simd_float2(float s) : x(s), y(s) {}
    simd_float2(float a, float b) : x(a), y(b) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
} __attribute__((aligned(8)));

struct simd_float3 {
    float x, y, z, _pad;
    simd_float3() : x(0), y(0), z(0), _pad(0) {}
    simd_float3(float s) : x(s), y(s), z(s), _pad(0) {}
    simd_float3(float a, float b, float c) : x(a), y(b), z(c), _pad(0) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
} __attribute__((aligned(16)));

struct simd_float4 {
    float x, y, z, w;
    simd_float4() : x(0), y(0), z(0), w(0) {}
    simd_float4(float s) : x(s), y(s), z(s), w(s) {}
    simd_float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
} __attribute__((aligned(16)));

#define TRQ_SIMD_OPS(T, N)                                                                                      \
    inline T operator+(const T& a, const T& b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }  \
    inline T operator-(const T& a, const T& b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }  \
    inline T operator*(const T& a, const T& b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; }  \
    inline T operator/(const T& a, const T& b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; }  \
    inline T operator*(const T& a, float s) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }        \
    inline T operator*(float s, const T& a) { T r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r; }        \
    inline T operator/(const T& a, float s) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }        \
    inline T operator-(const T& a) { T r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }                    \
    inline T& operator+=(T& a, const T& b) { a = a + b; return a; }                                               \
    inline T& operator-=(T& a, const T& b) { a = a - b; return a; }                                               \
    inline T& operator*=(T& a, float s) { a = a * s; return a; }
TRQ_SIMD_OPS(simd_float2, 2)
TRQ_SIMD_OPS(simd_float3, 3)
TRQ_SIMD_OPS(simd_float4, 4)
#undef TRQ_SIMD_OPS

inline simd_float2 simd_make_float2(float x, float y) { return simd_float2(x, y); }
inline simd_float3 simd_make_float3(float x, float y, float z) { return simd_float3(x, y, z); }
inline simd_float4 simd_make_float4(float x, float y, float z, float w) { return simd_float4(x, y, z, w); }
inline simd_float4 simd_make_float4(simd_float3 v, float w) { return simd_float4(v.x, v.y, v.z, w); }

struct simd_float2x2 { simd_float2 columns[2]; };
struct simd_float3x3 { simd_float3 columns[3]; };
struct simd_float4x4 { simd_float4 columns[4]; };

static const simd_float4x4 matrix_identity_float4x4 = {{simd_float4(1, 0, 0, 0), simd_float4(0, 1, 0, 0), simd_float4(0, 0, 1, 0), simd_float4(0, 0, 0, 1)}};

inline simd_float4 simd_mul(const simd_float4x4& m, const simd_float4& v) {
    simd_float4 r = m.columns[0] * v.x;
    r = r + m.columns[1] * v.y;
    r = r + m.columns[2] * v.z;
    r = r + m.columns[3] * v.w;
    return r;
}

/* ---- what RT_Metal/Tracer/Tracer.mm uses on top of the above (scene constants, camera) -------------------------
 * Apple's definitions, unfused fp32: dot = (x*x' + y*y') + z*z'; normalize(v) = v * (1 / sqrt(dot(v, v))) (simd/geometry.h,
 * precise variant); matrix product column by column; 4x4 inverse by cofactors in double (Apple's is closed source: the
 * inverse / normal matrices are compared to a tolerance, everything else exactly). */
inline float simd_dot(const simd_float3& a, const simd_float3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline simd_float3 simd_cross(const simd_float3& a, const simd_float3& b) {
    return simd_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float simd_length_squared(const simd_float3& a) { return simd_dot(a, a); }
inline float simd_length(const simd_float3& a) { return sqrtf(simd_dot(a, a)); }
inline simd_float3 simd_normalize(const simd_float3& a) { return a * (1.0f / sqrtf(simd_length_squared(a))); }
inline simd_float3 vector_normalize(const simd_float3& a) { return simd_normalize(a); }
inline simd_float3 simd_make_float3(float s) { return simd_float3(s); }

inline simd_float4x4 simd_transpose(const simd_float4x4& m) {
    simd_float4x4 r;
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r.columns[c][k] = m.columns[k][c];
    return r;
}
inline simd_float4x4 simd_mul(const simd_float4x4& a, const simd_float4x4& b) {
    simd_float4x4 r;
    for (int c = 0; c < 4; ++c) r.columns[c] = simd_mul(a, b.columns[c]);
    return r;
}
inline simd_float4x4 simd_inverse(const simd_float4x4& m) {
    double a[16], inv[16];
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) a[4 * c + k] = m.columns[c][k];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    const double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    simd_float4x4 r;
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r.columns[c][k] = (float)(inv[4 * c + k] / det);
    return r;
}

struct simd_quatf { simd_float4 vector; };       /* (imaginary xyz, real w) */
inline simd_quatf simd_quaternion(float angle, simd_float3 axis) {
    const float s = sinf(angle / 2), c = cosf(angle / 2);
    simd_quatf q; q.vector = simd_float4(s * axis.x, s * axis.y, s * axis.z, c);
    return q;
}
inline simd_quatf simd_mul(const simd_quatf& p, const simd_quatf& q) {
    const simd_float4 &a = p.vector, &b = q.vector;
    simd_quatf r;
    r.vector = simd_float4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
                           a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
    return r;
}
inline simd_float3 simd_act(const simd_quatf& q, const simd_float3& v) {      /* v + 2 w (u x v) + 2 u x (u x v) */
    const simd_float3 u(q.vector.x, q.vector.y, q.vector.z);
    const simd_float3 t = simd_cross(u, v) * 2.0f;
    return v + t * q.vector.w + simd_cross(u, t);
}

namespace simd {
inline simd_float3 normalize(const simd_float3& a) { return simd_normalize(a); }
inline simd_float3 cross(const simd_float3& a, const simd_float3& b) { return simd_cross(a, b); }
inline float length(const simd_float3& a) { return simd_length(a); }
inline float dot(const simd_float3& a, const simd_float3& b) { return simd_dot(a, b); }
inline simd_float4x4 inverse(const simd_float4x4& m) { return simd_inverse(m); }
}  // namespace simd
