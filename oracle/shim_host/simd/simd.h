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
