// tracer_b200/csrc/kernels/intersect.cuh
//
// Device intersectors of the ray query, written from the reference's arithmetic order
// (RT_Metal/Metal: AABB.hh:73-209, Triangle.hh:31-85, Sphere.hh:19-78, Square.hh:60-113,
// Cube.hh:17-47, HitRecord.hh:26-29). All fp32 ops via strict_math.cuh.
#pragma once
#include <cfloat>
#include <math_constants.h>

#include "../host/layout.h"
#include "strict_math.cuh"

namespace trq {

#define TRQ_PI_F   3.14159265358979323846264338327950288f
#define TRQ_PI_2_F 1.57079632679489661923132169163975144f

struct RayCtx {
    f3 o, d, inv;      // inv = 1.0f / d, hoisted: same bits as recomputing it per box test (AABB.hh:75,94)
};

__device__ __forceinline__ RayCtx make_ray_ctx(float ox, float oy, float oz, float dx, float dy, float dz) {
    RayCtx r;
    r.o = make_f3(ox, oy, oz);
    r.d = make_f3(dx, dy, dz);
    r.inv = make_f3(fdiv(1.0f, dx), fdiv(1.0f, dy), fdiv(1.0f, dz));
    return r;
}

// AABB::hit / hit_t  (AABB.hh:73-112). `t` is written only on success, like the reference.
__device__ __forceinline__ bool box_hit_t(const f3& mini, const f3& maxi, const RayCtx& r,
                                          float range_x, float range_y, float& t) {
    f3 ts = mul3(sub3(mini, r.o), r.inv);
    f3 te = mul3(sub3(maxi, r.o), r.inv);
    float tmin = fmaxf(fmaxf(fminf(ts.x, te.x), fminf(ts.y, te.y)), fminf(ts.z, te.z));
    float tmax = fminf(fminf(fmaxf(ts.x, te.x), fmaxf(ts.y, te.y)), fmaxf(ts.z, te.z));
    tmin = fmaxf(tmin, range_x);
    tmax = fminf(tmax, range_y);
    if (tmax < tmin || tmax < 0.0f) return false;
    t = (tmin < 0.0f) ? tmax : tmin;
    return true;
}

// hit_t specialised for Scene::hit's range_t.x == FLT_MIN (Render.hh:143), same results with fewer
// instructions: after `tmin = max(tmin, FLT_MIN)` tmin is >= FLT_MIN > 0 for every input (fmaxf drops
// NaNs), so `tmax < 0` implies `tmax < tmin` and the `tmin < 0 ? tmax : tmin` select always takes tmin.
__device__ __forceinline__ bool box_entry(const f3& mini, const f3& maxi, const f3& o, const f3& inv, float range_y, float& t) {
    f3 ts = mul3(sub3(mini, o), inv);
    f3 te = mul3(sub3(maxi, o), inv);
    float tmin = fmaxf(fmaxf(fminf(ts.x, te.x), fminf(ts.y, te.y)), fminf(ts.z, te.z));
    float tmax = fminf(fminf(fmaxf(ts.x, te.x), fmaxf(ts.y, te.y)), fmaxf(ts.z, te.z));
    tmin = fmaxf(tmin, FLT_MIN);
    tmax = fminf(tmax, range_y);
    if (tmax < tmin) return false;
    t = tmin;
    return true;
}

__device__ __forceinline__ bool box_hit(const f3& mini, const f3& maxi, const RayCtx& r,
                                        float range_x, float range_y) {
    float t;
    return box_hit_t(mini, maxi, r, range_x, range_y, t);
}

// HitRecord::checkFace
__device__ __forceinline__ bool front_face(const f3& dir, const f3& gn) { return dot3(dir, gn) <= 0.0f; }

// Triangle::hit_test on (v0, e1 = v1 - v0, e2 = v2 - v0)  (Triangle.hh:43-71).
// Precomputing e1/e2 is bit-compatible: those subtractions are the first operations of the test.
__device__ __forceinline__ bool tri_hit(const f3& v0, const f3& e1, const f3& e2, const RayCtx& r,
                                        float range_x, float range_y, float& t, float& u, float& v) {
    // Same tests in the same order as the reference, evaluated without early exits: in a warp the lanes that
    // fail early wait for the slowest lane anyway, and every rejection below is a pure comparison (no side
    // effects), so and-ing them gives exactly the reference's decision, including its NaN behaviour.
    f3 pvec = cross3(r.d, e2);
    float det = dot3(e1, pvec);
    bool ok = !(fabsf(det) < FLT_EPSILON);                 // Triangle.hh:55 (CULLING undefined)
    float invDet = fdiv(1.0f, det);
    f3 tvec = sub3(r.o, v0);
    float uu = fmul(dot3(tvec, pvec), invDet);
    ok = ok && !(uu < 0.0f || uu > 1.0f);                  // :62
    f3 qvec = cross3(tvec, e1);
    float vv = fmul(dot3(r.d, qvec), invDet);
    ok = ok && !(vv < 0.0f || fadd(uu, vv) > 1.0f);        // :66
    float tt = fmul(dot3(e2, qvec), invDet);
    ok = ok && !(tt > range_y || tt < range_x);            // :71  (t == range.y passes)
    if (ok) { t = tt; u = uu; v = vv; }
    return ok;
}

// Sphere::hit_test  (Sphere.hh:35-75); strict interval.
__device__ __forceinline__ bool sphere_hit(const f3& c, float radius, const RayCtx& r,
                                           float range_x, float range_y, float& t) {
    f3 oc = sub3(r.o, c);
    float a = dot3(r.d, r.d);
    float half_b = dot3(oc, r.d);
    float cc = fsub(dot3(oc, oc), fmul(radius, radius));
    float disc = fsub(fmul(half_b, half_b), fmul(a, cc));
    if (disc <= 0.0f) return false;
    float root = fsqrt(disc);
    float temp = fdiv(fsub(-half_b, root), a);
    if (!(temp < range_y && temp > range_x)) {
        temp = fdiv(fadd(-half_b, root), a);
        if (!(temp < range_y && temp > range_x)) return false;
    }
    t = temp;
    return true;
}

// Fields of HitRecord the query writes; filled on demand (final hit / expand).
struct Surface {
    f3 p, gn, sn;
    float uvx, uvy;
    uint32_t front, material;
};

__device__ __forceinline__ void finish_surface(const RayCtx& r, Surface& s) {
    bool f = front_face(r.d, s.gn);
    s.front = f ? 1u : 0u;
    s.sn = f ? s.gn : neg3(s.gn);
}

// Triangle.hh:73-82 given barycentrics
__device__ __forceinline__ void tri_surface(const RefVertex* __restrict__ triList, const uint32_t* __restrict__ idxList,
                                            uint32_t pIndex, float u, float v, const RayCtx& r, Surface& s) {
    const RefVertex* A = &triList[idxList[3 * pIndex]];
    const RefVertex* B = &triList[idxList[3 * pIndex + 1]];
    const RefVertex* C = &triList[idxList[3 * pIndex + 2]];
    float w = fsub(fsub(1.0f, u), v);
    s.p = add3(add3(scale3(ld3(B->v), u), scale3(ld3(C->v), v)), scale3(ld3(A->v), w));
    s.gn = add3(add3(scale3(ld3(B->n), u), scale3(ld3(C->n), v)), scale3(ld3(A->n), w));
    s.uvx = fadd(fadd(fmul(u, B->uv[0]), fmul(v, C->uv[0])), fmul(w, A->uv[0]));
    s.uvy = fadd(fadd(fmul(u, B->uv[1]), fmul(v, C->uv[1])), fmul(w, A->uv[1]));
    finish_surface(r, s);
    s.material = 19;
}

// Sphere.hh:51-56 given t
__device__ __forceinline__ void sphere_surface(const RefSphere* __restrict__ sp, float t, const RayCtx& r, Surface& s) {
    f3 c = ld3(sp->center);
    s.p = add3(r.o, scale3(r.d, t));
    s.gn = divs3(sub3(s.p, c), sp->radius);
    finish_surface(r, s);
    float phi = atan2f(s.gn.z, s.gn.x);
    float theta = asinf(s.gn.y);
    s.uvx = fsub(1.0f, fdiv(fadd(phi, TRQ_PI_F), fmul(2.0f, TRQ_PI_F)));
    s.uvy = fdiv(fadd(theta, TRQ_PI_2_F), TRQ_PI_F);
    s.material = sp->material;
}

// Square::hit_test  (Square.hh:82-111)
__device__ __forceinline__ bool square_hit(const RefSquare* __restrict__ sq, const RayCtx& r,
                                        float range_x, float range_y, float& t, Surface* s) {
    unsigned ai = sq->axis_i, aj = sq->axis_j, ak = sq->axis_k;
    float tt = fdiv(fsub(sq->value_k, get3(r.o, ak)), get3(r.d, ak));
    if (isinf(tt) || isnan(tt)) return false;
    if (tt < range_x || tt > range_y) return false;
    float a = fadd(get3(r.o, ai), fmul(tt, get3(r.d, ai)));
    if (a < sq->range_i[0] || a > sq->range_i[1]) return false;
    float b = fadd(get3(r.o, aj), fmul(tt, get3(r.d, aj)));
    if (b < sq->range_j[0] || b > sq->range_j[1]) return false;
    t = tt;
    if (s) {
        s->uvx = fdiv(fsub(a, sq->range_i[0]), fsub(sq->range_i[1], sq->range_i[0]));
        s->uvy = fdiv(fsub(b, sq->range_j[0]), fsub(sq->range_j[1], sq->range_j[0]));
        f3 gn = make_f3(0.0f, 0.0f, 0.0f);
        set3(gn, ak, 1.0f);
        s->gn = gn;
        finish_surface(r, *s);
        s->gn = s->sn;                                       // Square.hh:101
        f3 p = make_f3(0.0f, 0.0f, 0.0f);
        set3(p, ak, sq->value_k); set3(p, ai, a); set3(p, aj, b);
        s->p = p;
        s->material = sq->material;
    }
    return true;
}

// float4x4 * float4, columns summed left to right (oracle/shim/metal_stdlib)
__device__ __forceinline__ f3 m4_mul3(const float* __restrict__ m, const f3& v, float w) {
    f3 out;
    out.x = fadd(fadd(fadd(fmul(m[0], v.x), fmul(m[4], v.y)), fmul(m[8], v.z)), fmul(m[12], w));
    out.y = fadd(fadd(fadd(fmul(m[1], v.x), fmul(m[5], v.y)), fmul(m[9], v.z)), fmul(m[13], w));
    out.z = fadd(fadd(fadd(fmul(m[2], v.x), fmul(m[6], v.y)), fmul(m[10], v.z)), fmul(m[14], w));
    return out;
}

// AABB::hit(ray, range, record)  (AABB.hh:114-209), local-space box of a Cube
__device__ __forceinline__ bool box_hit_record(const RefAABB& b, const f3& o, const f3& d, float range_y,
                                               float& t_out, f3& p_out, f3& gn_out, float& uvx, float& uvy) {
    const float g3 = (3 * FLT_EPSILON * 0.5f) / (1 - 3 * FLT_EPSILON * 0.5f);    // gamma(3)  Math.hh:51-55
    const float widen = 1 + 2 * g3;
    float tmin = -FLT_MAX, tmax = range_y;
    unsigned axis = 0;
    f3 mini = make_f3(b.mini[0], b.mini[1], b.mini[2]), maxi = make_f3(b.maxi[0], b.maxi[1], b.maxi[2]);
    f3 ddd = sub3(o, mini), bbb = sub3(o, maxi);
    bool inside = (ddd.x > 0.0f && ddd.y > 0.0f && ddd.z > 0.0f) && (bbb.x < 0.0f && bbb.y < 0.0f && bbb.z < 0.0f);
    f3 gn = make_f3(0.0f, 0.0f, 0.0f), hp, p;
    if (inside) {
        for (unsigned i = 0; i < 3; ++i) {
            float lo = fdiv(fsub(get3(mini, i), get3(o, i)), get3(d, i));
            float hi = fdiv(fsub(get3(maxi, i), get3(o, i)), get3(d, i));
            float ts = fminf(hi, lo), te = fmaxf(hi, lo);
            te = fmul(te, widen);
            tmin = fmaxf(ts, tmin);
            if (te < tmax) { tmax = te; axis = i; }
            if (tmax < tmin || tmax < 0.0f) return false;
        }
        t_out = tmax;
        set3(gn, axis, get3(d, axis) > 0.0f ? 1.0f : -1.0f);
        hp = add3(o, scale3(d, tmax));
        p = hp;
        set3(p, axis, get3(d, axis) > 0.0f ? get3(maxi, axis) : get3(mini, axis));
    } else {
        for (unsigned i = 0; i < 3; ++i) {
            float lo = fdiv(fsub(get3(mini, i), get3(o, i)), get3(d, i));
            float hi = fdiv(fsub(get3(maxi, i), get3(o, i)), get3(d, i));
            float ts = fminf(hi, lo), te = fmaxf(hi, lo);
            te = fmul(te, widen);
            tmax = fminf(te, tmax);
            if (ts > tmin) { tmin = ts; axis = i; }
            if (tmax < tmin || tmax < 0.0f) return false;
        }
        t_out = tmin;
        set3(gn, axis, get3(d, axis) > 0.0f ? -1.0f : 1.0f);
        hp = add3(o, scale3(d, tmin));
        p = hp;
        set3(p, axis, get3(d, axis) > 0.0f ? get3(mini, axis) : get3(maxi, axis));
    }
    p_out = p; gn_out = gn;
    uvx = get3(hp, (1 + axis) % 3);
    uvy = get3(hp, (2 + axis) % 3);
    return true;
}

// Cube::hit_test  (Cube.hh:17-47)
__device__ __forceinline__ bool cube_hit(const RefCube* __restrict__ cb, const RayCtx& r,
                                      float range_x, float range_y, float& t, Surface* s) {
    (void)range_x;
    f3 lo = m4_mul3(cb->inverse, r.o, 1.0f);
    f3 ld = normalize3(m4_mul3(cb->inverse, r.d, 0.0f));     // Ray ctor normalises  Cube.hh:23
    float lt, uvx, uvy; f3 lp, lgn;
    if (!box_hit_record(cb->box, lo, ld, range_y, lt, lp, lgn, uvx, uvy)) return false;
    f3 wp = m4_mul3(cb->model, lp, 1.0f);
    f3 dv = sub3(r.o, wp);
    float tt = fsqrt(dot3(dv, dv));                           // distance(ray.origin, p)
    if (tt >= range_y) return false;
    t = tt;
    if (s) {
        s->gn = normalize3(m4_mul3(cb->normal, lgn, 0.0f));
        finish_surface(r, *s);
        s->p = wp;
        s->uvx = uvx; s->uvy = uvy;
        s->material = cb->material;
    }
    return true;
}

// The HitRecord fields of a cube hit the query has already found (trq_hit: t, u, v). Cube::hit_test's local box test starts
// its exit search at the CALLER's range.y (AABB.hh:116-145), which trq_hit does not carry, so the test cannot simply be run
// again. It does not have to be: re-running it with range.y = FLT_MAX gives the original outcome unless the ray started inside
// the local box and the original range.y was <= every exit distance (possible when the model matrix scales DOWN: local
// distances are longer than world distances). In that case the original kept axisPick = 0 and local t = range.y, so its
// local hit point was (the x plane the ray leaves through, uv.x, uv.y) and its normal +-x: everything follows from (u, v).
// The two cases are told apart by (t, u, v), which the re-run reproduces bit for bit in the first case.
__device__ __forceinline__ void cube_surface(const RefCube* __restrict__ cb, const RayCtx& r, float ht, float hu, float hv, Surface& s) {
    float t = 0.0f;
    Surface a;
    if (cube_hit(cb, r, FLT_MIN, FLT_MAX, t, &a) && __float_as_uint(t) == __float_as_uint(ht) &&
        __float_as_uint(a.uvx) == __float_as_uint(hu) && __float_as_uint(a.uvy) == __float_as_uint(hv)) { s = a; return; }
    const f3 ld = normalize3(m4_mul3(cb->inverse, r.d, 0.0f));                       // Cube.hh:20-23
    const bool pos = ld.x > 0.0f;
    const f3 lp = make_f3(pos ? cb->box.maxi[0] : cb->box.mini[0], hu, hv);          // AABB.hh:147-157 with axisPick = 0
    const f3 lgn = make_f3(pos ? 1.0f : -1.0f, 0.0f, 0.0f);
    s.gn = normalize3(m4_mul3(cb->normal, lgn, 0.0f));                               // Cube.hh:41-42
    finish_surface(r, s);
    s.p = m4_mul3(cb->model, lp, 1.0f);                                              // Cube.hh:26-27
    s.uvx = hu; s.uvy = hv;
    s.material = cb->material;
}

}  // namespace trq
