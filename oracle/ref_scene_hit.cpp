// oracle/ref_scene_hit.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the reference's OWN ray-query path as host C++: this translation unit
// `#include`s /root/reference/RT_Metal/Metal/Render.hh UNMODIFIED (through the
// metal_stdlib shim in oracle/shim/) and exports thin extern "C" entry points around
//   Scene::hit            RT_Metal/Metal/Render.hh:135-252
//   AABB::hit / hit_t     RT_Metal/Metal/AABB.hh:73-112
//   Triangle::hit_test    RT_Metal/Metal/Triangle.hh:31-85
//   Sphere::hit_test      RT_Metal/Metal/Sphere.hh:33-78
//   Square::hit_test      RT_Metal/Metal/Square.hh:60-113
//   Cube::hit_test        RT_Metal/Metal/Cube.hh:17-47
//   offset_ray            RT_Metal/Metal/Math.hh:62-74
//   Ray(o, d) ctor        RT_Metal/Metal/Ray.hh:21-23
//   CoordinateSystem / CosineSampleHemisphere   RT_Metal/Metal/Sampling.hh:18-34,125-129
// No reference source is copied into this repository: the headers are read where they
// lie. The build recipe is oracle/Makefile (target `ref`), output oracle/_ref/libtracer_ref.so.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the resulting library.

// Standard headers first: the Metal address-space keywords below are #defined to nothing /
// const and would break <thread> & friends if those were included afterwards.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <thread>
#include <vector>
#include <sys/types.h>

#include "shim/metal_stdlib"

// Blank out the shading-side headers Render.hh:5-22 would pull in (not on the hot path). Random.hh stays: it only
// declares, and its pcg32_t {state, inc} (Random.hh:6-12) is what toRNG / exRNG (Render.hh:96-120) are written against.
#define Light_h
#define Spectrum_h
#define Medium_h
#define Material_h

#define __METAL_VERSION__ 1
#define constant const
#define thread
#define device
#define threadgroup

#include "Render.hh"   // -I/root/reference/RT_Metal/Metal   (verbatim reference source)
#include "Random.metal" // pcg32_srandom_r / pcg32_random_r / randomF: the definitions behind Random.hh's declarations

#undef constant
#undef thread
#undef device
#undef threadgroup

extern "C" {

// 64-byte dump of the reference HitRecord fields the ray query writes (HitRecord.hh:9-30).
struct ref_record {
    uint32_t hit;        // return value of Scene::hit / hit_test
    float    t;
    float    p[3];
    float    gn[3];
    float    sn[3];
    float    uv[2];
    uint32_t front;      // HitRecord::f
    uint32_t material;
    uint32_t pad;
};

// Mirrors `struct Primitive` (Render.hh:122-130): six pointers to reference-layout arrays.
struct ref_prims {
    const void* sphereList;
    const void* squareList;
    const void* cubeList;
    const void* triList;
    const uint32_t* idxList;
    const void* bvhList;
};

// 32-byte ray as used by the C-ABI: origin.xyz, tmax (= test_t), direction.xyz, flags.
struct ref_ray { float ox, oy, oz, tmax, dx, dy, dz; uint32_t flags; };

static inline void dump(const HitRecord& h, bool hit, ref_record* o) {
    o->hit = hit ? 1u : 0u;
    o->t = h.t;
    o->p[0] = h.p.x;  o->p[1] = h.p.y;  o->p[2] = h.p.z;
    o->gn[0] = h.gn.x; o->gn[1] = h.gn.y; o->gn[2] = h.gn.z;
    o->sn[0] = h.sn.x; o->sn[1] = h.sn.y; o->sn[2] = h.sn.z;
    o->uv[0] = h.uv.x; o->uv[1] = h.uv.y;
    o->front = h.f ? 1u : 0u;
    o->material = h.material;
    o->pad = 0;
}

// Scene::hit takes an already-constructed Ray; it never normalises. Build one without
// going through the normalising ctor so the direction bits are exactly the caller's.
static inline Ray make_ray(const float* o, const float* d) {
    Ray r;
    r.origin = float3(o[0], o[1], o[2]);
    r.direction = float3(d[0], d[1], d[2]);
    return r;
}

static inline Primitive to_primitive(const ref_prims* p) {
    Primitive q;
    q.sphereList = (const Sphere*)p->sphereList;
    q.squareList = (const Square*)p->squareList;
    q.cubeList = (const Cube*)p->cubeList;
    q.triList = (const TriangleVertex*)p->triList;
    q.idxList = p->idxList;
    q.bvhList = (const BVH*)p->bvhList;
    return q;
}

void ref_sizes(uint32_t* out) {
    out[0] = sizeof(BVH);            out[1] = sizeof(AABB);
    out[2] = sizeof(Sphere);         out[3] = sizeof(Square);
    out[4] = sizeof(Cube);           out[5] = sizeof(TriangleVertex);
    out[6] = sizeof(Ray);            out[7] = sizeof(HitRecord);
    out[8] = offsetof(BVH, pType);   out[9] = offsetof(BVH, pIndex);
    out[10] = offsetof(BVH, bBOX);   out[11] = offsetof(Sphere, center);
    out[12] = offsetof(Sphere, material);
    out[13] = offsetof(Square, value_k);
    out[14] = offsetof(Square, model_matrix);
    out[15] = offsetof(Square, material);
    out[16] = offsetof(Cube, inverse_matrix);
    out[17] = offsetof(Cube, box);
    out[18] = offsetof(Cube, material);
    out[19] = offsetof(Square, range_i);
    out[20] = offsetof(Square, range_j);
    out[21] = offsetof(Square, axis_k);
    out[22] = offsetof(Sphere, boundingBOX);
    out[23] = offsetof(Square, boundingBOX);
}

// One call of the reference's Scene::hit per ray, statically split over `nthreads`.
void ref_scene_hit(const ref_prims* prims, const ref_ray* rays, uint64_t n, int any,
                   int nthreads, ref_record* out) {
    Primitive prim = to_primitive(prims);
    auto work = [&](uint64_t lo, uint64_t hi) {
        Scene scene{prim};
        for (uint64_t i = lo; i < hi; ++i) {
            Ray ray = make_ray(&rays[i].ox, &rays[i].dx);
            HitRecord rec;
            std::memset((void*)&rec, 0, sizeof(rec));
            bool h = scene.hit(ray, rec, rays[i].tmax, any != 0);
            dump(rec, h, &out[i]);
        }
    };
    if (nthreads <= 1 || n < 1024) { work(0, n); return; }
    std::vector<std::thread> pool;
    for (int k = 0; k < nthreads; ++k) {
        uint64_t lo = n * (uint64_t)k / (uint64_t)nthreads, hi = n * (uint64_t)(k + 1) / (uint64_t)nthreads;
        pool.emplace_back(work, lo, hi);
    }
    for (auto& t : pool) t.join();
}

// Timing variant for the CPU baseline: only (hit, t) are kept, 8 bytes per ray.
void ref_scene_hit_lite(const ref_prims* prims, const ref_ray* rays, uint64_t n, int any,
                        int nthreads, uint32_t* hit_out, float* t_out) {
    Primitive prim = to_primitive(prims);
    auto work = [&](uint64_t lo, uint64_t hi) {
        Scene scene{prim};
        for (uint64_t i = lo; i < hi; ++i) {
            Ray ray = make_ray(&rays[i].ox, &rays[i].dx);
            HitRecord rec;
            rec.t = 0;
            bool h = scene.hit(ray, rec, rays[i].tmax, any != 0);
            hit_out[i] = h ? 1u : 0u;
            t_out[i] = h ? rec.t : 0.0f;
        }
    };
    if (nthreads <= 1 || n < 1024) { work(0, n); return; }
    std::vector<std::thread> pool;
    for (int k = 0; k < nthreads; ++k) {
        uint64_t lo = n * (uint64_t)k / (uint64_t)nthreads, hi = n * (uint64_t)(k + 1) / (uint64_t)nthreads;
        pool.emplace_back(work, lo, hi);
    }
    for (auto& t : pool) t.join();
}

// ---- leaf / box intersectors, one call each --------------------------------------------
// box: 8 floats in AABB layout {mini.xyz, pad, maxi.xyz, pad}; range: {t_min, t_max}.
int ref_aabb_hit(const float* box, const float* o, const float* d, const float* range) {
    const AABB* b = (const AABB*)box;
    Ray ray = make_ray(o, d);
    float2 r(range[0], range[1]);
    return b->hit(ray, r) ? 1 : 0;
}
int ref_aabb_hit_t(const float* box, const float* o, const float* d, const float* range, float* t) {
    const AABB* b = (const AABB*)box;
    Ray ray = make_ray(o, d);
    float2 r(range[0], range[1]);
    return b->hit_t(ray, r, *t) ? 1 : 0;
}
int ref_triangle_hit(const void* triList, const uint32_t* abc, const float* o, const float* d,
                     float* range, ref_record* out) {
    uint3 idx(abc[0], abc[1], abc[2]);
    Triangle tri((const TriangleVertex*)triList, idx);
    Ray ray = make_ray(o, d);
    float2 r(range[0], range[1]);
    HitRecord rec; std::memset((void*)&rec, 0, sizeof(rec));
    bool h = tri.hit_test(ray, r, rec);
    range[0] = r.x; range[1] = r.y;
    dump(rec, h, out);
    return h ? 1 : 0;
}
int ref_sphere_hit(const void* sphere, const float* o, const float* d, float* range, ref_record* out) {
    const Sphere* s = (const Sphere*)sphere;
    Ray ray = make_ray(o, d);
    float2 r(range[0], range[1]);
    HitRecord rec; std::memset((void*)&rec, 0, sizeof(rec));
    bool h = s->hit_test(ray, r, rec);
    range[0] = r.x; range[1] = r.y;
    dump(rec, h, out);
    return h ? 1 : 0;
}
int ref_square_hit(const void* square, const float* o, const float* d, float* range, ref_record* out) {
    const Square* s = (const Square*)square;
    Ray ray = make_ray(o, d);
    float2 r(range[0], range[1]);
    HitRecord rec; std::memset((void*)&rec, 0, sizeof(rec));
    bool h = s->hit_test(ray, r, rec);
    range[0] = r.x; range[1] = r.y;
    dump(rec, h, out);
    return h ? 1 : 0;
}
int ref_cube_hit(const void* cube, const float* o, const float* d, float* range, ref_record* out) {
    const Cube* c = (const Cube*)cube;
    Ray ray = make_ray(o, d);
    float2 r(range[0], range[1]);
    HitRecord rec; std::memset((void*)&rec, 0, sizeof(rec));
    bool h = c->hit_test(ray, r, rec);
    range[0] = r.x; range[1] = r.y;
    dump(rec, h, out);
    return h ? 1 : 0;
}

// ---- ray construction helpers (used by the harness restatements' tests) ----------------
void ref_ray_ctor(const float* o, const float* d, float* out_o, float* out_d) {
    Ray ray(float3(o[0], o[1], o[2]), float3(d[0], d[1], d[2]));   // normalising ctor
    out_o[0] = ray.origin.x; out_o[1] = ray.origin.y; out_o[2] = ray.origin.z;
    out_d[0] = ray.direction.x; out_d[1] = ray.direction.y; out_d[2] = ray.direction.z;
}
void ref_offset_ray(const float* p, const float* n, float* out) {
    float3 r = offset_ray(float3(p[0], p[1], p[2]), float3(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void ref_coordinate_system(const float* a, float* b, float* c) {
    float3 aa(a[0], a[1], a[2]), bb, cc;
    CoordinateSystem(aa, bb, cc);
    b[0] = bb.x; b[1] = bb.y; b[2] = bb.z;
    c[0] = cc.x; c[1] = cc.y; c[2] = cc.z;
}
void ref_cosine_sample_hemisphere(const float* u, float* out) {
    float2 uu(u[0], u[1]);
    float3 r = CosineSampleHemisphere(uu);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float ref_next_float_up(float v) { return NextFloatUp(v); }
float ref_next_float_down(float v) { return NextFloatDown(v); }

// castRay (Camera.hh:59-69) with a camera given as MakeCamera's outputs {lookFrom, u, v, vertical, horizontal, cornerLowLeft}
// and lenRadius; the sampler is seeded (seed, seq) because castRay draws sampleUnitInDisk() even when lenRadius is 0.
void ref_cast_ray(const float* cam18, float lenRadius, float s, float t, uint64_t seed, uint64_t seq, float* origin, float* direction) {
    Camera c;
    c.lookFrom = float3(cam18[0], cam18[1], cam18[2]);
    c.u = float3(cam18[3], cam18[4], cam18[5]);
    c.v = float3(cam18[6], cam18[7], cam18[8]);
    c.vertical = float3(cam18[9], cam18[10], cam18[11]);
    c.horizontal = float3(cam18[12], cam18[13], cam18[14]);
    c.cornerLowLeft = float3(cam18[15], cam18[16], cam18[17]);
    c.lenRadius = lenRadius;
    pcg32_t r;
    pcg32_srandom_r(&r, seed, seq);
    RandomSampler rs { &r };
    Ray ray = castRay(&c, s, t, &rs);
    origin[0] = ray.origin.x; origin[1] = ray.origin.y; origin[2] = ray.origin.z;
    direction[0] = ray.direction.x; direction[1] = ray.direction.y; direction[2] = ray.direction.z;
}

// Square::sample (Square.hh:40-58): the light sample of the NEE shadow ray (Render.metal:313-323)
void ref_square_sample(const void* square, const float* u2, const float* pos3, float* p_out, float* n_out) {
    const Square* sq = reinterpret_cast<const Square*>(square);
    LightSampleRecord lsr;
    float2 uu(u2[0], u2[1]);
    float3 pos(pos3[0], pos3[1], pos3[2]);
    sq->sample(uu, pos, lsr);
    p_out[0] = lsr.p.x; p_out[1] = lsr.p.y; p_out[2] = lsr.p.z;
    n_out[0] = lsr.n.x; n_out[1] = lsr.n.y; n_out[2] = lsr.n.z;
}

// PCG32 as the reference's kernels run it (Random.metal:3-26) and RandomSampler::sample2D (RandomSampler.hh:16-21)
void ref_pcg32_fill(uint64_t initstate, uint64_t initseq, uint32_t n, uint32_t* u32_out, float* f32_out) {
    pcg32_t a, b;
    pcg32_srandom_r(&a, initstate, initseq);
    pcg32_srandom_r(&b, initstate, initseq);
    for (uint32_t i = 0; i < n; ++i) { u32_out[i] = pcg32_random_r(&a); f32_out[i] = randomF(&b); }
}
void ref_sample2d(uint64_t initstate, uint64_t initseq, float* out2) {
    pcg32_t r;
    pcg32_srandom_r(&r, initstate, initseq);
    RandomSampler rs { &r };
    float2 uu = rs.sample2D();
    out2[0] = uu.x; out2[1] = uu.y;
}

// the per-pixel RNG state texture <-> pcg32_t (Render.hh:96-120); out / in = {inc, state} BY MEMBER NAME.
// toRNG brace-initialises `pcg32_t { rng_inc, rng_state }` positionally against a struct declared {state, inc}
// (Random.hh:6-12), so the texel words it calls "state" (r, g) land in .inc and (b, a) in .state, while exRNG writes
// .state to (r, g) and .inc to (b, a): the two halves of a texel trade places once per frame. That is the reference's
// behaviour and it is what these two exports expose.
void ref_to_rng(const uint32_t* rgba, uint64_t* inc_state) {
    vec<uint32_t, 4> c; c.r = rgba[0]; c.g = rgba[1]; c.b = rgba[2]; c.a = rgba[3];
    pcg32_t r = toRNG(c);
    inc_state[0] = r.inc; inc_state[1] = r.state;
}
void ref_ex_rng(const uint64_t* inc_state, uint32_t* rgba) {
    pcg32_t r;
    r.inc = inc_state[0]; r.state = inc_state[1];
    vec<uint32_t, 4> c = exRNG(r);
    rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a;
}

}  // extern "C"
