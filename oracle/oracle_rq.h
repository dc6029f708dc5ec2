/* oracle/oracle_rq.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, strict IEEE fp32, no FMA contraction) of the reference's
 * ray-query hot path, instrumented with the per-ray step counters SURVEY.md section 8(d)
 * uses for "algorithmic bytes per ray", and extended with the primitive id of the winning
 * leaf (the reference HitRecord carries none, HitRecord.hh:9-30).
 *
 * Parity pin: this restatement is checked bit-for-bit (hit flag, t, p, gn, sn, front,
 * material; uv bit-for-bit too because both sides call the same libm) against the
 * reference's own source compiled verbatim (oracle/ref_scene_hit.cpp -> oracle/_ref/) and
 * against golden vectors generated from that build (tests/golden/, script
 * tests/golden/make_golden.py). The reference has no tests or golden vectors of its own
 * (SURVEY.md section 4), so "outputs of the reference itself run here" is the pin.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use anything in oracle/. The product path (tracer_b200/, include/) never does.
 *
 * Byte layouts are the reference's (Metal / Apple simd rules: float3 is 16 B):
 *   BVH            64 B  RT_Metal/Metal/BVH.hh:15-22
 *   AABB           32 B  RT_Metal/Metal/AABB.hh:7-9
 *   TriangleVertex 32 B  RT_Metal/Metal/Triangle.hh:12-18
 *   Sphere        272 B  RT_Metal/Metal/Sphere.hh:6-15
 *   Square        272 B  RT_Metal/Metal/Square.hh:12-27
 *   Cube          240 B  RT_Metal/Metal/Cube.hh:6-13
 */
#ifndef ORACLE_RQ_H
#define ORACLE_RQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORQ_SPHERE = 0, ORQ_SQUARE = 1, ORQ_CUBE = 2, ORQ_TRIANGLE = 3, ORQ_BVH = 4, ORQ_UNKNOW = 5 };

typedef struct { float mini[3]; float pad0; float maxi[3]; float pad1; } orq_aabb;
typedef struct {
    uint32_t parent, left, right, axis;
    int32_t  pType;
    uint32_t pIndex;
    uint32_t pad[2];
    orq_aabb bBOX;
} orq_bvh;
typedef struct { float v[3]; float n[3]; float uv[2]; } orq_vertex;
typedef struct {
    float radius; float pad0[3];
    float center[3]; float pad1;
    float model[16], normal[16], inverse[16];     /* column-major */
    uint32_t material; uint32_t pad2[3];
    orq_aabb boundingBOX;
} orq_sphere;
typedef struct {
    uint8_t axis_i, axis_j; uint8_t pad0[6];
    float range_i[2];
    float range_j[2];
    uint8_t axis_k; uint8_t pad1[3];
    float value_k;
    float model[16], normal[16], inverse[16];
    uint32_t material; uint32_t pad2[3];
    orq_aabb boundingBOX;
} orq_square;
typedef struct {
    float model[16], normal[16], inverse[16];
    orq_aabb box;
    uint32_t material; uint32_t pad[3];
} orq_cube;

/* mirrors `struct Primitive` (Render.hh:122-130) */
typedef struct {
    const orq_sphere* sphereList;
    const orq_square* squareList;
    const orq_cube*   cubeList;
    const orq_vertex* triList;
    const uint32_t*   idxList;
    const orq_bvh*    bvhList;
} orq_prims;

/* 32-byte ray / hit, identical to trq_ray / trq_hit in include/tracer_rq.h */
typedef struct { float ox, oy, oz, tmax, dx, dy, dz; uint32_t flags; } orq_ray;
typedef struct {
    float    t;
    uint32_t pType, pIndex, leafNode;
    float    u, v;          /* triangle: barycentrics (Triangle.hh:61,65); others: record.uv */
    uint32_t material;
    uint32_t flags;         /* bit0 = hit, bit1 = front face (HitRecord::f) */
} orq_hit;

/* the HitRecord fields the query writes; same 64-byte layout as ref_record */
typedef struct {
    uint32_t hit; float t; float p[3]; float gn[3]; float sn[3]; float uv[2];
    uint32_t front; uint32_t material; uint32_t pad;
} orq_record;

/* per-ray work counters */
typedef struct {
    uint32_t n_fp;     /* from-parent interior steps (two child box tests each)   Render.hh:155-187 */
    uint32_t n_fc;     /* from-child steps                                         Render.hh:189-209 */
    uint32_t n_disp;   /* child dispatches (pType/pIndex read)                     Render.hh:211-213 */
    uint32_t n_tri, n_sph, n_sq, n_cube;   /* leaf tests by type */
    uint32_t n_tie;    /* accepted hits with t == previous range.y (order-dependent winner) */
    uint32_t n_quirk;  /* "selected the missed child" events    Render.hh:174 */
    uint32_t max_level;
} orq_counters;

/* totals over a batch, as uint64 in the same order as orq_counters (max_level = max) */
typedef struct { uint64_t v[10]; } orq_totals;

int  orq_aabb_hit(const orq_aabb* box, const float o[3], const float d[3], const float range[2]);
int  orq_aabb_hit_t(const orq_aabb* box, const float o[3], const float d[3], const float range[2], float* t);
int  orq_triangle_hit(const orq_vertex* triList, const uint32_t abc[3], const float o[3], const float d[3],
                      float range[2], orq_record* rec, float bary[2]);
int  orq_sphere_hit(const orq_sphere* s, const float o[3], const float d[3], float range[2], orq_record* rec);
int  orq_square_hit(const orq_square* s, const float o[3], const float d[3], float range[2], orq_record* rec);
int  orq_cube_hit(const orq_cube* c, const float o[3], const float d[3], float range[2], orq_record* rec);
void orq_offset_ray(const float p[3], const float n[3], float out[3]);
void orq_normalize(const float d[3], float out[3]);

/* Scene::hit for one ray. Returns the hit flag; fills hit / rec / cnt when non-NULL. */
int  orq_scene_hit(const orq_prims* prims, const orq_ray* ray, int any,
                   orq_hit* hit, orq_record* rec, orq_counters* cnt);

/* Batch over n rays, statically split over nthreads. hits required; recs, cnts, totals optional. */
void orq_trace(const orq_prims* prims, const orq_ray* rays, uint64_t n, int any, int nthreads,
               orq_hit* hits, orq_record* recs, orq_counters* cnts, orq_totals* totals);

/* algorithmic bytes from totals: 60*N_fp + 12*N_fc + 8*N_disp + 48*N_tri + 16*N_sph + 48*n_rays
 * (+ 32 per square test, + 160 per cube test: the fields those tests read) */
uint64_t orq_algorithmic_bytes(const orq_totals* totals, uint64_t n_rays);

#ifdef __cplusplus
}
#endif
#endif
