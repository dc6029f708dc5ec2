/* oracle/nextweek_bvh.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C restatement of RT_Nextweek's CPU ray query, the baseline BASELINE.json's configs[0] names
 * ("RT_Nextweek random-spheres final scene, bvh_node closest-hit, 1280x720 primary rays, host CPU"):
 *   class BVH            RT_Nextweek/Tracer/BVH.swift:3-73      (build :9-51, hitTest :53-68)
 *   AABB.hit             RT_Nextweek/Tracer/AABB.swift:27-40
 *   surroundingBox       RT_Nextweek/Tracer/AABB.swift:43-48
 *   Sphere.hitTest       RT_Nextweek/Tracer/Sphere.swift:15-42, boundingBox :44-46
 *   HittableList.hitTest RT_Nextweek/Tracer/Hittable.swift:26-39
 * Swift is not installed here, so this restatement is NOT pinned against an execution of the reference
 * ("parity unpinned" for this row); it is used as a timed CPU baseline and as a sanity
 * check of hit sphere ids, never as the parity oracle of the CUDA path.
 *
 * Semantics kept from the Swift code: the split axis is random per node (arc4random there, a seeded PCG32
 * here); nodes of one or two primitives ignore the sort; the query tests BOTH children with the caller's
 * (t_min, t_max) and returns the nearer record (ties -> right); AABB.hit re-declares tmin/tmax per axis, so
 * the interval is not accumulated across axes; directions are not normalised.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float c[3]; float r; } nw_sphere;
typedef struct { float mn[3], mx[3]; int32_t left, right; } nw_node;   /* child >= 0: node index; < 0: ~sphere index */
typedef struct { float ox, oy, oz, tmax, dx, dy, dz; uint32_t flags; } nw_ray;   /* same 32 B as trq_ray */

typedef struct {
    nw_sphere* spheres; uint32_t n_spheres;
    nw_node* nodes; uint32_t n_nodes, cap_nodes;
    uint64_t rng_state, rng_inc;
} nw_scene;

static uint32_t pcg32(nw_scene* s) {
    uint64_t old = s->rng_state;
    s->rng_state = old * 6364136223846793005ULL + s->rng_inc;
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((-rot) & 31));
}
static float random_float(nw_scene* s) { return (float)pcg32(s) / (float)0xFFFFFFFFu; }   /* Random.swift:3-6 */

static void sphere_box(const nw_sphere* sp, float mn[3], float mx[3]) {                   /* Sphere.swift:44-46 */
    for (int k = 0; k < 3; ++k) { mn[k] = sp->c[k] - sp->r; mx[k] = sp->c[k] + sp->r; }
}

typedef struct { uint32_t id; float key; uint32_t order; } sort_item;
static int cmp_item(const void* a, const void* b) {
    const sort_item* x = (const sort_item*)a; const sort_item* y = (const sort_item*)b;
    if (x->key < y->key) return -1;
    if (y->key < x->key) return 1;
    return (x->order > y->order) - (x->order < y->order);                                 /* stable, like Swift's sort */
}

/* BVH.init  BVH.swift:9-51.  ids: sphere indices of this node's list (in list order). */
static int32_t build(nw_scene* s, const uint32_t* ids, uint32_t n) {
    int axis = (int)(3 * random_float(s));                                                /* :11 */
    if (axis > 2) axis = 2;
    sort_item* items = (sort_item*)malloc(sizeof(sort_item) * n);
    for (uint32_t i = 0; i < n; ++i) {
        float mn[3], mx[3];
        sphere_box(&s->spheres[ids[i]], mn, mx);
        items[i].id = ids[i]; items[i].key = mn[axis]; items[i].order = i;                /* :13-21 */
    }
    qsort(items, n, sizeof(sort_item), cmp_item);
    uint32_t me = s->n_nodes++;
    int32_t left, right;
    if (n == 1) { left = right = ~(int32_t)ids[0]; }                                      /* :35-37 */
    else if (n == 2) { left = ~(int32_t)ids[0]; right = ~(int32_t)ids[1]; }               /* :38-40 (unsorted list) */
    else {
        uint32_t mid = n / 2;                                                             /* :42 */
        uint32_t* sub = (uint32_t*)malloc(sizeof(uint32_t) * n);
        for (uint32_t i = 0; i < n; ++i) sub[i] = items[i].id;
        left = build(s, sub, mid);                                                        /* :43 */
        right = build(s, sub + mid, n - mid);                                             /* :44 */
        free(sub);
    }
    free(items);
    float lmn[3], lmx[3], rmn[3], rmx[3];
    if (left < 0) sphere_box(&s->spheres[~left], lmn, lmx); else { memcpy(lmn, s->nodes[left].mn, 12); memcpy(lmx, s->nodes[left].mx, 12); }
    if (right < 0) sphere_box(&s->spheres[~right], rmn, rmx); else { memcpy(rmn, s->nodes[right].mn, 12); memcpy(rmx, s->nodes[right].mx, 12); }
    nw_node* nd = &s->nodes[me];
    for (int k = 0; k < 3; ++k) { nd->mn[k] = fminf(lmn[k], rmn[k]); nd->mx[k] = fmaxf(lmx[k], rmx[k]); }   /* :47-50, AABB.swift:43-48 */
    nd->left = left; nd->right = right;
    return (int32_t)me;
}

nw_scene* nw_build(const float* centers_radii, uint32_t n, uint64_t seed, uint64_t seq) {
    nw_scene* s = (nw_scene*)calloc(1, sizeof(nw_scene));
    s->spheres = (nw_sphere*)malloc(sizeof(nw_sphere) * n);
    memcpy(s->spheres, centers_radii, sizeof(nw_sphere) * n);
    s->n_spheres = n;
    s->cap_nodes = 2 * n + 2;
    s->nodes = (nw_node*)calloc(s->cap_nodes, sizeof(nw_node));
    s->rng_state = 0; s->rng_inc = (seq << 1u) | 1u; pcg32(s); s->rng_state += seed; pcg32(s);
    uint32_t* ids = (uint32_t*)malloc(sizeof(uint32_t) * n);
    for (uint32_t i = 0; i < n; ++i) ids[i] = i;
    build(s, ids, n);                                                                     /* root = node 0 */
    free(ids);
    return s;
}

void nw_free(nw_scene* s) { if (s) { free(s->spheres); free(s->nodes); free(s); } }
uint32_t nw_node_count(const nw_scene* s) { return s->n_nodes; }

/* AABB.hit  AABB.swift:27-40 */
static int box_hit(const nw_node* b, const float o[3], const float d[3], float tmin, float tmax) {
    for (int i = 0; i < 3; ++i) {
        float minB = (b->mn[i] - o[i]) / d[i];
        float maxB = (b->mx[i] - o[i]) / d[i];
        float t0 = d[i] < 0.0f ? maxB : minB, t1 = d[i] < 0.0f ? minB : maxB;
        float lo = (tmin >= t0) ? tmin : t0;          /* Swift.max(t0, tmin) = tmin >= t0 ? tmin : t0 */
        float hi = (tmax < t1) ? tmax : t1;           /* Swift.min(t1, tmax) = tmax <  t1 ? tmax : t1 */
        if (hi <= lo) return 0;
    }
    return 1;
}

/* Sphere.hitTest  Sphere.swift:15-42 */
static int sphere_hit(const nw_sphere* sp, const float o[3], const float d[3], float t_min, float t_max, float* t) {
    float oc[3] = {o[0] - sp->c[0], o[1] - sp->c[1], o[2] - sp->c[2]};
    float a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    float b = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
    float c = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - sp->r * sp->r;
    float disc = b * b - a * c;
    if (disc <= 0) return 0;
    float tmp = (-b - sqrtf(disc)) / a;
    if (tmp < t_max && tmp > t_min) { *t = tmp; return 1; }
    tmp = (-b + sqrtf(disc)) / a;
    if (tmp < t_max && tmp > t_min) { *t = tmp; return 1; }
    return 0;
}

/* BVH.hitTest  BVH.swift:53-68: both children, always, with the caller's interval */
static int node_hit(const nw_scene* s, int32_t ref, const float o[3], const float d[3], float t_min, float t_max,
                    float* t, uint32_t* id) {
    if (ref < 0) {
        if (!sphere_hit(&s->spheres[~ref], o, d, t_min, t_max, t)) return 0;
        *id = (uint32_t)~ref;
        return 1;
    }
    const nw_node* n = &s->nodes[ref];
    if (!box_hit(n, o, d, t_min, t_max)) return 0;
    float tl, tr; uint32_t il, ir;
    int hl = node_hit(s, n->left, o, d, t_min, t_max, &tl, &il);
    int hr = node_hit(s, n->right, o, d, t_min, t_max, &tr, &ir);
    if (!hl && !hr) return 0;
    if (!hl) { *t = tr; *id = ir; return 1; }
    if (!hr) { *t = tl; *id = il; return 1; }
    if (tl < tr) { *t = tl; *id = il; } else { *t = tr; *id = ir; }
    return 1;
}

typedef struct { const nw_scene* s; const nw_ray* rays; uint64_t lo, hi; float t_min; uint32_t* hit_id; float* hit_t; } nw_job;

static void* job_main(void* arg) {
    nw_job* j = (nw_job*)arg;
    for (uint64_t i = j->lo; i < j->hi; ++i) {
        const float o[3] = {j->rays[i].ox, j->rays[i].oy, j->rays[i].oz};
        const float d[3] = {j->rays[i].dx, j->rays[i].dy, j->rays[i].dz};
        float t = 0; uint32_t id = 0xffffffffu;
        /* HittableList.hitTest over [treeBVH]  Hittable.swift:26-39 */
        int h = node_hit(j->s, 0, o, d, j->t_min, j->rays[i].tmax, &t, &id);
        j->hit_id[i] = h ? id : 0xffffffffu;
        j->hit_t[i] = h ? t : 0.0f;
    }
    return NULL;
}

/* world.hitTest(ray, 0.001, +max) for n rays (Render.swift:311-340), statically split over nthreads */
void nw_trace(const nw_scene* s, const nw_ray* rays, uint64_t n, float t_min, int nthreads, uint32_t* hit_id, float* hit_t) {
    if (nthreads < 1) nthreads = 1;
    if (n < 1024) nthreads = 1;
    nw_job* jobs = (nw_job*)calloc((size_t)nthreads, sizeof(nw_job));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int k = 0; k < nthreads; ++k) {
        jobs[k].s = s; jobs[k].rays = rays; jobs[k].t_min = t_min; jobs[k].hit_id = hit_id; jobs[k].hit_t = hit_t;
        jobs[k].lo = n * (uint64_t)k / (uint64_t)nthreads; jobs[k].hi = n * (uint64_t)(k + 1) / (uint64_t)nthreads;
        if (nthreads == 1) job_main(&jobs[k]); else pthread_create(&th[k], NULL, job_main, &jobs[k]);
    }
    if (nthreads > 1) for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
    free(jobs); free(th);
}
