// tracer_b200/csrc/kernels/trace_kernels.cuh -- the ray-query kernels (sm_100a).
//
//   trace_reflayout_kernel  1:1 transcription of Scene::hit (Render.hh:135-252) over the reference-
//                           layout buffers: parent links + 32-bit trail, one thread per ray.
//                           Correctness anchor and naive baseline.
//   trace_packed_kernel     the product kernel: persistent warps pull rays from a global queue with
//                           one warp-aggregated atomic per refill, traverse the packed 64-B fused
//                           nodes with 256-bit loads, and keep the far children of "both hit" nodes
//                           on a per-lane shared-memory stack. The stack replaces the reference's
//                           from-child steps (re-reading parent links, Render.hh:189-209); the visit
//                           ORDER per ray is exactly the reference's: near child first by hit_t's t,
//                           ties to the right child, the far child decided at first arrival and never
//                           re-tested, including the "select the missed child" quirk of Render.hh:174.
//                           FUSED variant (triangle-only scenes): writes the final trq_hit when a ray retires.
//   resolve_hits_kernel     compact result -> trq_hit (pType, pIndex, front, material, sphere uv) for scenes with
//                           sphere / square / cube leaves; GATHER variant also stores the record to every peer GPU
//   sort_*_kernel           optional ordering of the work queue (TRQ_SORT_RAYS)
//   expand_hits_kernel      trq_hit -> HitRecord fields (p, gn, sn, uv, f, material)
#pragma once
#include <cfloat>

#include "../../../include/tracer_rq.h"
#include "intersect.cuh"
#include "scene_dev.cuh"

namespace trq {

// Batch size of a launch: `n` from the host, or -- for trq_trace_indirect, where the count was produced on the
// device by a spawn kernel -- *nPtr clamped to the capacity n.
__device__ __forceinline__ uint64_t live_count(uint64_t n, const unsigned long long* nPtr) {
    if (nPtr == nullptr) return n;
    const unsigned long long m = *nPtr;
    return m < n ? (uint64_t)m : n;
}

// ---------------------------------------------------------------------------------------------
// loads
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// 256-bit read-only load (LDG.E.256 on sm_100): two float4 from a 32-byte aligned address.
__device__ __forceinline__ void ldg8(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// 256-bit store (STG.E.256 on sm_100): one request for a whole 32-byte record.
__device__ __forceinline__ void stg8(void* p, const float4& a, const float4& b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// aux: cube hits = front | material << 1; every other hit = bits of the ray direction's x, with y and z in dy / dz,
// so that resolve_hits_kernel can do a triangle's front-face test without reading the ray again.
__device__ __forceinline__ void store_compact(trq_hit* hits, uint64_t i, bool hit, float t, uint32_t leaf,
                                              float u, float v, uint32_t aux, float dy = 0.0f, float dz = 0.0f) {
    float4 a, b;
    a.x = hit ? t : 0.0f;
    a.y = __uint_as_float(hit ? leaf : 0u);
    a.z = hit ? u : 0.0f;
    a.w = hit ? v : 0.0f;
    b.x = __uint_as_float(hit ? aux : 0u);
    b.y = __uint_as_float(hit ? 1u : 0u);
    b.z = dy; b.w = dz;
    stg8(hits + i, a, b);
}

// Leaf tests that read reference-layout structs (rare leaf types: Square, Cube). Out of line and fed with
// scalars by value so that the hot loops keep the ray in registers (no local-memory RayCtx); the ray context is
// rebuilt here (1/d recomputed: same bits). out = {t, u, v, bits(aux)}. Returns true on accept.
__device__ __noinline__ bool leaf_square_cube(const RefBVH* __restrict__ bvh, const RefSquare* __restrict__ squares,
                                              const RefCube* __restrict__ cubes, uint32_t kind, uint32_t leafNode,
                                              float ox, float oy, float oz, float dx, float dy, float dz,
                                              float range_y, float4* out) {
    const RayCtx ray = make_ray_ctx(ox, oy, oz, dx, dy, dz);
    const uint32_t pIndex = bvh[leafNode].pIndex;
    float t = 0.0f;
    if (kind == REF_SQUARE) {
        if (!square_hit(&squares[pIndex], ray, FLT_MIN, range_y, t, nullptr)) return false;
        *out = make_float4(t, 0.0f, 0.0f, 0.0f);
        return true;
    }
    Surface s;
    if (!cube_hit(&cubes[pIndex], ray, FLT_MIN, range_y, t, &s)) return false;
    *out = make_float4(t, s.uvx, s.uvy, __uint_as_float(s.front | (s.material << 1)));
    return true;
}

// ---------------------------------------------------------------------------------------------
// v0: reference layout, reference control flow.
template <bool ANY>
__global__ void __launch_bounds__(128)
trace_reflayout_kernel(SceneDev S, const trq_ray* __restrict__ rays, trq_hit* __restrict__ hits, uint64_t n,
                       const unsigned long long* __restrict__ nPtr) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= live_count(n, nPtr)) return;
    const float4 r0 = ldg4(reinterpret_cast<const float4*>(rays + i));
    const float4 r1 = ldg4(reinterpret_cast<const float4*>(rays + i) + 1);
    const RayCtx ray = make_ray_ctx(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
    const float test_t = r0.w;
    const RefBVH* __restrict__ N = S.bvh;

    uint32_t the_index = 0, tested_index = 0xffffffffu;                   // Render.hh:137-138
    uint32_t stack_mark = 0, stack_level = 0;                             // :140-141
    float range_y = test_t;                                               // :143  range_t = (FLT_MIN, test_t)
    uint32_t best = 0xffffffffu, aux = 0; float bu = 0.0f, bv = 0.0f;
    bool done_any = false;

    if (box_hit(ld3(N[0].bBOX.mini), ld3(N[0].bBOX.maxi), ray, FLT_MIN, range_y)) {   // :145
        do {
            uint32_t sel;
            const uint4 links = *reinterpret_cast<const uint4*>(&N[the_index]);       // parent,left,right,axis :151-153
            const uint32_t p = links.x, l = links.y, r = links.z;
            if (tested_index != l && tested_index != r) {                 // :155
                float tl = range_y, tr = range_y;                         // :157
                const bool lt = box_hit_t(ld3(N[l].bBOX.mini), ld3(N[l].bBOX.maxi), ray, FLT_MIN, range_y, tl);   // :159
                const bool rt = box_hit_t(ld3(N[r].bBOX.mini), ld3(N[r].bBOX.maxi), ray, FLT_MIN, range_y, tr);   // :160
                if (!lt && !rt) { tested_index = the_index; the_index = p; stack_level -= 1; continue; }          // :162-169
                if (lt && rt) stack_mark |= 1u << (stack_level & 31u);    // :171-172
                sel = (tl < tr) ? l : r;                                  // :174
            } else {
                const uint32_t need = (stack_mark >> (stack_level & 31u)) & 1u;       // :191
                stack_mark &= ~(1u << (stack_level & 31u));               // :193
                if (need == 0) { tested_index = the_index; the_index = p; stack_level -= 1; continue; }           // :195-202
                sel = (tested_index == l) ? r : l;                        // :204-208
            }
            const int32_t pType = N[sel].pType;                           // :211-213
            const uint32_t pIndex = N[sel].pIndex;
            bool h = false; float t = 0.0f, u = 0.0f, v = 0.0f; uint32_t a = 0;
            if (pType == TRQ_BVH) { the_index = sel; stack_level += 1; continue; }    // :215-220
            if (pType == TRQ_TRIANGLE) {                                  // :230-240
                const f3 v0 = ld3(S.verts[S.idx[3 * pIndex]].v);
                const f3 v1 = ld3(S.verts[S.idx[3 * pIndex + 1]].v);
                const f3 v2 = ld3(S.verts[S.idx[3 * pIndex + 2]].v);
                h = tri_hit(v0, sub3(v1, v0), sub3(v2, v0), ray, FLT_MIN, range_y, t, u, v);
            } else if (pType == TRQ_SPHERE) {                             // :221-223
                h = sphere_hit(ld3(S.spheres[pIndex].center), S.spheres[pIndex].radius, ray, FLT_MIN, range_y, t);
            } else if (pType == TRQ_SQUARE || pType == TRQ_CUBE) {        // :224-229
                float4 o4;
                h = leaf_square_cube(S.bvh, S.squares, S.cubes, pType == TRQ_SQUARE ? REF_SQUARE : REF_CUBE, sel,
                                     ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, range_y, &o4);
                if (h) { t = o4.x; u = o4.y; v = o4.z; a = __float_as_uint(o4.w); }
            }
            if (h) {                                                      // aux word: cube front/material, else d.x (see store_compact)
                range_y = t; best = sel; bu = u; bv = v;
                aux = (pType == TRQ_SQUARE || pType == TRQ_CUBE) ? a : __float_as_uint(ray.d.x);
            }
            if (ANY && range_y < test_t) { done_any = true; break; }      // :244
            tested_index = sel;                                           // :246
        } while (tested_index != 0);                                      // :248
    }
    const bool hit = done_any || (range_y < test_t);                      // :251
    store_compact(hits, i, hit && best != 0xffffffffu, range_y, best, bu, bv, aux, ray.d.y, ray.d.z);
}

// ---------------------------------------------------------------------------------------------
// v1: packed layout, persistent warps, shared-memory far-child stack.
#ifndef TRQ_BLOCK
#define TRQ_BLOCK 256
#endif

#define TRQ_TRI_STRIDE 4u      // float4 per packed triangle: 64-byte records (48 used) so that a test is two requests, never three

struct TraceParams {
    const trq_ray* rays;
    trq_hit*       hits;
    uint64_t       n;
    unsigned long long* counter;   // global ray queue head
    uint32_t       stackDepth;     // entries per lane
    uint32_t       refillMin;      // refill when at least this many lanes of a warp are idle
    uint32_t       leafBatch;      // leave the interior phase when this many lanes wait at a leaf
    const uint32_t* order;         // optional queue order (TRQ_SORT_RAYS): queue slot -> ray index; NULL = identity
    const unsigned long long* nPtr;    // optional device-resident batch size (trq_trace_indirect); n is then the capacity
};

// ---------------------------------------------------------------------------------------------
// Optional ray ordering (TRQ_SORT_RAYS): a counting sort of ray INDICES by (Morton cell of the origin inside the
// scene box, direction octant), so that the rays one warp pulls from the queue start in the same region and head
// the same way. Rays are not moved; hits are still written at the ray's own index; the per-ray traversal is
// untouched, so results are identical with and without it.
#define TRQ_SORT_CELL_BITS 4u                                   // 16^3 cells (B200, C5: 4 bits 829, 5 bits 772, 6 bits 572 Mrays/s)
#define TRQ_SORT_BINS (1u << (3u * TRQ_SORT_CELL_BITS + 3u))    // x 8 octants = 32,768 bins

__device__ __forceinline__ uint32_t spread3(uint32_t v) {       // up to 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ uint32_t ray_sort_key(const SceneDev& S, const float4& r0, const float4& r1) {
    const float cells = (float)(1u << TRQ_SORT_CELL_BITS);
    uint32_t c[3];
    const float o[3] = {r0.x, r0.y, r0.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float ext = S.rootMax[k] - S.rootMin[k];
        float f = ext > 0.0f ? (o[k] - S.rootMin[k]) / ext * cells : 0.0f;
        f = fminf(fmaxf(f, 0.0f), cells - 1.0f);                // origins outside the scene box clamp to the border cells
        c[k] = (uint32_t)f;
    }
    const uint32_t morton = spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2);
    const uint32_t octant = (r1.x < 0.0f ? 1u : 0u) | (r1.y < 0.0f ? 2u : 0u) | (r1.z < 0.0f ? 4u : 0u);
    return (morton << 3) | octant;
}

// Both atomic passes are warp-aggregated (__match_any_sync): lanes with the same key elect one leader that
// issues a single atomicAdd for the group, so a batch whose rays all share one bin (e.g. primary rays from one
// eye point) costs n/32 same-address atomics, not n.
__global__ void __launch_bounds__(256)
sort_count_kernel(SceneDev S, const trq_ray* __restrict__ rays, uint64_t n, const unsigned long long* __restrict__ nPtr,
                  uint32_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < live_count(n, nPtr);
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const float4 r0 = ldg4(reinterpret_cast<const float4*>(rays + i));
    const float4 r1 = ldg4(reinterpret_cast<const float4*>(rays + i) + 1);
    const uint32_t key = ray_sort_key(S, r0, r1);
    keys[i] = key;
    const unsigned peers = __match_any_sync(vm, key);
    if ((threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[key], (uint32_t)__popc(peers));
}

// exclusive prefix sum of TRQ_SORT_BINS counters, one CTA of 1024 threads (256 bins per thread)
__global__ void __launch_bounds__(1024)
sort_scan_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t partial[1024];
    const uint32_t per = TRQ_SORT_BINS / 1024u;
    uint32_t* mine = hist + threadIdx.x * per;
    uint32_t sum = 0;
    for (uint32_t k = 0; k < per; ++k) sum += mine[k];
    partial[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024u; off <<= 1) {            // Hillis-Steele inclusive scan
        uint32_t v = threadIdx.x >= off ? partial[threadIdx.x - off] : 0u;
        __syncthreads();
        partial[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = partial[threadIdx.x] - sum;                  // exclusive base of this thread's bins
    for (uint32_t k = 0; k < per; ++k) { const uint32_t c = mine[k]; mine[k] = run; run += c; }
}

__global__ void __launch_bounds__(256)
sort_scatter_kernel(const uint32_t* __restrict__ keys, uint64_t n, const unsigned long long* __restrict__ nPtr,
                    uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < live_count(n, nPtr);
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const uint32_t key = keys[i];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned peers = __match_any_sync(vm, key);
    const unsigned leader = (unsigned)(__ffs(peers) - 1);
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = (uint32_t)i;     // lanes keep their index order inside a group
}

// Per-ray state that the interior loop never touches is parked in shared memory ([word][TRQ_BLOCK],
// lane-major => bank-conflict free) so that the hot loop fits in 48 registers (5 resident CTAs per SM).
enum : uint32_t { COLD_TEST_T = 0, COLD_BEST, COLD_U, COLD_V, COLD_AUX, COLD_RAY, COLD_DX, COLD_DY, COLD_DZ, COLD_WORDS };

#ifndef TRQ_INTERIOR_UNROLL
#define TRQ_INTERIOR_UNROLL 2
#endif
#ifndef TRQ_MIN_BLOCKS
#define TRQ_MIN_BLOCKS 5
#endif

#ifdef TRQ_STATS
__device__ unsigned long long g_stats[8];
#endif

// FUSED (triangle-only scenes): a retiring ray's final trq_hit is written by this kernel; otherwise a compact result is
// written and resolve_hits_kernel finishes it.
template <bool ANY, bool FUSED>
__global__ void __launch_bounds__(TRQ_BLOCK, TRQ_MIN_BLOCKS)
trace_packed_kernel(const SceneDev S, const TraceParams P) {
    extern __shared__ uint32_t smem_u32[];
    uint32_t* const stk = smem_u32 + threadIdx.x;                                   // [stackDepth][TRQ_BLOCK]
    uint32_t* const cold = smem_u32 + P.stackDepth * TRQ_BLOCK + threadIdx.x;        // [COLD_WORDS][TRQ_BLOCK]
    float* const coldf = reinterpret_cast<float*>(cold);
    const unsigned lane = threadIdx.x & 31u;

    const uint64_t N = live_count(P.n, P.nPtr);
    bool active = false, exhausted = false;
    f3 ro = make_f3(0.f, 0.f, 0.f), rinv = make_f3(0.f, 0.f, 0.f);
    float range_y = 0.0f;
    uint32_t cur = TRQ_REF_DONE_WORD, sp = 0;

    // A finished ray only retires its lane; its compact result (resolved into trq_hit by resolve_hits_kernel) is
    // written by flush() when the warp next refills, for all retired lanes at once: the six shared-memory reads and the
    // store are then issued once per refill instead of once per finishing ray (most rays finish alone in their warp step).
    bool pending = false;
    auto finish = [&]() { active = false; pending = true; };
    auto flush = [&]() {
        if (pending) {
            const uint32_t best = cold[COLD_BEST * TRQ_BLOCK];
            const bool hit = (range_y < coldf[COLD_TEST_T * TRQ_BLOCK]) && best != 0xffffffffu;   // Render.hh:251
            if (FUSED) {
                // triangle-only scene: `best` is the triangle slot; finish the record here (Triangle.hh:73-82: interpolated
                // normal, checkFace) and write the final trq_hit -- no resolve pass over the batch afterwards
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                if (hit) {
                    const float4* np = S.triN + (size_t)best * 4u;
                    float4 q0, q1;
                    ldg8(np, q0, q1);
                    const float4 q2 = ldg4(np + 2);
                    const float u = coldf[COLD_U * TRQ_BLOCK], v = coldf[COLD_V * TRQ_BLOCK];
                    const float w = fsub(fsub(1.0f, u), v);
                    const f3 gn = add3(add3(scale3(make_f3(q1.x, q1.y, q1.z), u), scale3(make_f3(q2.x, q2.y, q2.z), v)),
                                       scale3(make_f3(q0.x, q0.y, q0.z), w));
                    const f3 d = make_f3(__uint_as_float(cold[COLD_AUX * TRQ_BLOCK]), coldf[COLD_DY * TRQ_BLOCK], coldf[COLD_DZ * TRQ_BLOCK]);
                    const uint32_t front = front_face(d, gn) ? TRQ_HIT_FLAG_FRONT : 0u;
                    o0 = make_float4(range_y, __uint_as_float((uint32_t)TRQ_TRIANGLE), q1.w, q0.w);
                    o1 = make_float4(u, v, __uint_as_float(19u), __uint_as_float(TRQ_HIT_FLAG_HIT | front));
                }
                stg8(P.hits + cold[COLD_RAY * TRQ_BLOCK], o0, o1);
            } else {
                // triangle / sphere hits carry the ray direction (aux word = d.x) so that resolve_hits_kernel does not
                // have to read the ray again for the front-face test
                store_compact(P.hits, cold[COLD_RAY * TRQ_BLOCK], hit, range_y, best,
                              coldf[COLD_U * TRQ_BLOCK], coldf[COLD_V * TRQ_BLOCK], cold[COLD_AUX * TRQ_BLOCK],
                              coldf[COLD_DY * TRQ_BLOCK], coldf[COLD_DZ * TRQ_BLOCK]);
            }
            pending = false;
        }
    };
    auto pop = [&]() {
        if (sp == 0) cur = TRQ_REF_DONE_WORD; else { --sp; cur = stk[sp * TRQ_BLOCK]; }
    };

    for (;;) {
        // ---- refill idle lanes from the global queue: one atomic per warp ----
        const unsigned idleMask = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (idleMask == 0xffffffffu || __popc(idleMask) >= (int)P.refillMin)) {
            const int want = __popc(idleMask);
#ifdef TRQ_STATS
            if (lane == 0) { atomicAdd(&g_stats[4], 1ull); atomicAdd(&g_stats[5], (unsigned long long)want); }
#endif
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.counter, (unsigned long long)want);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + (unsigned long long)want >= N) exhausted = true;
            flush();
            if (!active) {
                const uint64_t slot = base + (uint64_t)__popc(idleMask & ((1u << lane) - 1u));
                if (slot < N) {
                    const uint64_t idx = P.order ? (uint64_t)__ldg(P.order + slot) : slot;
                    float4 r0, r1;
                    ldg8(reinterpret_cast<const float4*>(P.rays + idx), r0, r1);          // trq_ray is one 32-byte record
                    const RayCtx ray = make_ray_ctx(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
                    ro = ray.o; rinv = ray.inv;
                    range_y = r0.w;                                        // Render.hh:143  range_t = (FLT_MIN, test_t)
                    sp = 0;
                    const f3 rootMin = make_f3(S.rootMin[0], S.rootMin[1], S.rootMin[2]);
                    const f3 rootMax = make_f3(S.rootMax[0], S.rootMax[1], S.rootMax[2]);
                    if (box_hit(rootMin, rootMax, ray, FLT_MIN, range_y)) {            // :145
                        coldf[COLD_TEST_T * TRQ_BLOCK] = r0.w;
                        cold[COLD_BEST * TRQ_BLOCK] = 0xffffffffu;
                        coldf[COLD_U * TRQ_BLOCK] = 0.0f; coldf[COLD_V * TRQ_BLOCK] = 0.0f;
                        cold[COLD_AUX * TRQ_BLOCK] = 0u;
                        cold[COLD_RAY * TRQ_BLOCK] = (uint32_t)idx;
                        coldf[COLD_DX * TRQ_BLOCK] = r1.x; coldf[COLD_DY * TRQ_BLOCK] = r1.y; coldf[COLD_DZ * TRQ_BLOCK] = r1.z;
                        cur = S.rootRef; active = true;
                    } else {
                        store_compact(P.hits, idx, false, 0.0f, 0u, 0.0f, 0.0f, 0u);
                    }
                }
            }
        }
        const unsigned actMask = __ballot_sync(0xffffffffu, active);
        if (actMask == 0u) { if (exhausted) { flush(); break; } else continue; }

        // ---- traverse until enough lanes have retired to make a refill worthwhile ----
        // Two phases per round so that the (rarer) leaf code is not issued on every interior step:
        //   interior phase: lanes sitting on an interior node step; lanes that reached a leaf wait (a lane
        //     must test its leaf BEFORE it continues, or pruning/tie order would differ from the reference);
        //     the phase ends when no lane is on an interior node or enough lanes wait at leaves;
        //   leaf phase: every waiting lane tests its primitive and pops.
        const int keepGoing = exhausted ? 0 : (32 - (int)P.refillMin);
        do {
            for (;;) {
                const bool onInterior = active && TRQ_REF_KIND(cur) == REF_INTERIOR;
                const int nInt = __popc(__ballot_sync(0xffffffffu, onInterior));
                const int nWait = __popc(__ballot_sync(0xffffffffu, active && !onInterior));
                // leave the interior phase once the lanes waiting at a leaf are more than half of those still stepping
                // (B200 sweep, C3 / 1 M soup Mrays/s: never 4094 / 1141, nWait > nInt 4650 / 1154, 2 nWait > nInt 4727 / 1158,
                // 3x 4701 / 1157, 4x 4678 / 1158, 6x 4614 / 1155)
#ifndef TRQ_WAIT_MUL
#define TRQ_WAIT_MUL 2
#endif
                if (nInt == 0 || nWait >= (int)P.leafBatch || TRQ_WAIT_MUL * nWait > nInt) break;
#pragma unroll
                for (int rep = 0; rep < TRQ_INTERIOR_UNROLL; ++rep) {          // steps per vote round
#ifdef TRQ_STATS
                    { const unsigned m_ = __ballot_sync(0xffffffffu, active && TRQ_REF_KIND(cur) == REF_INTERIOR);
                      if (lane == 0 && m_) { atomicAdd(&g_stats[0], 1ull); atomicAdd(&g_stats[1], (unsigned long long)__popc(m_)); } }
#endif
                    if (active && TRQ_REF_KIND(cur) == REF_INTERIOR) {
                        const float4* np = S.nodes + (size_t)TRQ_REF_INDEX(cur) * 4u;
                        float4 q0, q1, q2, q3;
                        ldg8(np, q0, q1);
                        ldg8(np + 2, q2, q3);
                        float tl = range_y, tr = range_y;                     // :157
                        const bool lt = box_entry(make_f3(q0.x, q0.y, q0.z), make_f3(q1.x, q1.y, q1.z), ro, rinv, range_y, tl);   // :159
                        const bool rt = box_entry(make_f3(q2.x, q2.y, q2.z), make_f3(q3.x, q3.y, q3.z), ro, rinv, range_y, tr);   // :160
                        const uint32_t lref = __float_as_uint(q0.w), rref = __float_as_uint(q1.w);
                        if (!lt && !rt) {                                     // :162-169  pop
                            pop();
                        } else {
                            const bool selLeft = tl < tr;                     // :174 (ties -> right, quirk included)
                            if (lt && rt) { stk[sp * TRQ_BLOCK] = selLeft ? rref : lref; ++sp; }   // :171-172
                            cur = selLeft ? lref : rref;
                        }
                        if (cur == TRQ_REF_DONE_WORD) finish();
                    }
                }
            }
#ifdef TRQ_STATS
            { const unsigned m_ = __ballot_sync(0xffffffffu, active && TRQ_REF_KIND(cur) != REF_INTERIOR);
              if (lane == 0 && m_) { atomicAdd(&g_stats[2], 1ull); atomicAdd(&g_stats[3], (unsigned long long)__popc(m_)); } }
#endif
            if (active && TRQ_REF_KIND(cur) != REF_INTERIOR) {
                const uint32_t kind = TRQ_REF_KIND(cur);
                RayCtx ray;
                ray.o = ro; ray.inv = rinv;
                ray.d = make_f3(coldf[COLD_DX * TRQ_BLOCK], coldf[COLD_DY * TRQ_BLOCK], coldf[COLD_DZ * TRQ_BLOCK]);
                bool h = false; float t = 0.0f, u = 0.0f, v = 0.0f; uint32_t leaf = 0, a = 0;
                if (kind == REF_TRI) {
                    const float4* tp = S.tris + (size_t)TRQ_REF_INDEX(cur) * TRQ_TRI_STRIDE;
                    float4 t0, t1;
                    ldg8(tp, t0, t1);                                       // one 256-bit + one 128-bit request per test
                    const float4 t2 = ldg4(tp + 2);
                    leaf = __float_as_uint(t0.w);
                    h = tri_hit(make_f3(t0.x, t0.y, t0.z), make_f3(t1.x, t1.y, t1.z), make_f3(t2.x, t2.y, t2.z),
                                ray, FLT_MIN, range_y, t, u, v);
                } else if (kind == REF_SPHERE) {
                    const float4* spp = S.sph + (size_t)TRQ_REF_INDEX(cur) * 2u;
                    const float4 s0 = ldg4(spp), s1 = ldg4(spp + 1);
                    leaf = __float_as_uint(s1.x);
                    h = sphere_hit(make_f3(s0.x, s0.y, s0.z), s0.w, ray, FLT_MIN, range_y, t);
                } else if (kind == REF_SQUARE || kind == REF_CUBE) {
                    leaf = TRQ_REF_INDEX(cur);
                    float4 o4;
                    h = leaf_square_cube(S.bvh, S.squares, S.cubes, kind, leaf, ro.x, ro.y, ro.z, ray.d.x, ray.d.y, ray.d.z, range_y, &o4);
                    if (h) { t = o4.x; u = o4.y; v = o4.z; a = __float_as_uint(o4.w); }
                }
                if (h) {
                    range_y = t;
                    cold[COLD_BEST * TRQ_BLOCK] = FUSED ? TRQ_REF_INDEX(cur) : leaf;
                    coldf[COLD_U * TRQ_BLOCK] = u; coldf[COLD_V * TRQ_BLOCK] = v;
                    cold[COLD_AUX * TRQ_BLOCK] = (kind == REF_SQUARE || kind == REF_CUBE) ? a : __float_as_uint(ray.d.x);
                }
                if (ANY && range_y < coldf[COLD_TEST_T * TRQ_BLOCK]) cur = TRQ_REF_DONE_WORD;   // :244
                else pop();
                if (cur == TRQ_REF_DONE_WORD) finish();
            }
        } while (__popc(__ballot_sync(0xffffffffu, active)) > keepGoing);
    }
}

// ---------------------------------------------------------------------------------------------
// compact -> trq_hit, in place. One thread per ray, fully coalesced.
// GATHER: the resolved record is also stored into this rank's slot of every peer's buffer (NVLink peer memory, one
// coalesced 1-KB run per warp and peer); the last CTA to finish publishes count and step number to every peer with
// system-scope release stores, which the peers' gather_wait_kernel acquires.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <bool GATHER>
__global__ void __launch_bounds__(256)
resolve_hits_kernel(SceneDev S, const trq_ray* __restrict__ rays, trq_hit* hits, uint64_t n, const unsigned long long* __restrict__ nPtr,
                    unsigned long long* queueHead, const GatherDev G) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && queueHead) *queueHead = 0ull;        // the trace that used this queue head finished before this kernel started
    const uint64_t live = live_count(n, nPtr);
    if (i < live) {
        float4* io = reinterpret_cast<float4*>(hits + i);
        const float4 a = io[0], b = io[1];
        trq_hit out;
        out.t = 0.0f; out.pType = 0; out.pIndex = 0; out.leafNode = 0; out.u = 0.0f; out.v = 0.0f; out.material = 0; out.flags = 0;
        if (__float_as_uint(b.y) != 0u) {
            const uint32_t leaf = __float_as_uint(a.y);
            const int32_t pType = S.bvh[leaf].pType;
            const uint32_t pIndex = S.bvh[leaf].pIndex;
            RayCtx ray;
            if (pType == TRQ_TRIANGLE) {                          // only the direction matters (checkFace), and the trace kernel left it here
                ray = make_ray_ctx(0.0f, 0.0f, 0.0f, b.x, b.z, b.w);
            } else {
                const float4 r0 = ldg4(reinterpret_cast<const float4*>(rays + i));
                const float4 r1 = ldg4(reinterpret_cast<const float4*>(rays + i) + 1);
                ray = make_ray_ctx(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
            }
            out.t = a.x; out.pType = (uint32_t)pType; out.pIndex = pIndex; out.leafNode = leaf;
            Surface s; s.front = 0; s.material = 0; s.uvx = s.uvy = 0.0f;
            if (pType == TRQ_TRIANGLE) {
                tri_surface(S.verts, S.idx, pIndex, a.z, a.w, ray, s);
                out.u = a.z; out.v = a.w;
            } else if (pType == TRQ_SPHERE) {
                sphere_surface(&S.spheres[pIndex], a.x, ray, s);
                out.u = s.uvx; out.v = s.uvy;
            } else if (pType == TRQ_SQUARE) {
                float t;
                square_hit(&S.squares[pIndex], ray, a.x, a.x, t, &s);          // same t -> same a, b, uv
                out.u = s.uvx; out.v = s.uvy;
            } else if (pType == TRQ_CUBE) {
                const uint32_t aux = __float_as_uint(b.x);
                s.front = aux & 1u; s.material = aux >> 1;
                out.u = a.z; out.v = a.w;
            }
            out.material = s.material;
            out.flags = TRQ_HIT_FLAG_HIT | (s.front ? TRQ_HIT_FLAG_FRONT : 0u);
        }
        float4 o0, o1;
        o0.x = out.t; o0.y = __uint_as_float(out.pType); o0.z = __uint_as_float(out.pIndex); o0.w = __uint_as_float(out.leafNode);
        o1.x = out.u; o1.y = out.v; o1.z = __uint_as_float(out.material); o1.w = __uint_as_float(out.flags);
        io[0] = o0; io[1] = o1;
        if (GATHER) {
#pragma unroll 1
            for (uint32_t p = 0; p < G.nPeer; ++p) stg8(G.peerSlot[p] + i, o0, o1);
        }
    }
    if (GATHER) {
        __threadfence_system();                                   // this thread's peer stores before the CTA's arrival
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(G.blocksDone, 1u);
            if (prev == gridDim.x - 1) {                          // last CTA: every record of this rank is on its way
                *G.blocksDone = 0u;
                __threadfence_system();
                *G.ownCount = live;
                st_release_sys(G.ownFlag, G.step);
                for (uint32_t p = 0; p < G.nPeer; ++p) {
                    *G.peerCount[p] = live;
                    st_release_sys(G.peerFlag[p], G.step);
                }
            }
        }
    }
}

// Waits (on the stream) until every rank has published `step`; gives up after timeoutNs and raises *status.
__global__ void gather_wait_kernel(const unsigned long long* flags, uint32_t world, unsigned long long step,
                                   unsigned long long timeoutNs, volatile unsigned int* status) {
    const uint32_t p = threadIdx.x;
    if (p >= world) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(flags + p) < step) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeoutNs) { *status = 1u + p; break; }
        __nanosleep(256);
    }
}

// trq_hit -> HitRecord fields.
__global__ void __launch_bounds__(256)
expand_hits_kernel(SceneDev S, const trq_ray* __restrict__ rays, const trq_hit* __restrict__ hits,
                   trq_hit_record* __restrict__ recs, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const trq_hit h = hits[i];
    trq_hit_record out;
    out.hit = 0; out.t = 0.0f; out.front = 0; out.material = 0; out.pad = 0;
    out.p[0] = out.p[1] = out.p[2] = 0.0f; out.gn[0] = out.gn[1] = out.gn[2] = 0.0f;
    out.sn[0] = out.sn[1] = out.sn[2] = 0.0f; out.uv[0] = out.uv[1] = 0.0f;
    if (h.flags & TRQ_HIT_FLAG_HIT) {
        const float4 r0 = ldg4(reinterpret_cast<const float4*>(rays + i));
        const float4 r1 = ldg4(reinterpret_cast<const float4*>(rays + i) + 1);
        const RayCtx ray = make_ray_ctx(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z);
        Surface s; s.p = s.gn = s.sn = make_f3(0.f, 0.f, 0.f); s.uvx = s.uvy = 0.0f; s.front = 0; s.material = 0;
        float t;
        if (h.pType == TRQ_TRIANGLE)     tri_surface(S.verts, S.idx, h.pIndex, h.u, h.v, ray, s);
        else if (h.pType == TRQ_SPHERE)  sphere_surface(&S.spheres[h.pIndex], h.t, ray, s);
        else if (h.pType == TRQ_SQUARE)  square_hit(&S.squares[h.pIndex], ray, h.t, h.t, t, &s);
        else if (h.pType == TRQ_CUBE)    cube_hit(&S.cubes[h.pIndex], ray, FLT_MIN, FLT_MAX, t, &s);
        out.hit = 1; out.t = h.t;
        out.p[0] = s.p.x; out.p[1] = s.p.y; out.p[2] = s.p.z;
        out.gn[0] = s.gn.x; out.gn[1] = s.gn.y; out.gn[2] = s.gn.z;
        out.sn[0] = s.sn.x; out.sn[1] = s.sn.y; out.sn[2] = s.sn.z;
        out.uv[0] = s.uvx; out.uv[1] = s.uvy;
        out.front = s.front; out.material = s.material;
    }
    recs[i] = out;
}

// ---------------------------------------------------------------------------------------------
// derive the packed layout from the reference-layout buffers (one thread per BVH node)
__global__ void __launch_bounds__(256)
pack_scene_kernel(const RefBVH* __restrict__ bvh, const uint32_t* __restrict__ ref, uint32_t nNode,
                  const RefVertex* __restrict__ verts, const uint32_t* __restrict__ idx,
                  const RefSphere* __restrict__ spheres,
                  float4* __restrict__ nodes, float4* __restrict__ tris, float4* __restrict__ sph, float4* __restrict__ triN) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nNode) return;
    const uint32_t my = ref[i];
    const uint32_t kind = TRQ_REF_KIND(my), slot = TRQ_REF_INDEX(my);
    if (my == TRQ_REF_DONE_WORD) return;                      // unreachable node
    if (kind == REF_INTERIOR) {
        const uint32_t l = bvh[i].left, r = bvh[i].right;
        const RefAABB lb = bvh[l].bBOX, rb = bvh[r].bBOX;
        float4* out = nodes + (size_t)slot * 4u;
        out[0] = make_float4(lb.mini[0], lb.mini[1], lb.mini[2], __uint_as_float(ref[l]));
        out[1] = make_float4(lb.maxi[0], lb.maxi[1], lb.maxi[2], __uint_as_float(ref[r]));
        out[2] = make_float4(rb.mini[0], rb.mini[1], rb.mini[2], 0.0f);
        out[3] = make_float4(rb.maxi[0], rb.maxi[1], rb.maxi[2], 0.0f);
    } else if (kind == REF_TRI) {
        const uint32_t p = bvh[i].pIndex;
        const f3 v0 = ld3(verts[idx[3 * p]].v), v1 = ld3(verts[idx[3 * p + 1]].v), v2 = ld3(verts[idx[3 * p + 2]].v);
        const f3 e1 = sub3(v1, v0), e2 = sub3(v2, v0);       // Triangle.hh:43-44
        float4* out = tris + (size_t)slot * TRQ_TRI_STRIDE;
        out[0] = make_float4(v0.x, v0.y, v0.z, __uint_as_float(i));
        out[1] = make_float4(e1.x, e1.y, e1.z, 0.0f);
        out[2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
        // what a finished triangle hit needs beyond (t, u, v): the vertex normals for checkFace, and its ids
        const f3 n0 = ld3(verts[idx[3 * p]].n), n1 = ld3(verts[idx[3 * p + 1]].n), n2 = ld3(verts[idx[3 * p + 2]].n);
        float4* on = triN + (size_t)slot * 4u;
        on[0] = make_float4(n0.x, n0.y, n0.z, __uint_as_float(i));
        on[1] = make_float4(n1.x, n1.y, n1.z, __uint_as_float(p));
        on[2] = make_float4(n2.x, n2.y, n2.z, 0.0f);
        on[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    } else if (kind == REF_SPHERE) {
        const RefSphere* s = &spheres[bvh[i].pIndex];
        float4* out = sph + (size_t)slot * 2u;
        out[0] = make_float4(s->center[0], s->center[1], s->center[2], s->radius);
        out[1] = make_float4(__uint_as_float(i), 0.0f, 0.0f, 0.0f);
    }
}

}  // namespace trq
