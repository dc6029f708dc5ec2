// tracer_b200/csrc/host/bvh_build.cpp
//
// Host-side scene preparation of the reference, restated: BVH::buildNode, BVH::make and
// BVH::buildTree (RT_Metal/Metal/BVH.hh:35-314, AABB helpers AABB.hh:18-49,213-253) and the
// per-triangle leaf creation of AAPLRenderer.mm:575-589. It produces the exact 64-byte node
// array the ray query consumes: root at 0, leaves at 1..N in creation order, interior nodes
// after, parent/left/right as final array indices.
//
// The reference runs the top three levels of BVH::make on GCD queues and appends nodes under
// a mutex, so ITS node numbering is racy. This restatement is the sequential variant (the one
// the reference hints at with depth=999, BVH.hh:261): interior nodes are numbered in post-order
// of completion. Because a subtree over k leaves always creates k-1 interior nodes, every
// node's final index is known before its subtree is built, so the top levels are built on
// worker threads here WITHOUT changing the numbering.
//
// Parity: the reference builder itself is compiled from its source as test infrastructure
// (oracle/ref_builder.cpp: blocks -> lambdas on the way into g++, libdispatch and Apple simd as
// small stand-ins, dispatch_async run in place = this sequential order); this file reproduces its
// node arrays byte for byte (tests/test_builder.py, tests/golden/bvh_golden.npz).

#include "../../../include/tracer_rq.h"
#include "error.h"
#include "layout.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <future>
#include <exception>
#include <vector>

namespace {

using trq::RefAABB;
using trq::RefBVH;
using trq::RefVertex;

struct V3 { float x, y, z; };

inline float comp(const V3& v, unsigned i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

struct Box {
    V3 mini{FLT_MAX, FLT_MAX, FLT_MAX};     // AABB.hh:8-9
    V3 maxi{-FLT_MAX, -FLT_MAX, -FLT_MAX};
    V3 diagonal() const { return {maxi.x - mini.x, maxi.y - mini.y, maxi.z - mini.z}; }   // :18-20
    float area() const {                                                                   // :27-30
        V3 d = diagonal();
        return 2 * (d.x * d.y + d.x * d.z + d.y * d.z);
    }
    unsigned maximumExtent() const {                                                       // :42-49
        V3 d = diagonal();
        if (d.x > d.y && d.x > d.z) return 0;
        return d.y > d.z ? 1 : 2;
    }
    void grow(const V3& p) {                                                               // :241-253
        mini = {fminf(mini.x, p.x), fminf(mini.y, p.y), fminf(mini.z, p.z)};
        maxi = {fmaxf(maxi.x, p.x), fmaxf(maxi.y, p.y), fmaxf(maxi.z, p.z)};
    }
    void grow(const Box& b) {                                                              // :227-239
        mini = {fminf(mini.x, b.mini.x), fminf(mini.y, b.mini.y), fminf(mini.z, b.mini.z)};
        maxi = {fmaxf(maxi.x, b.maxi.x), fmaxf(maxi.y, b.maxi.y), fmaxf(maxi.z, b.maxi.z)};
    }
};

inline Box box_of(const RefBVH& n) {
    Box b;
    b.mini = {n.bBOX.mini[0], n.bBOX.mini[1], n.bBOX.mini[2]};
    b.maxi = {n.bBOX.maxi[0], n.bBOX.maxi[1], n.bBOX.maxi[2]};
    return b;
}

// AABB::centroid  AABB.hh:22-25 :  mini + (maxi - mini) / 2
inline V3 centroid_of(const RefBVH& n) {
    const float* lo = n.bBOX.mini; const float* hi = n.bBOX.maxi;
    return {lo[0] + (hi[0] - lo[0]) / 2, lo[1] + (hi[1] - lo[1]) / 2, lo[2] + (hi[2] - lo[2]) / 2};
}

struct Builder {
    RefBVH* nodes;          // pre-shift indexing: leaves [0,N), interior [N, 2N-1)
    uint32_t* idx;          // idx_list
    const V3* cen;          // centroid per leaf (same bits as recomputing it on every use)
    uint32_t nLeaves;

    static constexpr unsigned nBuckets = 10;                                               // BVH.hh:91

    // bucket of leaf `leaf` along `dim` inside centroid box `cbox`   BVH.hh:96-100,143-147
    static inline unsigned bucket(const Box& cbox, const V3& c, unsigned dim) {
        float d = comp(cbox.diagonal(), dim);
        float rel = (comp(c, dim) - comp(cbox.mini, dim)) / d;                             // AABB::relative :37-40
        float scaled = nBuckets * rel;
        unsigned b = (scaled != scaled) ? 0u : (unsigned)scaled;     // NaN (0/0: all centroids equal) -> 0
        return std::min(b, nBuckets - 1);
    }

    // BVH::make  BVH.hh:35-244.  Returns the (pre-shift) index of the subtree root.
    // `base` = number of interior nodes created before this subtree in sequential post-order.
    uint32_t make(uint32_t start, uint32_t end, uint32_t depth, uint32_t base) {
        const uint32_t span = end - start;
        if (span == 1) return idx[start];                                                  // :52-54

        unsigned dim = 0;
        uint32_t left, right;

        if (span == 2) {                                                                   // :60-77
            const uint32_t ia = idx[start], ib = idx[start + 1];
            Box cbox; cbox.grow(cen[ia]); cbox.grow(cen[ib]);                              // AABB::make(a, b)
            dim = cbox.maximumExtent();
            if (comp(cen[ia], dim) < comp(cen[ib], dim)) { left = ia; right = ib; }
            else                                         { left = ib; right = ia; }
        } else {
            Box cbox;                                                                      // :81-87
            for (uint32_t i = start; i < end; ++i) cbox.grow(cen[idx[i]]);
            dim = cbox.maximumExtent();                                                    // :89

            struct { unsigned count = 0; Box bbox; } buckets[nBuckets];                    // :92
            for (uint32_t i = start; i < end; ++i) {                                       // :94-108
                const uint32_t leaf = idx[i];
                unsigned b = bucket(cbox, cen[leaf], dim);
                buckets[b].bbox.grow(box_of(nodes[leaf]));
                buckets[b].count++;
            }

            float cost[nBuckets - 1];                                                      // :110-130
            const float cboxArea = cbox.area();          // NB: area of the CENTROID box, as the reference does
            for (unsigned i = 0; i < nBuckets - 1; ++i) {
                Box b0, b1; int count0 = 0, count1 = 0;
                for (unsigned j = 0; j <= i; ++j) { b0.grow(buckets[j].bbox); count0 += (int)buckets[j].count; }
                for (unsigned j = i + 1; j < nBuckets; ++j) { b1.grow(buckets[j].bbox); count1 += (int)buckets[j].count; }
                cost[i] = 1 + (count0 * b0.area() + count1 * b1.area()) / cboxArea;
            }
            float minCost = cost[0];                                                       // :132-139
            unsigned minCostSplitBucket = 0;
            for (unsigned i = 1; i < nBuckets - 1; ++i)
                if (cost[i] < minCost) { minCost = cost[i]; minCostSplitBucket = i; }

            auto tester = [&](uint32_t i) { return bucket(cbox, cen[idx[i]], dim) <= minCostSplitBucket; };  // :141-150

            uint32_t mid;                                                                  // :152-168
            {
                uint32_t first = start, last = end;
                bool done = false;
                while (!done && first != last) {
                    while (tester(first)) { ++first; if (first == last) { done = true; break; } }
                    if (done) break;
                    do { --last; if (first == last) { done = true; break; } } while (!tester(last));
                    if (done) break;
                    std::swap(idx[first], idx[last]);
                    ++first;
                }
                mid = first;
            }

            if (mid <= start || mid >= end) {                                              // :187-195
                std::sort(idx + start, idx + end, [&](uint32_t a, uint32_t b) {
                    return comp(cen[a], dim) < comp(cen[b], dim);                          // box_compare :30-33
                });
                mid = start + span / 2;
            }

            const uint32_t leftBase = base, rightBase = base + (mid - start - 1);
            if (depth <= 3 && span > 65536) {                                              // parallel top levels (:197-219)
                auto fut = std::async(std::launch::async, [&] { return make(start, mid, depth + 1, leftBase); });
                right = make(mid, end, depth + 1, rightBase);
                left = fut.get();
            } else {
                left = make(start, mid, depth + 1, leftBase);                              // :199-200
                right = make(mid, end, depth + 1, rightBase);
            }
        }

        RefBVH nb;                                                                         // :222-231
        std::memset(&nb, 0, sizeof nb);
        nb.axis = dim;
        nb.left = left + 1;
        nb.right = right + 1;
        nb.pType = TRQ_BVH;
        Box u = box_of(nodes[left]); u.grow(box_of(nodes[right]));
        nb.bBOX.mini[0] = u.mini.x; nb.bBOX.mini[1] = u.mini.y; nb.bBOX.mini[2] = u.mini.z;
        nb.bBOX.maxi[0] = u.maxi.x; nb.bBOX.maxi[1] = u.maxi.y; nb.bBOX.maxi[2] = u.maxi.z;

        const uint32_t self = nLeaves + base + span - 2;         // = bvh_list.size() at emplace time, sequentially
        const uint32_t parent = self + 1;                                                  // :235
        nodes[self] = nb;                                                                  // :236
        nodes[left].parent = parent;                                                       // :240-241
        nodes[right].parent = parent;
        return parent - 1;                                                                 // :243
    }
};

}  // namespace

extern "C" {

int trq_bvh_build_node(const float box_min[3], const float box_max[3], const float model[16],
                       int32_t pType, uint32_t pIndex, void* node_out) {
    if (!box_min || !box_max || !node_out) return trq::fail(TRQ_ERR_INVALID, "trq_bvh_build_node: NULL argument");
    static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const float* m = model ? model : ident;
    const float* ele[2] = {box_min, box_max};                                              // BVH.hh:277
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};    // :279-280
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {                                                  // :282-299
                const float x = ele[i][0], y = ele[j][1], z = ele[k][2];
                for (int c = 0; c < 3; ++c) {
                    // simd_mul(model, (x,y,z,1)): column-major, columns summed left to right
                    const float t = ((m[0 + c] * x + m[4 + c] * y) + m[8 + c] * z) + m[12 + c] * 1.0f;
                    lo[c] = fminf(lo[c], t);
                    hi[c] = fmaxf(hi[c], t);
                }
            }
    RefBVH nb;                                                                             // :301-309
    std::memset(&nb, 0, sizeof nb);
    nb.pType = pType;
    nb.pIndex = pIndex;
    for (int c = 0; c < 3; ++c) { nb.bBOX.mini[c] = lo[c]; nb.bBOX.maxi[c] = hi[c]; }
    std::memcpy(node_out, &nb, sizeof nb);
    return TRQ_OK;
}

int trq_bvh_build_nodes_triangles(const void* triList, const uint32_t* idxList, uint32_t nTri,
                                  uint32_t pIndexBase, void* nodes_out) {
    if ((!triList || !idxList || !nodes_out) && nTri) return trq::fail(TRQ_ERR_INVALID, "trq_bvh_build_nodes_triangles: NULL argument");
    const RefVertex* tv = (const RefVertex*)triList;
    RefBVH* out = (RefBVH*)nodes_out;
    for (uint32_t i = 0; i < nTri; ++i) {
        const float* a = tv[idxList[3 * i]].v;
        const float* b = tv[idxList[3 * i + 1]].v;
        const float* c = tv[idxList[3 * i + 2]].v;
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {                                                      // AAPLRenderer.mm:575-586
            hi[k] = std::max({a[k], b[k], c[k]});
            lo[k] = std::min({a[k], b[k], c[k]});
        }
        int rc = trq_bvh_build_node(lo, hi, nullptr, TRQ_TRIANGLE, pIndexBase + i, &out[i]);   // :588-589
        if (rc != TRQ_OK) return rc;
    }
    return TRQ_OK;
}

int trq_bvh_build_tree(void* bvhList, uint32_t nLeaves, uint32_t* nNodeOut, uint32_t* maxDepthOut) {
    if (!bvhList || nLeaves == 0) return trq::fail(TRQ_ERR_INVALID, "trq_bvh_build_tree: empty leaf list");
    if (nLeaves > 0x7fffffffu) return trq::fail(TRQ_ERR_INVALID, "trq_bvh_build_tree: too many leaves");
    RefBVH* nodes = (RefBVH*)bvhList;
    const uint32_t nNode = 2 * nLeaves - 1;

    // no C++ exception may cross the C ABI: std::vector (bad_alloc) and std::async (system_error when no thread can be
    // started) both throw
    try {
    std::vector<uint32_t> idx(nLeaves);                                                    // BVH.hh:248-253
    std::vector<V3> cen(nLeaves);
    for (uint32_t i = 0; i < nLeaves; ++i) { idx[i] = i; cen[i] = centroid_of(nodes[i]); }

    Builder b{nodes, idx.data(), cen.data(), nLeaves};
    b.make(0, nLeaves, 0, 0);                                                              // :260
    } catch (const std::exception& e) {
        return trq::fail(TRQ_ERR_NOMEM, "trq_bvh_build_tree: %s", e.what());
    }

    // :263-268  move the root (last node) to the front; everything else shifts by +1, which the
    // +1s stored in left/right/parent anticipated.
    RefBVH root = nodes[nNode - 1];
    root.parent = 0;
    std::memmove(nodes + 1, nodes, sizeof(RefBVH) * (size_t)(nNode - 1));
    nodes[0] = root;
    nodes[root.left].parent = 0;
    nodes[root.right].parent = 0;

    // deepest interior level (root = 0): the reference's trail has 32 bits (Render.hh:140,172)
    uint32_t maxDepth = 0;
    if (nLeaves > 1) try {
        std::vector<uint32_t> depth(nNode, 0);
        // parents always have a larger pre-shift index than their children, i.e. interior nodes
        // appear after their descendants except the root at 0: walk from the end to the front.
        depth[0] = 0;
        for (uint32_t i = nNode - 1; i > nLeaves; --i) {
            // processed in decreasing index => parent before child
            if (nodes[i].pType == TRQ_BVH) {
                depth[i] = depth[nodes[i].parent] + 1;
                maxDepth = std::max(maxDepth, depth[i]);
            }
        }
    } catch (const std::exception& e) {
        return trq::fail(TRQ_ERR_NOMEM, "trq_bvh_build_tree: %s", e.what());
    }
    if (nNodeOut) *nNodeOut = nNode;
    if (maxDepthOut) *maxDepthOut = maxDepth;
    if (maxDepth > 31)
        return trq::fail(TRQ_ERR_DEPTH, "trq_bvh_build_tree: interior depth %u exceeds the 32-bit trail", maxDepth);
    return TRQ_OK;
}

}  // extern "C"
