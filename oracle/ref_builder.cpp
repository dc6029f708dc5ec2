// oracle/ref_builder.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's own host BVH builder (RT_Metal/Metal/BVH.hh:30-314: BVH::buildNode, BVH::make, BVH::buildTree, with
// AABB.hh's host helpers), compiled from the sources where they lie and exported through a C ABI, so that the
// restated builder (tracer_b200/csrc/host/bvh_build.cpp) and the GPU builder can be checked against it byte for byte.
//
// BVH.hh needs three things g++ does not have; none of them changes what the builder computes:
//   * clang blocks (`^{ ... }`, three of them, BVH.hh:152,207,213): the recipe in oracle/Makefile streams the file through
//     `sed 's/\^{/[\&]{/g'` into the compiler (this TU includes it as /dev/stdin) -- nothing of the reference is copied
//     into the repository or written to disk; `__block` is defined away;
//   * libdispatch: oracle/shim_host/dispatch/dispatch.h runs submitted work immediately, i.e. the reference's
//     sequential order (its commented variant, BVH.hh:261), which makes the node numbering deterministic;
//   * Apple <simd/simd.h>, <Metal/Metal.h>, <MetalKit/MetalKit.h> (Common.hh:20-22): oracle/shim_host/ (strict IEEE
//     fp32 vector structs; the Metal headers are empty). `__auto_type` (Common.hh:17-18) is C-only in GCC: -D__auto_type=auto.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include <dispatch/dispatch.h>

#include "/dev/stdin"        // RT_Metal/Metal/BVH.hh with `^{` -> `[&]{`  (see oracle/Makefile, target ref)

static_assert(sizeof(BVH) == 64, "struct BVH layout (BVH.hh:15-22)");
static_assert(sizeof(AABB) == 32, "struct AABB layout (AABB.hh:7-9)");

extern "C" {

int refb_sizes(uint32_t* bvh, uint32_t* aabb) { *bvh = sizeof(BVH); *aabb = sizeof(AABB); return 0; }

// BVH::buildNode (BVH.hh:273-314): appends one leaf; returns it in node_out (64 bytes).
void refb_build_node(const float* box_min, const float* box_max, const float* model16, int32_t pType, uint32_t pIndex, void* node_out) {
    AABB box;
    box.mini = float3(box_min[0], box_min[1], box_min[2]);
    box.maxi = float3(box_max[0], box_max[1], box_max[2]);
    float4x4 m = matrix_identity_float4x4;
    if (model16)
        for (int c = 0; c < 4; ++c) m.columns[c] = simd_make_float4(model16[4 * c], model16[4 * c + 1], model16[4 * c + 2], model16[4 * c + 3]);
    std::vector<BVH> list;
    BVH::buildNode(box, m, (PrimitiveType)pType, pIndex, list);
    memcpy(node_out, &list[0], sizeof(BVH));
}

// BVH::buildTree (BVH.hh:246-269): bvhList holds nLeaves leaves on entry (capacity 2*nLeaves-1), the whole tree on return.
uint32_t refb_build_tree(void* bvhList, uint32_t nLeaves) {
    std::vector<BVH> list(nLeaves);
    memcpy((void*)list.data(), bvhList, (size_t)nLeaves * sizeof(BVH));
    BVH::buildTree(list);
    memcpy(bvhList, (const void*)list.data(), list.size() * sizeof(BVH));
    return (uint32_t)list.size();
}

}  // extern "C"
