// tracer_b200/csrc/host/layout.h
// Reference byte layouts (Metal / Apple-simd rules: float3 is 16 B) as plain structs shared
// by the host code and the CUDA kernels. Sizes are the ones the reference compiles to:
//   BVH 64 (BVH.hh:15-22), AABB 32 (AABB.hh:7-9), TriangleVertex 32 (Triangle.hh:12-18),
//   Sphere 272 (Sphere.hh:6-15), Square 272 (Square.hh:12-27), Cube 240 (Cube.hh:6-13).
#pragma once
#include <stdint.h>

namespace trq {

struct RefAABB { float mini[3]; float pad0; float maxi[3]; float pad1; };
struct RefBVH {
    uint32_t parent, left, right, axis;
    int32_t  pType;
    uint32_t pIndex;
    uint32_t pad[2];
    RefAABB  bBOX;
};
struct RefVertex { float v[3]; float n[3]; float uv[2]; };
struct RefSphere {
    float radius; float pad0[3];
    float center[3]; float pad1;
    float model[16], normal[16], inverse[16];
    uint32_t material; uint32_t pad2[3];
    RefAABB boundingBOX;
};
struct RefSquare {
    uint8_t axis_i, axis_j; uint8_t pad0[6];
    float range_i[2];
    float range_j[2];
    uint8_t axis_k; uint8_t pad1[3];
    float value_k;
    float model[16], normal[16], inverse[16];
    uint32_t material; uint32_t pad2[3];
    RefAABB boundingBOX;
};
struct RefCube {
    float model[16], normal[16], inverse[16];
    RefAABB box;
    uint32_t material; uint32_t pad[3];
};

static_assert(sizeof(RefAABB) == 32, "AABB layout");
static_assert(sizeof(RefBVH) == 64, "BVH layout");
static_assert(sizeof(RefVertex) == 32, "TriangleVertex layout");
static_assert(sizeof(RefSphere) == 272, "Sphere layout");
static_assert(sizeof(RefSquare) == 272, "Square layout");
static_assert(sizeof(RefCube) == 240, "Cube layout");

}  // namespace trq
