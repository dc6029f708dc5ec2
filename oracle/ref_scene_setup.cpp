// oracle/ref_scene_setup.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's own scene set-up code (RT_Metal/Tracer/Tracer.mm: MakeCamera :87-125, MakeSquare / MakeCube /
// MakeSphere :127-172, prepareCubeList :174-243, prepareCornellBox :245-304, prepareSphereList :306-369,
// prepareCamera :371-411), compiled from the source where it lies and exported through a C ABI, so that the scene
// constants and the camera the workload generators restate (tracer_b200/harness.py, csrc/host/harness.cpp) are checked
// against what the reference builds. Tracer.mm is Objective-C++ only by extension: it is plain C++ over Apple simd,
// which oracle/shim_host/simd/simd.h stands in for (see there for what is exact and what is a tolerance).
// BVH.hh arrives first through /dev/stdin with its clang blocks turned into lambdas (oracle/Makefile), which also
// satisfies Tracer.hh's own #include "BVH.hh" through the include guard.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include <dispatch/dispatch.h>

#include "/dev/stdin"        // RT_Metal/Metal/BVH.hh with `^{` -> `[&]{`
#include "Tracer.mm"         // -I$(REF)/RT_Metal/Tracer

static_assert(sizeof(Square) == 272 && sizeof(Cube) == 240 && sizeof(Sphere) == 272, "host layout == Metal layout");

template <typename T>
static uint32_t copy_out(const std::vector<T>& list, void* out, uint32_t cap) {
    const uint32_t n = (uint32_t)list.size() < cap ? (uint32_t)list.size() : cap;
    if (out) memcpy(out, (const void*)list.data(), (size_t)n * sizeof(T));
    return (uint32_t)list.size();
}

static void camera_out(const Camera& c, float* out18) {
    const float3 f[6] = {c.lookFrom, c.u, c.v, c.vertical, c.horizontal, c.cornerLowLeft};
    for (int k = 0; k < 6; ++k) { out18[3 * k] = f[k].x; out18[3 * k + 1] = f[k].y; out18[3 * k + 2] = f[k].z; }
}

extern "C" {

// The three lists as the application builds them: one shared material list, cubes first, then the Cornell box, then the
// spheres (AAPLRenderer.mm:216,222,228) -- the material indices stored in the primitives depend on that order.
struct SceneLists { std::vector<Cube> cubes; std::vector<Square> squares; std::vector<Sphere> spheres; std::vector<Material> materials; };
static const SceneLists& scene_lists() {
    static const SceneLists L = [] {
        SceneLists l;
        prepareCubeList(l.cubes, l.materials);
        prepareCornellBox(l.squares, l.materials);
        prepareSphereList(l.spheres, l.materials);
        return l;
    }();
    return L;
}
uint32_t refs_cornell_squares(void* out, uint32_t cap) { return copy_out(scene_lists().squares, out, cap); }
uint32_t refs_cubes(void* out, uint32_t cap) { return copy_out(scene_lists().cubes, out, cap); }
uint32_t refs_spheres(void* out, uint32_t cap) { return copy_out(scene_lists().spheres, out, cap); }
uint32_t refs_material_count(void) { return (uint32_t)scene_lists().materials.size(); }
// prepareCamera with no mouse / key offsets: {lookFrom, u, v, vertical, horizontal, cornerLowLeft}
void refs_prepare_camera(float viewW, float viewH, float* out18) {
    Camera c;
    prepareCamera(&c, float2(viewW, viewH), float2(0, 0), float3(0, 0, 0));
    camera_out(c, out18);
}
void refs_make_camera(const float* from, const float* at, const float* up, float aperture, float aspect, float vfov, float focus, float* out18) {
    Camera c;
    MakeCamera(&c, float3(from[0], from[1], from[2]), float3(at[0], at[1], at[2]), float3(up[0], up[1], up[2]), aperture, aspect, vfov, focus);
    camera_out(c, out18);
}

}  // extern "C"
