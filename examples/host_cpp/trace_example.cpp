// examples/host_cpp/trace_example.cpp -- C++ host driving the ray query through the C-ABI only.
//
// Mirrors what RT_Metal's host does around the hot path (AAPLRenderer.mm:546-624, Render.metal:523-532):
// make per-triangle leaves, build the BVH, upload the scene, cast a small grid of camera rays, trace
// closest-hit and any-hit, expand the hit records, then shard the same batch over every visible GPU. Build (from the repo root):
//   g++ -std=c++17 -O2 -Iinclude examples/host_cpp/trace_example.cpp -Ltracer_b200 -ltracer_rq \
//       -Wl,-rpath,$PWD/tracer_b200 -o examples/host_cpp/trace_example
// Exit code 0 on success, 3 when there is no CUDA device (there is no CPU fallback), 1 on any other error.
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <vector>

#include "tracer_rq.h"
#include "tracer_rq_harness.h"

struct Vertex { float v[3], n[3], uv[2]; };          // TriangleVertex, 32 B (Triangle.hh:12-18)
struct Node { unsigned char bytes[64]; };            // struct BVH, 64 B (BVH.hh:15-22)

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        if (rc_ != TRQ_OK) {                                                                \
            std::fprintf(stderr, "%s -> %d: %s\n", #call, rc_, trq_last_error_string());   \
            return rc_ == TRQ_ERR_NO_DEVICE ? 3 : 1;                                        \
        }                                                                                   \
    } while (0)

int main() {
    // a 2x2 m floor (two triangles) with a pyramid of four triangles standing on it
    const float P[][3] = {{-1, 0, -1}, {1, 0, -1}, {1, 0, 1}, {-1, 0, 1}, {-0.5f, 0, -0.5f}, {0.5f, 0, -0.5f},
                          {0.5f, 0, 0.5f}, {-0.5f, 0, 0.5f}, {0, 1, 0}};
    const uint32_t I[] = {0, 1, 2, 0, 2, 3, 4, 5, 8, 5, 6, 8, 6, 7, 8, 7, 4, 8};
    const uint32_t nTri = sizeof(I) / sizeof(I[0]) / 3;
    std::vector<Vertex> verts(9);
    for (int k = 0; k < 9; ++k) {
        std::memcpy(verts[k].v, P[k], 12);
        verts[k].n[0] = 0; verts[k].n[1] = 1; verts[k].n[2] = 0;
        verts[k].uv[0] = verts[k].uv[1] = 0;
    }

    std::vector<Node> bvh(2 * nTri - 1);
    CHECK(trq_bvh_build_nodes_triangles(verts.data(), I, nTri, 0, bvh.data()));
    uint32_t nNode = 0, depth = 0;
    CHECK(trq_bvh_build_tree(bvh.data(), nTri, &nNode, &depth));

    trq_scene_desc desc = {};
    desc.triList = verts.data(); desc.nVert = (uint32_t)verts.size();
    desc.idxList = I;            desc.nTri = nTri;
    desc.bvhList = bvh.data();   desc.nNode = nNode;
    trq_scene* scene = nullptr;
    CHECK(trq_scene_create(&desc, 0, &scene));

    const uint32_t W = 8, H = 6;
    std::vector<trq_ray> rays(W * H);
    const float from[3] = {0, 2.5f, -4}, at[3] = {0, 0.3f, 0}, up[3] = {0, 1, 0};
    trqh_gen_camera_rays(from, at, up, 0.6f, float(W) / H, 1.0f, W, H, rays.data());

    std::vector<trq_hit> hits(rays.size()), occl(rays.size());
    std::vector<trq_hit_record> recs(rays.size());
    CHECK(trq_trace(scene, rays.data(), rays.size(), TRQ_HOST_PTRS, hits.data(), nullptr));
    CHECK(trq_trace(scene, rays.data(), rays.size(), TRQ_HOST_PTRS | TRQ_TRACE_ANY, occl.data(), nullptr));
    CHECK(trq_expand_hits(scene, rays.data(), hits.data(), rays.size(), TRQ_HOST_PTRS, recs.data(), nullptr));

    unsigned nHit = 0;
    for (uint32_t y = H; y-- > 0;) {
        for (uint32_t x = 0; x < W; ++x) {
            const trq_hit& h = hits[y * W + x];
            if (((h.flags ^ occl[y * W + x].flags) & TRQ_HIT_FLAG_HIT) != 0) { std::fprintf(stderr, "any/closest disagree\n"); return 1; }
            if (h.flags & TRQ_HIT_FLAG_HIT) { ++nHit; std::putchar(h.pIndex < 2 ? '.' : 'A' + (int)h.pIndex - 2); }
            else std::putchar(' ');
        }
        std::putchar('\n');
    }
    std::printf("%u nodes, depth %u, %u of %zu rays hit; centre hit p = (%.3f, %.3f, %.3f) t = %.3f\n", nNode, depth, nHit,
                rays.size(), recs[(H / 2) * W + W / 2].p[0], recs[(H / 2) * W + W / 2].p[1], recs[(H / 2) * W + W / 2].p[2],
                recs[(H / 2) * W + W / 2].t);
    CHECK(trq_scene_destroy(scene));

    // the same batch sharded over every visible GPU from this one process (one scene per device, no collective)
    trq_mgpu* all = nullptr;
    CHECK(trq_mgpu_create(&desc, nullptr, 0, &all));
    std::vector<trq_hit> sharded(rays.size());
    CHECK(trq_mgpu_trace(all, rays.data(), rays.size(), 0, sharded.data()));
    if (std::memcmp(sharded.data(), hits.data(), hits.size() * sizeof(trq_hit)) != 0) { std::fprintf(stderr, "sharded trace differs\n"); return 1; }
    std::printf("sharded over %d device(s): identical\n", trq_mgpu_device_count(all));
    CHECK(trq_mgpu_destroy(all));
    return nHit > 0 ? 0 : 1;
}
