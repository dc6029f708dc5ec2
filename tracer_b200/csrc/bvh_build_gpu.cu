// tracer_b200/csrc/bvh_build_gpu.cu -- BVH::buildTree (BVH.hh:246-269) on the GPU.
//   trq_bvh_build_tree_device   leaves and result in DEVICE memory (feeds trq_scene_create_device without a host round trip)
//   trq_bvh_build_tree_gpu      the same build for a HOST array (copy in, build, copy out)
// Orchestrates the level-synchronous kernels of kernels/bvh_build.cuh; same node array as trq_bvh_build_tree.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/tracer_rq.h"
#include "host/error.h"
#include "host/layout.h"
#include "host/scratch.h"
#include "kernels/bvh_build.cuh"

using namespace trq;
using namespace trq::gpubuild;

namespace trq {
void note_launches(uint64_t n);

// One stream-ordered pool per device that keeps its memory between calls: scratch for the builder and for scene
// creation comes back without a cudaMalloc after the first use.
cudaMemPool_t scratch_pool(int device) {
    static std::mutex m;
    static std::map<int, cudaMemPool_t> pools;
    std::lock_guard<std::mutex> lock(m);
    auto it = pools.find(device);
    if (it != pools.end()) return it->second;
    cudaMemPoolProps pp = {};
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &pp) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pools[device] = pool;
    return pool;
}

}  // namespace trq

namespace {

#define BCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return trq::fail(TRQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline unsigned blocks(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

struct DeviceScope {
    int prev = -1;
    bool ok = false;
    explicit DeviceScope(int dev) { cudaGetDevice(&prev); ok = cudaSetDevice(dev) == cudaSuccess; }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

// dLeaves[0..n) -> dOut[0..2n-1): root at 0, leaves at 1..n in list order, interiors after (BVH.hh:263-268). n >= 2.
int build_on_device(const RefBVH* dLeaves, uint32_t n, RefBVH* dOut, int device, uint32_t* maxDepthOut) {
    const uint32_t nNode = 2 * n - 1;
    cudaStream_t st = nullptr;                          // legacy default stream: ordered with the caller's cudaMemcpy / kernels
    PoolScratch mem(device, st);
    uint32_t *idx, *seg, *flag, *scanT, *tileSums, *leftFalse, *rightTrue, *arrivals, *counters;
    float4* cen;
    BNode *tabA, *tabB;
    const uint32_t cap = n / 2 + 1;
    const uint32_t nTiles = blocks((uint64_t)n + 1, kScanTile);
    bool ok = mem.alloc(&idx, n) && mem.alloc(&seg, n) &&
              mem.alloc(&flag, (size_t)n + 1) && mem.alloc(&scanT, (size_t)n + 1) && mem.alloc(&tileSums, nTiles) &&
              mem.alloc(&leftFalse, n) && mem.alloc(&rightTrue, n) && mem.alloc(&arrivals, nNode) &&
              mem.alloc(&counters, 4) && mem.alloc(&cen, n) && mem.alloc(&tabA, cap) && mem.alloc(&tabB, cap);
    if (!ok) return trq::fail(TRQ_ERR_NOMEM, "BVH build on the GPU: out of device memory");

    BCK(cudaMemsetAsync(dOut, 0, (size_t)nNode * sizeof(RefBVH), st));
    BCK(cudaMemcpyAsync(dOut + 1, dLeaves, (size_t)n * sizeof(RefBVH), cudaMemcpyDeviceToDevice, st));     // leaves at 1..N (BVH.hh:265)
    BCK(cudaMemsetAsync(arrivals, 0, (size_t)nNode * sizeof(uint32_t), st));
    BCK(cudaMemsetAsync(counters, 0, 4 * sizeof(uint32_t), st));
    uint64_t launches = 0;
    init_elements_kernel<<<blocks(n, 256), 256, 0, st>>>(dLeaves, n, idx, seg, cen); ++launches;
    BNode root; std::memset(&root, 0, sizeof root);
    root.start = 0; root.end = n; root.base = 0;
    BCK(cudaMemcpyAsync(tabA, &root, sizeof root, cudaMemcpyHostToDevice, st));
    BCK(cudaStreamSynchronize(st));                     // `root` lives on this stack frame

    BNode *cur = tabA, *nxt = tabB;
    uint32_t nActive = 1, depth = 0;
    uint32_t* nNext = counters;          // [0] next-level count, [1] max depth, [2] level-table overflow
    uint32_t* dMaxDepth = counters + 1;
    uint32_t* dOverflow = counters + 2;
    while (nActive > 0) {
        if (depth > 64) return trq::fail(TRQ_ERR_DEPTH, "BVH build on the GPU: tree deeper than 64 levels");
        reset_nodes_kernel<<<blocks(nActive, 256), 256, 0, st>>>(cur, nActive);
        centroid_bounds_kernel<<<blocks(n, 256), 256, 0, st>>>(idx, seg, cen, n, cur);
        choose_axis_kernel<<<blocks(nActive, 256), 256, 0, st>>>(cur, nActive);
        bucket_kernel<<<blocks(n, 256), 256, 0, st>>>(dLeaves, idx, seg, cen, n, cur);
        choose_split_kernel<<<blocks(nActive, 128), 128, 0, st>>>(cur, nActive);
        predicate_kernel<<<blocks((uint64_t)n + 1, 256), 256, 0, st>>>(idx, seg, cen, n, cur, flag);
        scan_tiles_kernel<<<nTiles, 1024, 0, st>>>(flag, scanT, n + 1, tileSums);
        scan_sums_kernel<<<1, 1024, 0, st>>>(tileSums, nTiles);
        scan_add_kernel<<<nTiles, 1024, 0, st>>>(scanT, n + 1, tileSums);
        midpoint_kernel<<<blocks(nActive, 256), 256, 0, st>>>(cur, nActive, scanT);
        mispl_kernel<<<blocks(n, 256), 256, 0, st>>>(seg, n, cur, flag, scanT, leftFalse, rightTrue);
        swap_kernel<<<blocks(n, 256), 256, 0, st>>>(seg, n, cur, leftFalse, rightTrue, idx);
        BCK(cudaMemsetAsync(nNext, 0, sizeof(uint32_t), st));
        emit_kernel<<<blocks(nActive, 128), 128, 0, st>>>(cur, nActive, idx, cen, dOut, n, nxt, cap, nNext, dOverflow, dMaxDepth, depth);
        reseg_kernel<<<blocks(n, 256), 256, 0, st>>>(seg, n, cur);
        launches += 14;
        uint32_t lvl[3] = {0, 0, 0};
        BCK(cudaMemcpyAsync(lvl, counters, sizeof lvl, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        nActive = lvl[0];
        if (lvl[2] || nActive > cap) return trq::fail(TRQ_ERR_LAYOUT, "BVH build on the GPU: internal: level table overflow");
        BNode* t = cur; cur = nxt; nxt = t;
        ++depth;
    }
    refit_kernel<<<blocks(n, 256), 256, 0, st>>>(dOut, n, arrivals); ++launches;
    uint32_t maxDepth = 0;
    BCK(cudaMemcpyAsync(&maxDepth, dMaxDepth, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    BCK(cudaGetLastError());
    trq::note_launches(launches);
    if (maxDepthOut) *maxDepthOut = maxDepth;
    if (maxDepth > 31)
        return trq::fail(TRQ_ERR_DEPTH, "BVH build on the GPU: interior depth %u exceeds the 32-bit trail (Render.hh:140)", maxDepth);
    return TRQ_OK;
}

int check_device(int device, const char* who) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return trq::fail(TRQ_ERR_NO_DEVICE, "%s: no CUDA device (use trq_bvh_build_tree on the host)", who); }
    if (device < 0 || device >= ndev) return trq::fail(TRQ_ERR_INVALID, "%s: device %d out of range", who, device);
    return TRQ_OK;
}

}  // namespace

extern "C" int trq_bvh_build_tree_device(void* d_bvhList, uint32_t nLeaves, int device, uint32_t* nNodeOut, uint32_t* maxDepthOut) {
    if (!d_bvhList || nLeaves == 0) return trq::fail(TRQ_ERR_INVALID, "trq_bvh_build_tree_device: empty leaf list");
    if (nLeaves > 0x3fffffffu) return trq::fail(TRQ_ERR_INVALID, "trq_bvh_build_tree_device: too many leaves");
    int rc = check_device(device, "trq_bvh_build_tree_device");
    if (rc != TRQ_OK) return rc;
    DeviceScope scope(device);
    if (!scope.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    RefBVH* d = (RefBVH*)d_bvhList;
    const uint32_t n = nLeaves;
    if (nNodeOut) *nNodeOut = 2 * n - 1;
    if (n == 1) {                                   // BVH.hh:52-54 + :263-268: the lone leaf becomes node 0
        BCK(cudaMemset(&d[0].parent, 0, sizeof(uint32_t)));
        if (maxDepthOut) *maxDepthOut = 0;
        return TRQ_OK;
    }
    // the result overwrites the leaves' slots (they move to 1..n), so the build reads a copy of them
    PoolScratch mem(device, nullptr);
    RefBVH* dLeaves = nullptr;
    if (!mem.alloc(&dLeaves, n)) return trq::fail(TRQ_ERR_NOMEM, "trq_bvh_build_tree_device: out of device memory");
    BCK(cudaMemcpyAsync(dLeaves, d, (size_t)n * sizeof(RefBVH), cudaMemcpyDeviceToDevice, nullptr));
    return build_on_device(dLeaves, n, d, device, maxDepthOut);
}

extern "C" int trq_bvh_build_tree_gpu(void* bvhList, uint32_t nLeaves, int device, uint32_t* nNodeOut, uint32_t* maxDepthOut) {
    if (!bvhList || nLeaves == 0) return trq::fail(TRQ_ERR_INVALID, "trq_bvh_build_tree_gpu: empty leaf list");
    if (nLeaves > 0x3fffffffu) return trq::fail(TRQ_ERR_INVALID, "trq_bvh_build_tree_gpu: too many leaves");
    const uint32_t n = nLeaves, nNode = 2 * n - 1;
    RefBVH* host = (RefBVH*)bvhList;
    if (n == 1) {                                   // BVH.hh:52-54 + :263-268: the lone leaf becomes node 0
        host[0].parent = 0;
        if (nNodeOut) *nNodeOut = 1;
        if (maxDepthOut) *maxDepthOut = 0;
        return TRQ_OK;
    }
    int rc = check_device(device, "trq_bvh_build_tree_gpu");
    if (rc != TRQ_OK) return rc;
    DeviceScope scope(device);
    if (!scope.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    PoolScratch mem(device, nullptr);
    RefBVH *dLeaves = nullptr, *dOut = nullptr;
    if (!mem.alloc(&dLeaves, n) || !mem.alloc(&dOut, nNode)) return trq::fail(TRQ_ERR_NOMEM, "trq_bvh_build_tree_gpu: out of device memory");
    BCK(cudaMemcpyAsync(dLeaves, host, (size_t)n * sizeof(RefBVH), cudaMemcpyHostToDevice, nullptr));
    uint32_t maxDepth = 0;
    rc = build_on_device(dLeaves, n, dOut, device, &maxDepth);
    if (nNodeOut) *nNodeOut = nNode;
    if (maxDepthOut) *maxDepthOut = maxDepth;
    if (rc != TRQ_OK && rc != TRQ_ERR_DEPTH) return rc;
    BCK(cudaMemcpy(host, dOut, (size_t)nNode * sizeof(RefBVH), cudaMemcpyDeviceToHost));
    return rc;
}
