// tracer_b200/csrc/bvh_build_gpu.cu -- trq_bvh_build_tree_gpu: BVH::buildTree (BVH.hh:246-269) on the GPU.
// Orchestrates the level-synchronous kernels of kernels/bvh_build.cuh; same node array as trq_bvh_build_tree.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/tracer_rq.h"
#include "host/error.h"
#include "host/layout.h"
#include "kernels/bvh_build.cuh"

using namespace trq;
using namespace trq::gpubuild;

namespace trq { void note_launches(uint64_t n); }

namespace {

struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <typename T> bool alloc(T** out, size_t count) {
        void* p = nullptr;
        if (cudaMalloc(&p, (count ? count : 1) * sizeof(T)) != cudaSuccess) return false;
        ptrs.push_back(p);
        *out = (T*)p;
        return true;
    }
};

#define BCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            if (prev >= 0) cudaSetDevice(prev);                                                     \
            return trq::fail(TRQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                           \
    } while (0)

inline unsigned blocks(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

}  // namespace

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
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return trq::fail(TRQ_ERR_NO_DEVICE, "trq_bvh_build_tree_gpu: no CUDA device (use trq_bvh_build_tree on the host)"); }
    if (device < 0 || device >= ndev) return trq::fail(TRQ_ERR_INVALID, "trq_bvh_build_tree_gpu: device %d out of range", device);
    int prev = -1;
    cudaGetDevice(&prev);
    BCK(cudaSetDevice(device));

    Scratch mem;
    RefBVH *dLeaves, *dOut;
    uint32_t *idx, *seg, *flag, *scanT, *tileSums, *leftFalse, *rightTrue, *arrivals, *counters;
    float4* cen;
    BNode *tabA, *tabB;
    const uint32_t cap = n / 2 + 1;
    const uint32_t nTiles = blocks((uint64_t)n + 1, kScanTile);
    bool ok = mem.alloc(&dLeaves, n) && mem.alloc(&dOut, nNode) && mem.alloc(&idx, n) && mem.alloc(&seg, n) &&
              mem.alloc(&flag, (size_t)n + 1) && mem.alloc(&scanT, (size_t)n + 1) && mem.alloc(&tileSums, nTiles) &&
              mem.alloc(&leftFalse, n) && mem.alloc(&rightTrue, n) && mem.alloc(&arrivals, nNode) &&
              mem.alloc(&counters, 4) && mem.alloc(&cen, n) && mem.alloc(&tabA, cap) && mem.alloc(&tabB, cap);
    if (!ok) { if (prev >= 0) cudaSetDevice(prev); cudaGetLastError(); return trq::fail(TRQ_ERR_NOMEM, "trq_bvh_build_tree_gpu: out of device memory"); }

    BCK(cudaMemcpy(dLeaves, host, (size_t)n * sizeof(RefBVH), cudaMemcpyHostToDevice));
    BCK(cudaMemset(dOut, 0, (size_t)nNode * sizeof(RefBVH)));
    BCK(cudaMemcpy(dOut + 1, dLeaves, (size_t)n * sizeof(RefBVH), cudaMemcpyDeviceToDevice));     // leaves at 1..N (BVH.hh:265)
    BCK(cudaMemset(arrivals, 0, (size_t)nNode * sizeof(uint32_t)));
    BCK(cudaMemset(counters, 0, 4 * sizeof(uint32_t)));
    uint64_t launches = 0;
    init_elements_kernel<<<blocks(n, 256), 256>>>(dLeaves, n, idx, seg, cen); ++launches;
    BNode root; std::memset(&root, 0, sizeof root);
    root.start = 0; root.end = n; root.base = 0;
    BCK(cudaMemcpy(tabA, &root, sizeof root, cudaMemcpyHostToDevice));

    BNode *cur = tabA, *nxt = tabB;
    uint32_t nActive = 1, depth = 0;
    uint32_t* nNext = counters;          // [0] next-level count, [1] max depth
    uint32_t* dMaxDepth = counters + 1;
    while (nActive > 0) {
        if (depth > 64) { if (prev >= 0) cudaSetDevice(prev); return trq::fail(TRQ_ERR_DEPTH, "trq_bvh_build_tree_gpu: tree deeper than 64 levels"); }
        reset_nodes_kernel<<<blocks(nActive, 256), 256>>>(cur, nActive);
        centroid_bounds_kernel<<<blocks(n, 256), 256>>>(idx, seg, cen, n, cur);
        choose_axis_kernel<<<blocks(nActive, 256), 256>>>(cur, nActive);
        bucket_kernel<<<blocks(n, 256), 256>>>(dLeaves, idx, seg, cen, n, cur);
        choose_split_kernel<<<blocks(nActive, 128), 128>>>(cur, nActive);
        predicate_kernel<<<blocks((uint64_t)n + 1, 256), 256>>>(idx, seg, cen, n, cur, flag);
        scan_tiles_kernel<<<nTiles, 1024>>>(flag, scanT, n + 1, tileSums);
        scan_sums_kernel<<<1, 1024>>>(tileSums, nTiles);
        scan_add_kernel<<<nTiles, 1024>>>(scanT, n + 1, tileSums);
        midpoint_kernel<<<blocks(nActive, 256), 256>>>(cur, nActive, scanT);
        mispl_kernel<<<blocks(n, 256), 256>>>(seg, n, cur, flag, scanT, leftFalse, rightTrue);
        swap_kernel<<<blocks(n, 256), 256>>>(seg, n, cur, leftFalse, rightTrue, idx);
        BCK(cudaMemsetAsync(nNext, 0, sizeof(uint32_t)));
        emit_kernel<<<blocks(nActive, 128), 128>>>(cur, nActive, idx, cen, dOut, n, nxt, nNext, dMaxDepth, depth);
        reseg_kernel<<<blocks(n, 256), 256>>>(seg, n, cur);
        launches += 14;
        BCK(cudaMemcpy(&nActive, nNext, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (nActive > cap) { if (prev >= 0) cudaSetDevice(prev); return trq::fail(TRQ_ERR_LAYOUT, "trq_bvh_build_tree_gpu: internal: level table overflow"); }
        BNode* t = cur; cur = nxt; nxt = t;
        ++depth;
    }
    refit_kernel<<<blocks(n, 256), 256>>>(dOut, n, arrivals); ++launches;
    uint32_t maxDepth = 0;
    BCK(cudaMemcpy(&maxDepth, dMaxDepth, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    BCK(cudaMemcpy(host, dOut, (size_t)nNode * sizeof(RefBVH), cudaMemcpyDeviceToHost));
    BCK(cudaGetLastError());
    trq::note_launches(launches);
    if (prev >= 0) cudaSetDevice(prev);
    if (nNodeOut) *nNodeOut = nNode;
    if (maxDepthOut) *maxDepthOut = maxDepth;
    if (maxDepth > 31)
        return trq::fail(TRQ_ERR_DEPTH, "trq_bvh_build_tree_gpu: interior depth %u exceeds the 32-bit trail", maxDepth);
    return TRQ_OK;
}
