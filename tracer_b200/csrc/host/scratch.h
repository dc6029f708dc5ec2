// tracer_b200/csrc/host/scratch.h -- stream-ordered device scratch from a per-device pool that keeps its memory between
// calls (the BVH builder and scene creation get their temporaries back without a cudaMalloc after the first use).
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace trq {

cudaMemPool_t scratch_pool(int device);            // bvh_build_gpu.cu

// Allocations released together when the scope ends -- also on error paths, after the stream has drained.
struct PoolScratch {
    cudaMemPool_t pool;
    cudaStream_t st;
    std::vector<void*> ptrs;
    PoolScratch(int device, cudaStream_t s) : pool(scratch_pool(device)), st(s) {}
    PoolScratch(const PoolScratch&) = delete;
    PoolScratch& operator=(const PoolScratch&) = delete;
    ~PoolScratch() {
        if (ptrs.empty()) return;
        cudaStreamSynchronize(st);                 // kernels of a failed call may still be running
        for (void* p : ptrs) cudaFreeAsync(p, st);
    }
    template <typename T> bool alloc(T** out, size_t count) {
        void* p = nullptr;
        const size_t bytes = (count ? count : 1) * sizeof(T);
        cudaError_t e = pool ? cudaMallocFromPoolAsync(&p, bytes, pool, st) : cudaMallocAsync(&p, bytes, st);
        if (e != cudaSuccess) { cudaGetLastError(); return false; }
        ptrs.push_back(p);
        *out = (T*)p;
        return true;
    }
};

}  // namespace trq
