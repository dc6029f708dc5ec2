// tracer_b200/csrc/trq_api.cu -- implementation of the C-ABI in include/tracer_rq.h.
//
// Host side of the drop-in boundary: it takes the six reference-layout arrays exactly as
// AAPLRenderer.mm:614-624,712-720 hands them to Metal (struct Primitive, Render.hh:122-130),
// uploads them, derives the packed traversal layout ON THE DEVICE (pack_scene_kernel), and runs
// batches of Scene::hit (Render.hh:135-252) as CUDA kernels. No CPU fallback exists: without a
// device every compute entry point returns TRQ_ERR_NO_DEVICE.

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/tracer_rq.h"
#include "host/error.h"
#include "host/layout.h"
#include "kernels/trace_kernels.cuh"

using namespace trq;

namespace {

std::atomic<uint64_t> g_launches{0};

}  // namespace

namespace trq { void note_launches(uint64_t n) { g_launches += n; } }   // other translation units (GPU builder)

namespace {

#define TRQ_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return trq::fail(TRQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr int kCounterRing = 256;
constexpr int kStageBufs = 4;
constexpr int kProfRing = 64;
#ifndef TRQ_DEFAULT_CHUNK_RAYS
#define TRQ_DEFAULT_CHUNK_RAYS (512ll << 10)
#endif

}  // namespace

struct trq_scene {
    int device = 0;
    int numSMs = 0;
    trq_scene_info_t info{};
    // reference-layout device copies (struct Primitive)
    RefSphere* d_spheres = nullptr;
    RefSquare* d_squares = nullptr;
    RefCube*   d_cubes = nullptr;
    RefVertex* d_verts = nullptr;
    uint32_t*  d_idx = nullptr;
    RefBVH*    d_bvh = nullptr;
    // packed layout
    float4* d_nodes = nullptr;
    float4* d_tris = nullptr;
    float4* d_sph = nullptr;
    float4* d_triN = nullptr;
    bool allTriangles = false;    // every leaf is a triangle: the trace kernel finishes its own records (no resolve pass)
    SceneDev dev{};
    uint32_t stackDepth = 1;
    size_t traceSmem = 0;
    int blocksPerSM[2][2] = {};   // resident trace_packed_kernel CTAs per SM: [closest-hit, any-hit][compact result, fused finish]
    // stream-ordered scratch for TRQ_SORT_RAYS: a private pool that keeps its memory across synchronisations
    // (the device's default pool hands it back at every sync and re-maps 64 MB on the next sorted launch)
    cudaMemPool_t scratchPool = nullptr;
    // ray-queue heads
    unsigned long long* d_counters = nullptr;
    std::atomic<uint32_t> counterNext{0};
    // staging for TRQ_HOST_PTRS
    std::mutex stageMutex;
    uint64_t stageCap = 0;
    trq_ray* d_stageRays[kStageBufs] = {};
    trq_hit* d_stageHits[kStageBufs] = {};
    cudaStream_t stageStream[kStageBufs] = {};   // one stream per staging buffer: copy in, trace, copy out in stream order
    bool stageReady = false;
    uint64_t stageSeq = 0;        // chunks ever staged (ring position)
    // optional per-kernel timing (trq_profile_enable): events around the trace and resolve kernels
    bool profile = false;
    cudaEvent_t evProf[kProfRing][3] = {};
    uint32_t profHead = 0, profCount = 0;
};

namespace {

void free_scene(trq_scene* s) {
    if (!s) return;
    cudaFree(s->d_spheres); cudaFree(s->d_squares); cudaFree(s->d_cubes);
    cudaFree(s->d_verts); cudaFree(s->d_idx); cudaFree(s->d_bvh);
    cudaFree(s->d_nodes); cudaFree(s->d_tris); cudaFree(s->d_sph); cudaFree(s->d_triN);
    cudaFree(s->d_counters);
    if (s->scratchPool) cudaMemPoolDestroy(s->scratchPool);
    for (int b = 0; b < kStageBufs; ++b) {
        cudaFree(s->d_stageRays[b]); cudaFree(s->d_stageHits[b]);
        if (s->stageStream[b]) cudaStreamDestroy(s->stageStream[b]);
    }
    for (int k = 0; k < kProfRing; ++k)
        for (int j = 0; j < 3; ++j) if (s->evProf[k][j]) cudaEventDestroy(s->evProf[k][j]);
    delete s;
}

template <typename T>
int upload(T** dst, const void* src, size_t count) {
    *dst = nullptr;
    if (count == 0) return TRQ_OK;
    TRQ_CUDA(cudaMalloc((void**)dst, count * sizeof(T)));
    TRQ_CUDA(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return TRQ_OK;
}

// Walk the tree once on the host: validate the layout contract and assign every reachable node its
// packed reference (interior nodes and leaves numbered in depth-first order, left child first).
int plan_layout(const trq_scene_desc* d, std::vector<uint32_t>& ref, trq_scene_info_t& info) {
    const RefBVH* N = (const RefBVH*)d->bvhList;
    const uint32_t n = d->nNode;
    ref.assign(n, TRQ_REF_DONE_WORD);
    uint32_t nInterior = 0, nTriLeaf = 0, nSphLeaf = 0, nLeaf = 0, maxDepth = 0;
    struct Item { uint32_t node, depth; };
    std::vector<Item> stack;
    stack.push_back({0u, 0u});
    while (!stack.empty()) {
        const Item it = stack.back(); stack.pop_back();
        const uint32_t i = it.node;
        if (i >= n) return trq::fail(TRQ_ERR_LAYOUT, "bvhList: child index %u out of range (nNode %u)", i, n);
        if (ref[i] != TRQ_REF_DONE_WORD) return trq::fail(TRQ_ERR_LAYOUT, "bvhList: node %u reachable twice (not a tree)", i);
        const RefBVH& b = N[i];
        if (b.pType == TRQ_BVH) {
            if (it.depth > 31) return trq::fail(TRQ_ERR_DEPTH, "bvhList: interior depth %u exceeds the 32-bit trail (Render.hh:140)", it.depth);
            if (b.left == 0 || b.right == 0 || b.left == b.right)
                return trq::fail(TRQ_ERR_LAYOUT, "bvhList: interior node %u has invalid children (%u, %u)", i, b.left, b.right);
            if (nInterior >= 0x1fffffffu) return trq::fail(TRQ_ERR_LAYOUT, "bvhList: too many interior nodes");
            ref[i] = TRQ_MAKE_REF(REF_INTERIOR, nInterior++);
            maxDepth = it.depth > maxDepth ? it.depth : maxDepth;
            stack.push_back({b.right, it.depth + 1});     // left is popped (numbered) first
            stack.push_back({b.left, it.depth + 1});
        } else {
            ++nLeaf;
            switch (b.pType) {
                case TRQ_TRIANGLE:
                    if (b.pIndex >= d->nTri) return trq::fail(TRQ_ERR_LAYOUT, "leaf %u: triangle pIndex %u >= nTri %u", i, b.pIndex, d->nTri);
                    for (int k = 0; k < 3; ++k)
                        if (d->idxList[3 * (size_t)b.pIndex + k] >= d->nVert)
                            return trq::fail(TRQ_ERR_LAYOUT, "triangle %u: vertex index out of range", b.pIndex);
                    ref[i] = TRQ_MAKE_REF(REF_TRI, nTriLeaf++);
                    break;
                case TRQ_SPHERE:
                    if (b.pIndex >= d->nSphere) return trq::fail(TRQ_ERR_LAYOUT, "leaf %u: sphere pIndex %u >= nSphere %u", i, b.pIndex, d->nSphere);
                    ref[i] = TRQ_MAKE_REF(REF_SPHERE, nSphLeaf++);
                    break;
                case TRQ_SQUARE:
                    if (b.pIndex >= d->nSquare) return trq::fail(TRQ_ERR_LAYOUT, "leaf %u: square pIndex %u >= nSquare %u", i, b.pIndex, d->nSquare);
                    ref[i] = TRQ_MAKE_REF(REF_SQUARE, i);
                    break;
                case TRQ_CUBE:
                    if (b.pIndex >= d->nCube) return trq::fail(TRQ_ERR_LAYOUT, "leaf %u: cube pIndex %u >= nCube %u", i, b.pIndex, d->nCube);
                    ref[i] = TRQ_MAKE_REF(REF_CUBE, i);
                    break;
                default:
                    ref[i] = TRQ_MAKE_REF(REF_NOP, i);        // Render.hh:241 `default: break`
                    break;
            }
        }
    }
    info.nNode = n; info.nInterior = nInterior; info.nLeaf = nLeaf; info.maxDepth = maxDepth;
    info.nTri = nTriLeaf; info.nSphere = nSphLeaf;
    info.nSquare = d->nSquare; info.nCube = d->nCube;
    return TRQ_OK;
}

int launch_trace(trq_scene* s, const trq_ray* d_rays, uint64_t n, uint32_t flags, trq_hit* d_hits, cudaStream_t st,
                 const unsigned long long* nPtr = nullptr, const GatherDev* gather = nullptr) {
    if (n == 0 && !gather) return TRQ_OK;
    constexpr uint64_t kMaxPerLaunch = 1ull << 31;             // the kernels keep a 32-bit ray index per lane
    if (n > kMaxPerLaunch) {
        if (nPtr || gather) return trq::fail(TRQ_ERR_INVALID, "trq_trace_indirect / trq_trace_gather: more than 2^31 rays");
        for (uint64_t off = 0; off < n; off += kMaxPerLaunch) {
            const uint64_t m = (n - off) < kMaxPerLaunch ? (n - off) : kMaxPerLaunch;
            const int rc = launch_trace(s, d_rays + off, m, flags, d_hits + off, st);
            if (rc != TRQ_OK) return rc;
        }
        return TRQ_OK;
    }
    const bool any = (flags & TRQ_TRACE_ANY) != 0;
    unsigned long long* usedCounter = nullptr;
    bool fused = false;
    cudaEvent_t* prof = nullptr;
    if (s->profile) {
        prof = s->evProf[s->profHead % kProfRing];
        s->profHead++; if (s->profCount < kProfRing) s->profCount++;
    }
    if (n == 0) {
        // gather with an empty batch: nothing to trace, but the peers still wait for this rank's step
    } else if (flags & TRQ_KERNEL_REFLAYOUT) {
        const unsigned block = 128;
        const uint64_t grid = (n + block - 1) / block;
        if (grid > 0x7fffffffull) return trq::fail(TRQ_ERR_INVALID, "trq_trace: batch too large");
        if (prof) TRQ_CUDA(cudaEventRecord(prof[0], st));
        if (any) trace_reflayout_kernel<true><<<(unsigned)grid, block, 0, st>>>(s->dev, d_rays, d_hits, n, nPtr);
        else     trace_reflayout_kernel<false><<<(unsigned)grid, block, 0, st>>>(s->dev, d_rays, d_hits, n, nPtr);
        g_launches++;
    } else {
        static const uint32_t refillMin = [] {
            const char* e = getenv("TRQ_REFILL_MIN");
            int v = e ? atoi(e) : 16;      // B200 sweeps (profiles/r01_sweep*): 12..20 is flat within 2% on C3
            return (uint32_t)(v < 1 ? 1 : (v > 32 ? 32 : v));
        }();
        static const uint32_t leafBatch = [] {
            const char* e = getenv("TRQ_LEAF_BATCH");
            int v = e ? atoi(e) : 12;
            return (uint32_t)(v < 1 ? 1 : (v > 32 ? 32 : v));
        }();
        static const int blocksPerSMOverride =[] { const char* e = getenv("TRQ_BLOCKS_PER_SM"); return e ? atoi(e) : 0; }();
        const size_t smem = s->traceSmem;
        // triangle-only scene (and no peer gather): the trace kernel writes final records, there is no resolve pass
        static const int fusedEnv = [] { const char* e = getenv("TRQ_FUSED_RESOLVE"); return e ? atoi(e) : 1; }();
        fused = s->allTriangles && !gather && fusedEnv != 0;
        int perSM = s->blocksPerSM[any ? 1 : 0][fused ? 1 : 0];                  // queried once, in trq_scene_create
        if (blocksPerSMOverride > 0 && blocksPerSMOverride < perSM) perSM = blocksPerSMOverride;
        uint64_t grid = (uint64_t)perSM * (uint64_t)s->numSMs;             // persistent: a multiple of the SM count
        const uint64_t need = (n + TRQ_BLOCK - 1) / TRQ_BLOCK;
        if (grid > need) grid = need;
        // queue heads are zero at creation and re-zeroed by the resolve kernel that follows each trace on the stream
        unsigned long long* counter = s->d_counters + (s->counterNext.fetch_add(1) % kCounterRing);
        usedCounter = counter;
        if (fused) TRQ_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));   // no resolve kernel to re-zero it
        TraceParams P;
        P.rays = d_rays; P.hits = d_hits; P.n = n; P.counter = counter;
        P.stackDepth = s->stackDepth; P.refillMin = refillMin; P.leafBatch = leafBatch;
        P.order = nullptr; P.nPtr = nPtr;
        if (prof) TRQ_CUDA(cudaEventRecord(prof[0], st));      // the ordering pass is part of the timed traversal
        // TRQ_SORT_RAYS: counting sort of ray indices by (origin cell, direction octant); stream-ordered scratch
        // strictly opt-in (the caller knows whether its batch is incoherent, e.g. bounce depth >= 1 in a scene
        // that does not fit L2); TRQ_SORT_RAYS=0/1 in the environment overrides for experiments.
        static const int sortEnv = [] { const char* e = getenv("TRQ_SORT_RAYS"); return e ? atoi(e) : -1; }();
        const bool sortRays = sortEnv >= 0 ? (sortEnv != 0) : ((flags & TRQ_SORT_RAYS) != 0);
        uint32_t* scratch = nullptr;
        if (sortRays && n >= 65536) {
            const size_t words = (size_t)TRQ_SORT_BINS + 2 * (size_t)n;          // hist | keys | order
            TRQ_CUDA(cudaMallocFromPoolAsync((void**)&scratch, words * sizeof(uint32_t), s->scratchPool, st));
            uint32_t* hist = scratch; uint32_t* keys = scratch + TRQ_SORT_BINS; uint32_t* order = keys + n;
            TRQ_CUDA(cudaMemsetAsync(hist, 0, (size_t)TRQ_SORT_BINS * sizeof(uint32_t), st));
            const unsigned gb = (unsigned)((n + 255) / 256);
            sort_count_kernel<<<gb, 256, 0, st>>>(s->dev, d_rays, n, nPtr, keys, hist);
            sort_scan_kernel<<<1, 1024, 0, st>>>(hist);
            sort_scatter_kernel<<<gb, 256, 0, st>>>(keys, n, nPtr, hist, order);
            g_launches += 3;
            P.order = order;
        }
        if (any) { if (fused) trace_packed_kernel<true, true><<<(unsigned)grid, TRQ_BLOCK, smem, st>>>(s->dev, P);
                   else       trace_packed_kernel<true, false><<<(unsigned)grid, TRQ_BLOCK, smem, st>>>(s->dev, P); }
        else     { if (fused) trace_packed_kernel<false, true><<<(unsigned)grid, TRQ_BLOCK, smem, st>>>(s->dev, P);
                   else       trace_packed_kernel<false, false><<<(unsigned)grid, TRQ_BLOCK, smem, st>>>(s->dev, P); }
        g_launches++;
        if (scratch) TRQ_CUDA(cudaFreeAsync(scratch, st));
    }
    if (prof) TRQ_CUDA(cudaEventRecord(prof[1], st));
    if (!fused) {
        const unsigned block = 256;
        const uint64_t grid = n ? (n + block - 1) / block : 1;
        if (gather) resolve_hits_kernel<true><<<(unsigned)grid, block, 0, st>>>(s->dev, d_rays, d_hits, n, nPtr, usedCounter, *gather);
        else        resolve_hits_kernel<false><<<(unsigned)grid, block, 0, st>>>(s->dev, d_rays, d_hits, n, nPtr, usedCounter, GatherDev{});
        g_launches++;
    }
    if (prof) TRQ_CUDA(cudaEventRecord(prof[2], st));
    TRQ_CUDA(cudaGetLastError());
    return TRQ_OK;
}

#ifdef TRQ_STAGE_TIMELINE
// developer instrumentation (make EXTRA=-DTRQ_STAGE_TIMELINE): device timestamps around every staged operation
std::vector<cudaEvent_t> g_tl;
cudaEvent_t* timeline_events(int k) {
    const size_t at = g_tl.size();
    g_tl.resize(at + k);
    for (int i = 0; i < k; ++i) cudaEventCreate(&g_tl[at + i]);
    return &g_tl[at];
}
void timeline_dump() {
    if (getenv("TRQ_STAGE_TIMELINE_PRINT"))
        for (size_t i = 0; i + 3 < g_tl.size(); i += 4) {
            float a, b, c, d;
            cudaEventElapsedTime(&a, g_tl[0], g_tl[i]); cudaEventElapsedTime(&b, g_tl[0], g_tl[i + 1]);
            cudaEventElapsedTime(&c, g_tl[0], g_tl[i + 2]); cudaEventElapsedTime(&d, g_tl[0], g_tl[i + 3]);
            fprintf(stderr, "chunk %3zu  h2d %8.1f..%8.1f  trace ..%8.1f  d2h ..%8.1f us\n", i / 4, a * 1e3, b * 1e3, c * 1e3, d * 1e3);
        }
    for (cudaEvent_t e : g_tl) cudaEventDestroy(e);
    g_tl.clear();
}
#endif

int sync_staging(trq_scene* s) {
    for (int b = 0; b < kStageBufs; ++b)
        if (s->stageStream[b]) TRQ_CUDA(cudaStreamSynchronize(s->stageStream[b]));
    return TRQ_OK;
}

int ensure_staging(trq_scene* s, uint64_t chunk) {
    if (!s->stageReady) {
        for (int b = 0; b < kStageBufs; ++b) TRQ_CUDA(cudaStreamCreateWithFlags(&s->stageStream[b], cudaStreamNonBlocking));
        s->stageReady = true;
    }
    if (chunk > s->stageCap) {
        int rc = sync_staging(s);                             // asynchronous calls may still be using the old buffers
        if (rc != TRQ_OK) return rc;
        for (int b = 0; b < kStageBufs; ++b) {
            cudaFree(s->d_stageRays[b]); cudaFree(s->d_stageHits[b]);
            s->d_stageRays[b] = nullptr; s->d_stageHits[b] = nullptr;
        }
        s->stageCap = 0;
        for (int b = 0; b < kStageBufs; ++b) {
            TRQ_CUDA(cudaMalloc((void**)&s->d_stageRays[b], chunk * sizeof(trq_ray)));
            TRQ_CUDA(cudaMalloc((void**)&s->d_stageHits[b], chunk * sizeof(trq_hit)));
        }
        s->stageCap = chunk;
    }
    return TRQ_OK;
}

// Host-pointer path: the batch is cut into chunks; chunk k goes through staging buffer k % kStageBufs on that
// buffer's own stream (copy in, trace, copy out, in stream order). Different chunks overlap on the two copy engines
// and the SMs, buffer reuse is ordered by the stream itself, and a chunk costs four driver calls.
int trace_host(trq_scene* s, const trq_ray* rays, uint64_t n, uint32_t flags, trq_hit* hits) {
    const uint64_t chunkRays = [] {                           // read per call so that one process can sweep it
        const char* e = getenv("TRQ_CHUNK_RAYS");
        long long v = e ? atoll(e) : (TRQ_DEFAULT_CHUNK_RAYS)  /* B200 sweep: profiles/r01_e2e_chunk_sweep.txt */;
        return (uint64_t)(v < 1024 ? 1024 : v);
    }();
    std::lock_guard<std::mutex> lock(s->stageMutex);
    // a sorted chunk is only as coherent as it is large: C5 e2e 461 / 605 / 599 / 504 Mrays/s at 512K / 1M / 2M / 4M rays
    const uint64_t want = (flags & TRQ_SORT_RAYS) ? 2 * chunkRays : chunkRays;
    const uint64_t chunk = n < want ? n : want;
    int rc = ensure_staging(s, chunk);
    if (rc != TRQ_OK) return rc;
    uint64_t done = 0;
    while (done < n) {
        const uint64_t m = (n - done) < chunk ? (n - done) : chunk;
        const int b = (int)(s->stageSeq++ % kStageBufs);       // runs across calls: TRQ_HOST_ASYNC calls share the ring
        cudaStream_t st = s->stageStream[b];
#ifdef TRQ_STAGE_TIMELINE
        cudaEvent_t* tl = timeline_events(4); cudaEventRecord(tl[0], st);
#endif
        TRQ_CUDA(cudaMemcpyAsync(s->d_stageRays[b], rays + done, m * sizeof(trq_ray), cudaMemcpyHostToDevice, st));
#ifdef TRQ_STAGE_TIMELINE
        cudaEventRecord(tl[1], st);
#endif
        rc = launch_trace(s, s->d_stageRays[b], m, flags, s->d_stageHits[b], st);
        if (rc != TRQ_OK) return rc;
#ifdef TRQ_STAGE_TIMELINE
        cudaEventRecord(tl[2], st);
#endif
        TRQ_CUDA(cudaMemcpyAsync(hits + done, s->d_stageHits[b], m * sizeof(trq_hit), cudaMemcpyDeviceToHost, st));
#ifdef TRQ_STAGE_TIMELINE
        cudaEventRecord(tl[3], st);
#endif
        done += m;
    }
#ifdef TRQ_STAGE_TIMELINE
    sync_staging(s); timeline_dump();
#endif
    if (flags & TRQ_HOST_ASYNC) return TRQ_OK;               // the caller collects with trq_host_sync()
    return sync_staging(s);
}

}  // namespace

extern "C" {

int trq_version(void) { return TRQ_VERSION; }

const char* trq_last_error_string(void) { return trq::last_error(); }

int trq_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

uint64_t trq_launch_count(void) { return g_launches.load(); }

#ifdef TRQ_STATS
void trq_debug_stats(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, trq::g_stats, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(trq::g_stats, z, sizeof z); }
}
#endif

int trq_scene_create(const trq_scene_desc* d, int device, trq_scene** out) {
    if (!d || !out) return trq::fail(TRQ_ERR_INVALID, "trq_scene_create: NULL argument");
    *out = nullptr;
    if (!d->bvhList || d->nNode == 0) return trq::fail(TRQ_ERR_INVALID, "trq_scene_create: empty bvhList");
    if ((d->nSphere && !d->sphereList) || (d->nSquare && !d->squareList) || (d->nCube && !d->cubeList) ||
        (d->nVert && !d->triList) || (d->nTri && !d->idxList))
        return trq::fail(TRQ_ERR_INVALID, "trq_scene_create: NULL array with non-zero count");

    trq_scene_info_t info{};
    std::vector<uint32_t> ref;
    int rc = plan_layout(d, ref, info);                        // pure host validation: runs without a GPU
    if (rc != TRQ_OK) return rc;

    int ndev = trq_device_count();
    if (ndev <= 0) return trq::fail(TRQ_ERR_NO_DEVICE, "trq_scene_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return trq::fail(TRQ_ERR_INVALID, "trq_scene_create: device %d out of range (%d devices)", device, ndev);

    DeviceGuard guard(device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", device);

    trq_scene* s = new (std::nothrow) trq_scene();
    if (!s) return trq::fail(TRQ_ERR_NOMEM, "trq_scene_create: out of host memory");
    s->device = device;
    auto bail = [&](int code) { free_scene(s); return code; };

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(trq::fail(TRQ_ERR_CUDA, "cudaGetDeviceProperties failed"));
    s->numSMs = prop.multiProcessorCount;

    if ((rc = upload(&s->d_spheres, d->sphereList, d->nSphere)) != TRQ_OK) return bail(rc);
    if ((rc = upload(&s->d_squares, d->squareList, d->nSquare)) != TRQ_OK) return bail(rc);
    if ((rc = upload(&s->d_cubes, d->cubeList, d->nCube)) != TRQ_OK) return bail(rc);
    if ((rc = upload(&s->d_verts, d->triList, d->nVert)) != TRQ_OK) return bail(rc);
    if ((rc = upload(&s->d_idx, d->idxList, (size_t)d->nTri * 3)) != TRQ_OK) return bail(rc);
    if ((rc = upload(&s->d_bvh, d->bvhList, d->nNode)) != TRQ_OK) return bail(rc);

    uint32_t* d_ref = nullptr;
    if ((rc = upload(&d_ref, ref.data(), ref.size())) != TRQ_OK) return bail(rc);
    auto bail2 = [&](int code) { cudaFree(d_ref); return bail(code); };

    const size_t nodeBytes = (size_t)info.nInterior * 64, triBytes = (size_t)info.nTri * 16 * TRQ_TRI_STRIDE, sphBytes = (size_t)info.nSphere * 32;
    if (nodeBytes && cudaMalloc((void**)&s->d_nodes, nodeBytes) != cudaSuccess) return bail2(trq::fail(TRQ_ERR_CUDA, "cudaMalloc(nodes %zu B) failed", nodeBytes));
    if (triBytes && cudaMalloc((void**)&s->d_tris, triBytes) != cudaSuccess) return bail2(trq::fail(TRQ_ERR_CUDA, "cudaMalloc(tris %zu B) failed", triBytes));
    if (sphBytes && cudaMalloc((void**)&s->d_sph, sphBytes) != cudaSuccess) return bail2(trq::fail(TRQ_ERR_CUDA, "cudaMalloc(spheres %zu B) failed", sphBytes));
    const size_t triNBytes = (size_t)info.nTri * 64;
    if (triNBytes && cudaMalloc((void**)&s->d_triN, triNBytes) != cudaSuccess) return bail2(trq::fail(TRQ_ERR_CUDA, "cudaMalloc(triangle normals %zu B) failed", triNBytes));
    if (cudaMalloc((void**)&s->d_counters, kCounterRing * sizeof(unsigned long long)) != cudaSuccess) return bail2(trq::fail(TRQ_ERR_CUDA, "cudaMalloc(counters) failed"));
    if (cudaMemset(s->d_counters, 0, kCounterRing * sizeof(unsigned long long)) != cudaSuccess) return bail2(trq::fail(TRQ_ERR_CUDA, "cudaMemset(counters) failed"));

    {
        const unsigned block = 256, grid = (d->nNode + block - 1) / block;
        pack_scene_kernel<<<grid, block>>>(s->d_bvh, d_ref, d->nNode, s->d_verts, s->d_idx, s->d_spheres,
                                           s->d_nodes, s->d_tris, s->d_sph, s->d_triN);
        g_launches++;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return bail2(trq::fail(TRQ_ERR_CUDA, "pack_scene_kernel failed: %s", cudaGetErrorString(e)));
    }
    cudaFree(d_ref);

    const RefBVH* N = (const RefBVH*)d->bvhList;
    s->dev.spheres = s->d_spheres; s->dev.squares = s->d_squares; s->dev.cubes = s->d_cubes;
    s->dev.verts = s->d_verts; s->dev.idx = s->d_idx; s->dev.bvh = s->d_bvh;
    s->dev.nodes = s->d_nodes; s->dev.tris = s->d_tris; s->dev.sph = s->d_sph; s->dev.triN = s->d_triN;
    s->allTriangles = info.nTri > 0 && info.nTri == info.nLeaf;
    s->dev.rootRef = ref[0];
    for (int k = 0; k < 3; ++k) { s->dev.rootMin[k] = N[0].bBOX.mini[k]; s->dev.rootMax[k] = N[0].bBOX.maxi[k]; }
    s->dev.nNode = d->nNode;
    s->stackDepth = info.maxDepth + 1;
    {
        cudaMemPoolProps pp = {};
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&s->scratchPool, &pp);
        unsigned long long keep = ~0ull;
        if (e == cudaSuccess) e = cudaMemPoolSetAttribute(s->scratchPool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return bail(trq::fail(TRQ_ERR_CUDA, "cudaMemPoolCreate failed: %s", cudaGetErrorString(e)));
    }
    s->traceSmem = ((size_t)s->stackDepth + COLD_WORDS) * TRQ_BLOCK * sizeof(uint32_t);
    {
        cudaError_t e = cudaSuccess;
        const void* variants[2][2] = {{(const void*)trace_packed_kernel<false, false>, (const void*)trace_packed_kernel<false, true>},
                                      {(const void*)trace_packed_kernel<true, false>, (const void*)trace_packed_kernel<true, true>}};
        for (int a = 0; a < 2; ++a)
            for (int f = 0; f < 2; ++f) {
                if (e == cudaSuccess && s->traceSmem > 48 * 1024)
                    e = cudaFuncSetAttribute(variants[a][f], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->traceSmem);
                if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->blocksPerSM[a][f], variants[a][f], TRQ_BLOCK, s->traceSmem);
            }
        if (e != cudaSuccess) return bail(trq::fail(TRQ_ERR_CUDA, "trace_packed_kernel occupancy query failed: %s", cudaGetErrorString(e)));
        if (s->blocksPerSM[0][0] < 1 || s->blocksPerSM[0][1] < 1 || s->blocksPerSM[1][0] < 1 || s->blocksPerSM[1][1] < 1)
            return bail(trq::fail(TRQ_ERR_CUDA, "trace_packed_kernel does not fit on an SM (smem %zu)", s->traceSmem));
    }

    info.bytesReferenceLayout = (uint64_t)d->nSphere * sizeof(RefSphere) + (uint64_t)d->nSquare * sizeof(RefSquare) +
                                (uint64_t)d->nCube * sizeof(RefCube) + (uint64_t)d->nVert * sizeof(RefVertex) +
                                (uint64_t)d->nTri * 12 + (uint64_t)d->nNode * sizeof(RefBVH);
    info.bytesPacked = nodeBytes + triBytes + sphBytes + triNBytes;
    info.device = device;
    s->info = info;
    *out = s;
    return TRQ_OK;
}

int trq_scene_destroy(trq_scene* s) {
    if (!s) return TRQ_OK;
    DeviceGuard guard(s->device);
    free_scene(s);
    return TRQ_OK;
}

int trq_scene_info(const trq_scene* s, trq_scene_info_t* info) {
    if (!s || !info) return trq::fail(TRQ_ERR_INVALID, "trq_scene_info: NULL argument");
    *info = s->info;
    return TRQ_OK;
}

int trq_profile_enable(trq_scene* s, int on) {
    if (!s) return trq::fail(TRQ_ERR_INVALID, "trq_profile_enable: NULL scene");
    DeviceGuard guard(s->device);
    if (on && !s->evProf[0][0]) {
        for (int k = 0; k < kProfRing; ++k)
            for (int j = 0; j < 3; ++j) TRQ_CUDA(cudaEventCreate(&s->evProf[k][j]));
    }
    s->profile = on != 0;
    s->profHead = 0; s->profCount = 0;
    return TRQ_OK;
}

int trq_profile_read(trq_scene* s, uint32_t* nLaunches, float* traceMs, float* resolveMs) {
    if (!s || !nLaunches || !traceMs || !resolveMs) return trq::fail(TRQ_ERR_INVALID, "trq_profile_read: NULL argument");
    DeviceGuard guard(s->device);
    *nLaunches = 0; *traceMs = 0.0f; *resolveMs = 0.0f;
    for (uint32_t k = 0; k < s->profCount; ++k) {
        cudaEvent_t* ev = s->evProf[(s->profHead - 1 - k) % kProfRing];
        TRQ_CUDA(cudaEventSynchronize(ev[2]));
        float a = 0.0f, b = 0.0f;
        TRQ_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
        TRQ_CUDA(cudaEventElapsedTime(&b, ev[1], ev[2]));
        *traceMs += a; *resolveMs += b; (*nLaunches)++;
    }
    s->profHead = 0; s->profCount = 0;
    return TRQ_OK;
}

int trq_trace(trq_scene* s, const trq_ray* rays, uint64_t n, uint32_t flags, trq_hit* hits, void* stream) {
    if (!s) return trq::fail(TRQ_ERR_INVALID, "trq_trace: NULL scene");
    if (n == 0) return TRQ_OK;
    if (!rays || !hits) return trq::fail(TRQ_ERR_INVALID, "trq_trace: NULL rays/hits");
    if (!(flags & TRQ_HOST_PTRS) && ((((uintptr_t)rays) | ((uintptr_t)hits)) & 31u))
        return trq::fail(TRQ_ERR_INVALID, "trq_trace: device rays/hits must be 32-byte aligned (one record = one 256-bit access)");
    DeviceGuard guard(s->device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    if (flags & TRQ_HOST_PTRS) return trace_host(s, rays, n, flags, hits);
    return launch_trace(s, rays, n, flags, hits, (cudaStream_t)stream);
}

int trq_host_sync(trq_scene* s) {
    if (!s) return trq::fail(TRQ_ERR_INVALID, "trq_host_sync: NULL scene");
    DeviceGuard guard(s->device);
    std::lock_guard<std::mutex> lock(s->stageMutex);
    return sync_staging(s);
}

int trq_expand_hits(trq_scene* s, const trq_ray* rays, const trq_hit* hits, uint64_t n, uint32_t flags,
                    trq_hit_record* records, void* stream) {
    if (!s) return trq::fail(TRQ_ERR_INVALID, "trq_expand_hits: NULL scene");
    if (n == 0) return TRQ_OK;
    if (!rays || !hits || !records) return trq::fail(TRQ_ERR_INVALID, "trq_expand_hits: NULL argument");
    DeviceGuard guard(s->device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    const unsigned block = 256;
    const uint64_t grid = (n + block - 1) / block;
    if (flags & TRQ_HOST_PTRS) {
        trq_ray* dr = nullptr; trq_hit* dh = nullptr; trq_hit_record* dq = nullptr;
        TRQ_CUDA(cudaMalloc((void**)&dr, n * sizeof(trq_ray)));
        if (cudaMalloc((void**)&dh, n * sizeof(trq_hit)) != cudaSuccess) { cudaFree(dr); return trq::fail(TRQ_ERR_CUDA, "cudaMalloc failed"); }
        if (cudaMalloc((void**)&dq, n * sizeof(trq_hit_record)) != cudaSuccess) { cudaFree(dr); cudaFree(dh); return trq::fail(TRQ_ERR_CUDA, "cudaMalloc failed"); }
        cudaError_t e = cudaMemcpy(dr, rays, n * sizeof(trq_ray), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dh, hits, n * sizeof(trq_hit), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) { expand_hits_kernel<<<(unsigned)grid, block>>>(s->dev, dr, dh, dq, n); g_launches++; e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaMemcpy(records, dq, n * sizeof(trq_hit_record), cudaMemcpyDeviceToHost);
        cudaFree(dr); cudaFree(dh); cudaFree(dq);
        if (e != cudaSuccess) return trq::fail(TRQ_ERR_CUDA, "trq_expand_hits: %s", cudaGetErrorString(e));
        return TRQ_OK;
    }
    expand_hits_kernel<<<(unsigned)grid, block, 0, (cudaStream_t)stream>>>(s->dev, rays, hits, records, n);
    g_launches++;
    TRQ_CUDA(cudaGetLastError());
    return TRQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Wavefront callers: device-side ray producers + a trace whose batch size lives
// on the device, so that cast -> trace -> spawn -> trace -> ... runs without a host round trip.
#include "../../include/tracer_rq_harness.h"
#include "kernels/producers.cuh"

extern "C" {

int trq_trace_indirect(trq_scene* s, const trq_ray* rays, const uint64_t* d_count, uint64_t capacity, uint32_t flags,
                       trq_hit* hits, void* stream) {
    if (!s) return trq::fail(TRQ_ERR_INVALID, "trq_trace_indirect: NULL scene");
    if (flags & TRQ_HOST_PTRS) return trq::fail(TRQ_ERR_INVALID, "trq_trace_indirect: device pointers only");
    if (capacity == 0) return TRQ_OK;
    if (!rays || !hits || !d_count) return trq::fail(TRQ_ERR_INVALID, "trq_trace_indirect: NULL argument");
    if ((((uintptr_t)rays) | ((uintptr_t)hits)) & 31u)
        return trq::fail(TRQ_ERR_INVALID, "trq_trace_indirect: rays/hits must be 32-byte aligned");
    DeviceGuard guard(s->device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    return launch_trace(s, rays, capacity, flags, hits, (cudaStream_t)stream, (const unsigned long long*)d_count);
}

int trq_cast_rays(trq_scene* s, const trq_camera* cam, uint32_t W, uint32_t H, trq_ray* rays, void* stream) {
    if (!s || !cam || !rays) return trq::fail(TRQ_ERR_INVALID, "trq_cast_rays: NULL argument");
    if (cam->aperture != 0.0f) return trq::fail(TRQ_ERR_INVALID, "trq_cast_rays: only aperture 0 (the reference's setting, Tracer.mm:380) is supported");
    if ((uint64_t)W * H == 0) return TRQ_OK;
    DeviceGuard guard(s->device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    float c[18];
    trqh_make_camera(cam->lookFrom, cam->lookAt, cam->viewUp, cam->vfov, cam->aspect, cam->focus_dist, c);   // MakeCamera, host
    CameraDev cd;
    for (int k = 0; k < 3; ++k) {
        cd.lookFrom[k] = c[k]; cd.u[k] = c[3 + k]; cd.v[k] = c[6 + k];
        cd.vertical[k] = c[9 + k]; cd.horizontal[k] = c[12 + k]; cd.corner[k] = c[15 + k];
    }
    cd.lenRadius = cam->aperture / 2;
    const uint64_t n = (uint64_t)W * H;
    cast_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cd, W, H, rays);
    g_launches++;
    TRQ_CUDA(cudaGetLastError());
    return TRQ_OK;
}

static int spawn_common(trq_scene* s, const trq_ray* rays, const trq_hit* hits, uint64_t n, const uint64_t* d_n,
                        uint64_t* d_count, const char* who) {
    if (!s) return trq::fail(TRQ_ERR_INVALID, "%s: NULL scene", who);
    if (!d_count) return trq::fail(TRQ_ERR_INVALID, "%s: NULL d_count", who);
    if (n && (!rays || !hits)) return trq::fail(TRQ_ERR_INVALID, "%s: NULL rays/hits", who);
    if (n > (1ull << 32)) return trq::fail(TRQ_ERR_INVALID, "%s: more than 2^32 rays in one call", who);

    return TRQ_OK;
}

int trq_spawn_bounce(trq_scene* s, const trq_ray* rays, const trq_hit* hits, uint64_t n, const uint64_t* d_n, uint64_t seedBase,
                     trq_ray* out, uint32_t* srcIndex, uint64_t* d_count, void* stream) {
    return trq_spawn_bounce_rng(s, rays, hits, n, d_n, seedBase, nullptr, nullptr, out, srcIndex, d_count, stream);
}

int trq_spawn_bounce_rng(trq_scene* s, const trq_ray* rays, const trq_hit* hits, uint64_t n, const uint64_t* d_n, uint64_t seedBase,
                         const uint32_t* pixelOf, uint32_t* rngState, trq_ray* out, uint32_t* srcIndex, uint64_t* d_count, void* stream) {
    int rc = spawn_common(s, rays, hits, n, d_n, d_count, "trq_spawn_bounce");
    if (rc == TRQ_OK && rngState && (((uintptr_t)rngState) & 15u)) rc = trq::fail(TRQ_ERR_INVALID, "trq_spawn_bounce_rng: rngState must be 16-byte aligned");
    if (rc != TRQ_OK) return rc;
    if (n && !out) return trq::fail(TRQ_ERR_INVALID, "trq_spawn_bounce: NULL out");
    DeviceGuard guard(s->device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    cudaStream_t st = (cudaStream_t)stream;
    TRQ_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
    if (n == 0) return TRQ_OK;
    spawn_bounce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->dev, rays, hits, n, (const unsigned long long*)d_n, seedBase, pixelOf, rngState, out, srcIndex,
                                                                      (unsigned long long*)d_count);

    g_launches++;
    TRQ_CUDA(cudaGetLastError());
    return TRQ_OK;
}

int trq_rng_frame_begin(trq_scene* s, uint32_t* rngState, uint64_t nPixels, void* stream) {
    if (!s) return trq::fail(TRQ_ERR_INVALID, "trq_rng_frame_begin: NULL scene");
    if (nPixels == 0) return TRQ_OK;
    if (!rngState || (((uintptr_t)rngState) & 15u)) return trq::fail(TRQ_ERR_INVALID, "trq_rng_frame_begin: rngState must be a 16-byte aligned device pointer");
    DeviceGuard guard(s->device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    rng_frame_begin_kernel<<<(unsigned)((nPixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>((uint4*)rngState, nPixels);
    g_launches++;
    TRQ_CUDA(cudaGetLastError());
    return TRQ_OK;
}

int trq_spawn_shadow(trq_scene* s, const trq_ray* rays, const trq_hit* hits, uint64_t n, const uint64_t* d_n, uint64_t seedBase,
                     uint32_t lightA, uint32_t lightB, trq_ray* out, uint32_t* srcIndex, uint64_t* d_count, void* stream) {
    return trq_spawn_shadow_rng(s, rays, hits, n, d_n, seedBase, nullptr, nullptr, lightA, lightB, out, srcIndex, d_count, stream);
}

int trq_spawn_shadow_rng(trq_scene* s, const trq_ray* rays, const trq_hit* hits, uint64_t n, const uint64_t* d_n, uint64_t seedBase,
                         const uint32_t* pixelOf, uint32_t* rngState, uint32_t lightA, uint32_t lightB, trq_ray* out, uint32_t* srcIndex,
                         uint64_t* d_count, void* stream) {
    int rc = spawn_common(s, rays, hits, n, d_n, d_count, "trq_spawn_shadow");
    if (rc == TRQ_OK && rngState && (((uintptr_t)rngState) & 15u)) rc = trq::fail(TRQ_ERR_INVALID, "trq_spawn_shadow_rng: rngState must be 16-byte aligned");
    if (rc != TRQ_OK) return rc;
    if (n && !out) return trq::fail(TRQ_ERR_INVALID, "trq_spawn_shadow: NULL out");
    if (lightA >= s->info.nSquare || lightB >= s->info.nSquare)
        return trq::fail(TRQ_ERR_INVALID, "trq_spawn_shadow: light index out of range (nSquare %u)", s->info.nSquare);
    DeviceGuard guard(s->device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    cudaStream_t st = (cudaStream_t)stream;
    TRQ_CUDA(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
    if (n == 0) return TRQ_OK;
    spawn_shadow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->dev, rays, hits, n, (const unsigned long long*)d_n, seedBase, pixelOf, rngState, lightA, lightB, out, srcIndex,
                                                                      (unsigned long long*)d_count);
    g_launches++;
    TRQ_CUDA(cudaGetLastError());
    return TRQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Peer-memory hit gather (a consumer that wants every rank's hits whole). One process per GPU;
// every rank owns a buffer [parity][rank][capacity] of trq_hit plus a small header (flags, counts), exported through
// CUDA IPC. trq_trace_gather traces this rank's rays and its resolve kernel stores each finished record into slot
// [rank] of EVERY rank's buffer over NVLink -- the all-gather is fused into the kernel that produces the records, no
// NCCL call and no second pass over the data -- then publishes (count, step) with system-scope release stores.
// trq_gather_wait enqueues a small kernel that acquires all ranks' step numbers. Two parities: a rank can run one
// step ahead of a peer that is still consuming the previous one, never two (it would need the peer's next flag).
namespace {
constexpr size_t kGatherHeader = 4096;                 // flags[16] u64 @0, counts[2][16] u64 @128, blocksDone u32 @2048
constexpr size_t kGatherCountsAt = 128, kGatherBlocksDoneAt = 2048;
}

struct trq_gather {
    trq_scene* scene = nullptr;
    uint32_t rank = 0, world = 1;
    uint64_t capacity = 0;
    uint8_t* base = nullptr;
    uint8_t* peerBase[TRQ_GATHER_MAX_RANKS] = {};
    unsigned int* h_status = nullptr;                  // pinned + mapped: raised by gather_wait_kernel on timeout
    unsigned int* d_status = nullptr;
    unsigned long long step = 0;
    bool connected = false;
    size_t slot_offset(unsigned parity, uint32_t r) const {
        return kGatherHeader + ((size_t)parity * world + r) * capacity * sizeof(trq_hit);
    }
};

extern "C" {

int trq_gather_create(trq_scene* s, uint32_t rank, uint32_t world, uint64_t capacity, trq_gather** out, void* handle) {
    if (!s || !out || !handle) return trq::fail(TRQ_ERR_INVALID, "trq_gather_create: NULL argument");
    *out = nullptr;
    if (world == 0 || world > TRQ_GATHER_MAX_RANKS || rank >= world)
        return trq::fail(TRQ_ERR_INVALID, "trq_gather_create: rank %u / world %u (at most %d ranks)", rank, world, TRQ_GATHER_MAX_RANKS);
    if (capacity == 0 || capacity > (1ull << 31)) return trq::fail(TRQ_ERR_INVALID, "trq_gather_create: capacity out of range");
    static_assert(sizeof(cudaIpcMemHandle_t) == TRQ_GATHER_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(s->device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    trq_gather* g = new (std::nothrow) trq_gather();
    if (!g) return trq::fail(TRQ_ERR_NOMEM, "trq_gather_create: out of host memory");
    g->scene = s; g->rank = rank; g->world = world; g->capacity = capacity;
    const size_t total = kGatherHeader + 2 * (size_t)world * capacity * sizeof(trq_hit);
    cudaError_t e = cudaMalloc((void**)&g->base, total);
    if (e == cudaSuccess) e = cudaMemset(g->base, 0, kGatherHeader);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&g->h_status, sizeof(unsigned int), cudaHostAllocMapped);
    if (e == cudaSuccess) { *g->h_status = 0; e = cudaHostGetDevicePointer((void**)&g->d_status, g->h_status, 0); }
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, g->base);
    if (e != cudaSuccess) {
        cudaFree(g->base); if (g->h_status) cudaFreeHost(g->h_status); delete g;
        return trq::fail(TRQ_ERR_CUDA, "trq_gather_create (%zu bytes): %s", total, cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof h);
    *out = g;
    return TRQ_OK;
}

int trq_gather_connect(trq_gather* g, const void* handles) {
    if (!g || (!handles && g->world > 1)) return trq::fail(TRQ_ERR_INVALID, "trq_gather_connect: NULL argument");
    if (g->connected) return trq::fail(TRQ_ERR_INVALID, "trq_gather_connect: already connected");
    DeviceGuard guard(g->scene->device);
    for (uint32_t r = 0; r < g->world; ++r) {
        if (r == g->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles + (size_t)r * sizeof h, sizeof h);
        cudaError_t e = cudaIpcOpenMemHandle((void**)&g->peerBase[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return trq::fail(TRQ_ERR_CUDA, "trq_gather_connect: cudaIpcOpenMemHandle(rank %u) failed: %s (one process per GPU, same node)",
                             r, cudaGetErrorString(e));
    }
    g->connected = true;
    return TRQ_OK;
}

int trq_trace_gather(trq_scene* s, trq_gather* g, const trq_ray* rays, uint64_t n, uint32_t flags, void* stream) {
    if (!s || !g || g->scene != s) return trq::fail(TRQ_ERR_INVALID, "trq_trace_gather: NULL or foreign scene / gather");
    if (!g->connected) return trq::fail(TRQ_ERR_INVALID, "trq_trace_gather: call trq_gather_connect first");
    if (flags & (TRQ_HOST_PTRS | TRQ_HOST_ASYNC)) return trq::fail(TRQ_ERR_INVALID, "trq_trace_gather: device pointers only");
    if (n > g->capacity) return trq::fail(TRQ_ERR_INVALID, "trq_trace_gather: %llu rays exceed the capacity %llu", (unsigned long long)n, (unsigned long long)g->capacity);
    if (n && !rays) return trq::fail(TRQ_ERR_INVALID, "trq_trace_gather: NULL rays");
    if (((uintptr_t)rays) & 31u) return trq::fail(TRQ_ERR_INVALID, "trq_trace_gather: rays must be 32-byte aligned");
    DeviceGuard guard(s->device);
    if (!guard.ok) return trq::fail(TRQ_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
    const unsigned long long step = ++g->step;
    const unsigned parity = (unsigned)(step & 1ull);
    GatherDev G{};
    G.step = step;
    G.blocksDone = (unsigned int*)(g->base + kGatherBlocksDoneAt);
    G.ownFlag = (unsigned long long*)g->base + g->rank;
    G.ownCount = (unsigned long long*)(g->base + kGatherCountsAt) + parity * TRQ_GATHER_MAX_RANKS + g->rank;
    for (uint32_t r = 0; r < g->world; ++r) {
        if (r == g->rank) continue;
        const uint32_t k = G.nPeer++;
        G.peerSlot[k] = (trq_hit*)(g->peerBase[r] + g->slot_offset(parity, g->rank));
        G.peerFlag[k] = (unsigned long long*)g->peerBase[r] + g->rank;
        G.peerCount[k] = (unsigned long long*)(g->peerBase[r] + kGatherCountsAt) + parity * TRQ_GATHER_MAX_RANKS + g->rank;
    }
    trq_hit* own = (trq_hit*)(g->base + g->slot_offset(parity, g->rank));
    return launch_trace(s, rays, n, flags, own, (cudaStream_t)stream, nullptr, &G);
}

int trq_gather_wait(trq_gather* g, void* stream, const trq_hit** hitsAll, const uint64_t** counts) {
    if (!g) return trq::fail(TRQ_ERR_INVALID, "trq_gather_wait: NULL gather");
    if (g->step == 0) return trq::fail(TRQ_ERR_INVALID, "trq_gather_wait: nothing traced yet");
    DeviceGuard guard(g->scene->device);
    static const unsigned long long timeoutNs = [] {
        const char* e = getenv("TRQ_GATHER_TIMEOUT_MS");
        return (unsigned long long)(e ? atoll(e) : 10000) * 1000000ull;
    }();
    gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)g->base, g->world, g->step, timeoutNs, g->d_status);
    g_launches++;
    TRQ_CUDA(cudaGetLastError());
    const unsigned parity = (unsigned)(g->step & 1ull);
    if (hitsAll) *hitsAll = (const trq_hit*)(g->base + g->slot_offset(parity, 0));
    if (counts) *counts = (const uint64_t*)(g->base + kGatherCountsAt) + parity * TRQ_GATHER_MAX_RANKS;
    return TRQ_OK;
}

int trq_gather_status(trq_gather* g) {
    if (!g) return trq::fail(TRQ_ERR_INVALID, "trq_gather_status: NULL gather");
    const unsigned int v = *(volatile unsigned int*)g->h_status;
    if (v) return trq::fail(TRQ_ERR_CUDA, "trq_gather_wait: rank %u did not publish its hits before the timeout", v - 1);
    return TRQ_OK;
}

int trq_gather_destroy(trq_gather* g) {
    if (!g) return TRQ_OK;
    DeviceGuard guard(g->scene->device);
    cudaDeviceSynchronize();
    for (uint32_t r = 0; r < g->world; ++r) if (g->peerBase[r]) cudaIpcCloseMemHandle(g->peerBase[r]);
    cudaFree(g->base);
    if (g->h_status) cudaFreeHost(g->h_status);
    delete g;
    return TRQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Single-process multi-GPU helper (SURVEY.md section 8b "multi-GPU helper trq_mgpu_* owns one scene per rank", 8e): the
// reference's host is ONE process, so this is the form of ray sharding it can adopt without a launcher: one scene per
// device, host rays cut into contiguous ranges [k*n/R, (k+1)*n/R), every device's range staged and traced through its
// own copy engines and streams at once (TRQ_HOST_ASYNC on every scene, then one wait per scene). No collective anywhere:
// the scene is replicated by uploading it R times, the hits land in the caller's host array.
struct trq_mgpu {
    std::vector<trq_scene*> scenes;
};

extern "C" {

int trq_mgpu_create(const trq_scene_desc* desc, const int* devices, int nDevices, trq_mgpu** out) {
    if (!desc || !out) return trq::fail(TRQ_ERR_INVALID, "trq_mgpu_create: NULL argument");
    *out = nullptr;
    const int have = trq_device_count();
    if (have <= 0) return trq::fail(TRQ_ERR_NO_DEVICE, "trq_mgpu_create: no CUDA device (there is no CPU fallback)");
    if (nDevices <= 0) nDevices = have;                        // all visible devices
    if (nDevices > have && !devices) return trq::fail(TRQ_ERR_INVALID, "trq_mgpu_create: %d devices requested, %d visible", nDevices, have);
    trq_mgpu* m = new (std::nothrow) trq_mgpu();
    if (!m) return trq::fail(TRQ_ERR_NOMEM, "trq_mgpu_create: out of host memory");
    for (int k = 0; k < nDevices; ++k) {
        trq_scene* s = nullptr;
        const int rc = trq_scene_create(desc, devices ? devices[k] : k, &s);
        if (rc != TRQ_OK) {
            for (trq_scene* t : m->scenes) trq_scene_destroy(t);
            delete m;
            return rc;
        }
        m->scenes.push_back(s);
    }
    *out = m;
    return TRQ_OK;
}

int trq_mgpu_device_count(const trq_mgpu* m) { return m ? (int)m->scenes.size() : 0; }

trq_scene* trq_mgpu_scene(trq_mgpu* m, int k) {
    if (!m || k < 0 || k >= (int)m->scenes.size()) { trq::fail(TRQ_ERR_INVALID, "trq_mgpu_scene: index out of range"); return nullptr; }
    return m->scenes[k];
}

int trq_mgpu_shard(const trq_mgpu* m, uint64_t n, int k, uint64_t* lo, uint64_t* hi) {
    if (!m || !lo || !hi || k < 0 || k >= (int)m->scenes.size()) return trq::fail(TRQ_ERR_INVALID, "trq_mgpu_shard: bad argument");
    const unsigned __int128 R = m->scenes.size();
    *lo = (uint64_t)((unsigned __int128)n * (unsigned)k / R);
    *hi = (uint64_t)((unsigned __int128)n * (unsigned)(k + 1) / R);
    return TRQ_OK;
}

int trq_mgpu_trace(trq_mgpu* m, const trq_ray* rays, uint64_t n, uint32_t flags, trq_hit* hits) {
    if (!m) return trq::fail(TRQ_ERR_INVALID, "trq_mgpu_trace: NULL handle");
    if (n == 0) return TRQ_OK;
    if (!rays || !hits) return trq::fail(TRQ_ERR_INVALID, "trq_mgpu_trace: NULL rays/hits");
    int first = TRQ_OK;
    for (int k = 0; k < (int)m->scenes.size(); ++k) {           // queue every device's range ...
        uint64_t lo, hi;
        trq_mgpu_shard(m, n, k, &lo, &hi);
        if (hi == lo) continue;
        const int rc = trq_trace(m->scenes[k], rays + lo, hi - lo, flags | TRQ_HOST_PTRS | TRQ_HOST_ASYNC, hits + lo, nullptr);
        if (rc != TRQ_OK && first == TRQ_OK) first = rc;
    }
    for (trq_scene* s : m->scenes) {                            // ... then wait for all of them
        const int rc = trq_host_sync(s);
        if (rc != TRQ_OK && first == TRQ_OK) first = rc;
    }
    return first;
}

int trq_mgpu_destroy(trq_mgpu* m) {
    if (!m) return TRQ_OK;
    for (trq_scene* s : m->scenes) trq_scene_destroy(s);
    delete m;
    return TRQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Memory-system probes for the roofline report ("no L2 figure is in MEASURED_PEAKS.json, so the
// harness must measure it"): read-only 16-byte loads that bypass L1 (ld.global.cg) over a working set that fits in L2
// (32 MB) or does not (2 GB), all SMs, best of five.
namespace {
__global__ void __launch_bounds__(256)
probe_read_kernel(const uint4* __restrict__ buf, uint64_t n16, int iters, uint32_t* sink) {
    uint32_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it)
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;          // keeps the loads alive
}
}  // namespace

extern "C" int trq_probe_bandwidth(int device, int which, double* gbs) {
    if (!gbs) return trq::fail(TRQ_ERR_INVALID, "trq_probe_bandwidth: NULL argument");
    int ndev = trq_device_count();
    if (ndev <= 0) return trq::fail(TRQ_ERR_NO_DEVICE, "trq_probe_bandwidth: no CUDA device");
    if (device < 0 || device >= ndev) return trq::fail(TRQ_ERR_INVALID, "trq_probe_bandwidth: device out of range");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    TRQ_CUDA(cudaGetDeviceProperties(&prop, device));
    const uint64_t bytes = which == 0 ? (32ull << 20) : (2ull << 30);
    const int iters = which == 0 ? 64 : 2;
    uint4* buf = nullptr; uint32_t* sink = nullptr;
    TRQ_CUDA(cudaMalloc((void**)&buf, bytes));
    if (cudaMalloc((void**)&sink, 4) != cudaSuccess) { cudaFree(buf); return trq::fail(TRQ_ERR_CUDA, "cudaMalloc failed"); }
    cudaMemset(buf, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned grid = (unsigned)prop.multiProcessorCount * 8u;
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        probe_read_kernel<<<grid, 256>>>(buf, bytes / 16, iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.0f; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
        g_launches++;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    cudaFree(buf); cudaFree(sink);
    if (e != cudaSuccess) return trq::fail(TRQ_ERR_CUDA, "trq_probe_bandwidth: %s", cudaGetErrorString(e));
    *gbs = (double)bytes * iters / (best * 1e-3) / 1e9;
    return TRQ_OK;
}
