// tools/hostlink_probe.cu -- microbenchmark: what the SMs can move over PCIe compared with the copy engines.
// Answers the question behind the "direct host path" experiment of trq_trace(TRQ_HOST_PTRS) (profiles/r02_e2e_direct_experiment.txt):
// can kernels that read rays from / write records to PINNED HOST memory keep both directions of the link as busy as
// cudaMemcpyAsync does?
//   ce_h2d / ce_d2h / ce_both     copy engines, one large copy per direction
//   zc_read  L lanes              every warp reads L x 32 contiguous bytes (LDG.256) per request from host memory and stores them
//                                 to device memory -- the ray fetch of a warp refill
//   tma_read                      one thread per CTA: cp.async.bulk host -> shared (12 KB, two stages) -> device
//   tma_write                     the same, device -> shared -> host (the gather sender with the host as its peer)
//   tma_both, zc+tma_write        two kernels at once on two streams
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hostlink_probe hostlink_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) zc_read(const uint4* __restrict__ host, uint4* __restrict__ dev, size_t nUnits32, int lanes) {
    // unit = 32 bytes = two uint4; a warp request covers `lanes` consecutive units
    const unsigned lane = threadIdx.x & 31u;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t base = warp * lanes; base < nUnits32; base += nWarps * lanes) {
        if (lane < (unsigned)lanes && base + lane < nUnits32) {
            const uint4* p = host + (base + lane) * 2;
            uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
            asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "l"(p));
            uint4* q = dev + (base + lane) * 2;
            q[0] = make_uint4(a0, a1, a2, a3); q[1] = make_uint4(b0, b1, b2, b3);
        }
    }
}

#define CHUNK 12288u
#define STAGES 2u
// src -> shared -> dst with TMA bulk copies, one elected thread per CTA, chunks strided over the grid
__global__ void __launch_bounds__(32) tma_copy(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, size_t bytes) {
    extern __shared__ __align__(128) unsigned char buf[];
    __shared__ __align__(8) unsigned long long bar[STAGES];
    if (threadIdx.x != 0) return;
    for (uint32_t k = 0; k < STAGES; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bar[k])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t nChunks = (bytes + CHUNK - 1) / CHUNK;
    uint32_t no = 0;
    for (size_t c = blockIdx.x; c < nChunks; c += gridDim.x, ++no) {
        const size_t off = c * CHUNK;
        const uint32_t len = (uint32_t)((bytes - off) < CHUNK ? (bytes - off) : CHUNK);
        const uint32_t stage = no % STAGES, phase = (no / STAGES) & 1u;
        const uint32_t sm = (uint32_t)__cvta_generic_to_shared(buf + stage * CHUNK), mb = (uint32_t)__cvta_generic_to_shared(&bar[stage]);
        asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(STAGES - 1) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(len) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(sm), "l"(src + off), "r"(len), "r"(mb) : "memory");
        uint32_t ready = 0;
        while (!ready)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ready) : "r"(mb), "r"(phase) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst + off), "r"(sm), "r"(len) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// stand-in for the trace of one chunk: streams the chunk from `in` to `out` `reps` times (reps sets how long it occupies the SMs)
__global__ void __launch_bounds__(256) chunk_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t nUnits, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nUnits; i += (size_t)gridDim.x * blockDim.x) {
            uint4 v = in[i]; v.x += (uint32_t)r; out[i] = v;
        }
}

int main(int argc, char** argv) {
    const size_t bytes = (argc > 1 ? (size_t)atol(argv[1]) : 197166528ull) & ~(size_t)31;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("# %s, %d SMs, %zu bytes per direction\n", prop.name, sms, bytes);
    unsigned char *hIn, *hOut, *dIn, *dOut, *dA, *dB;
    CK(cudaHostAlloc((void**)&hIn, bytes, cudaHostAllocMapped)); CK(cudaHostAlloc((void**)&hOut, bytes, cudaHostAllocMapped));
    for (size_t i = 0; i < bytes; i += 4096) hIn[i] = (unsigned char)i;
    CK(cudaHostGetDevicePointer((void**)&dIn, hIn, 0)); CK(cudaHostGetDevicePointer((void**)&dOut, hOut, 0));
    CK(cudaMalloc((void**)&dA, bytes)); CK(cudaMalloc((void**)&dB, bytes));
    CK(cudaMemset(dB, 1, bytes));
    cudaStream_t s0, s1; CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, f0, f1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
    const int smemT = STAGES * CHUNK;
    CK(cudaFuncSetAttribute(tma_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, smemT));
    auto report = [&](const char* name, float ms, double dirs) { printf("%-28s %7.3f ms  %6.1f GB/s per direction%s\n", name, ms, bytes / ms / 1e6, dirs > 1 ? " (both busy)" : ""); };
    auto timeit = [&](const char* name, auto&& a, auto&& b, bool two) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s0)); if (two) CK(cudaEventRecord(f0, s1));
            a(); if (two) b();
            CK(cudaEventRecord(e1, s0)); if (two) CK(cudaEventRecord(f1, s1));
            CK(cudaDeviceSynchronize());
            CK(cudaGetLastError());
            float m0, m1 = 0.f; CK(cudaEventElapsedTime(&m0, e0, e1)); if (two) CK(cudaEventElapsedTime(&m1, f0, f1));
            const float m = m0 > m1 ? m0 : m1;
            if (rep > 0 && m < best) best = m;
        }
        report(name, best, two ? 2 : 1);
    };
    auto none = [] {};
    timeit("ce_h2d", [&] { CK(cudaMemcpyAsync(dA, hIn, bytes, cudaMemcpyHostToDevice, s0)); }, none, false);
    timeit("ce_d2h", [&] { CK(cudaMemcpyAsync(hOut, dB, bytes, cudaMemcpyDeviceToHost, s0)); }, none, false);
    timeit("ce_both", [&] { CK(cudaMemcpyAsync(dA, hIn, bytes, cudaMemcpyHostToDevice, s0)); }, [&] { CK(cudaMemcpyAsync(hOut, dB, bytes, cudaMemcpyDeviceToHost, s1)); }, true);
    char name[64];
    for (int occ : {1, 2, 5, 8}) for (int lanes : {16, 32}) {
        snprintf(name, sizeof name, "zc_read %d lanes, %d CTA/SM", lanes, occ);
        timeit(name, [&] { zc_read<<<sms * occ, 256, 0, s0>>>((const uint4*)dIn, (uint4*)dA, bytes / 32, lanes); }, none, false);
    }
    for (int per : {1, 2, 4}) {
        snprintf(name, sizeof name, "tma_read %d CTA/SM", per);
        timeit(name, [&] { tma_copy<<<sms * per, 32, smemT, s0>>>(dIn, dA, bytes); }, none, false);
        snprintf(name, sizeof name, "tma_write %d CTA/SM", per);
        timeit(name, [&] { tma_copy<<<sms * per, 32, smemT, s0>>>(dB, dOut, bytes); }, none, false);
        snprintf(name, sizeof name, "tma_both %d CTA/SM", per);
        timeit(name, [&] { tma_copy<<<sms * per, 32, smemT, s0>>>(dIn, dA, bytes); }, [&] { tma_copy<<<sms * per, 32, smemT, s1>>>(dB, dOut, bytes); }, true);
    }
    timeit("zc_read32x5 + tma_write", [&] { zc_read<<<sms * 5, 256, 0, s0>>>((const uint4*)dIn, (uint4*)dA, bytes / 32, 32); }, [&] { tma_copy<<<sms, 32, smemT, s1>>>(dB, dOut, bytes); }, true);
    timeit("ce_h2d + tma_write", [&] { CK(cudaMemcpyAsync(dA, hIn, bytes, cudaMemcpyHostToDevice, s0)); }, [&] { tma_copy<<<sms, 32, smemT, s1>>>(dB, dOut, bytes); }, true);
    timeit("tma_read + ce_d2h", [&] { tma_copy<<<sms, 32, smemT, s0>>>(dIn, dA, bytes); }, [&] { CK(cudaMemcpyAsync(hOut, dB, bytes, cudaMemcpyDeviceToHost, s1)); }, true);
    // the chunk pipeline of trq_trace(TRQ_HOST_PTRS): chunk k on stream k % S: H2D, kernel, D2H in stream order
    cudaStream_t ps[8]; for (int k = 0; k < 8; ++k) CK(cudaStreamCreateWithFlags(&ps[k], cudaStreamNonBlocking));
    cudaEvent_t p0, p1; CK(cudaEventCreate(&p0)); CK(cudaEventCreate(&p1));
    cudaEvent_t done[8]; for (int k = 0; k < 8; ++k) CK(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
    for (int reps : {0, 8}) for (int S : {2, 3, 4, 8}) for (size_t chunkMB : {2, 4, 8, 16, 32}) {
        const size_t chunk = chunkMB << 20;
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaDeviceSynchronize());
            const auto t0 = std::chrono::steady_clock::now();
            int k = 0;
            for (size_t off = 0; off < bytes; off += chunk, ++k) {
                const size_t len = (bytes - off) < chunk ? (bytes - off) : chunk;
                cudaStream_t st = ps[k % S];
                CK(cudaMemcpyAsync(dA + off, hIn + off, len, cudaMemcpyHostToDevice, st));
                if (reps) chunk_kernel<<<sms * 4, 256, 0, st>>>((const uint4*)(dA + off), (uint4*)(dB + off), len / 16, reps);
                CK(cudaMemcpyAsync(hOut + off, dB + off, len, cudaMemcpyDeviceToHost, st));
            }
            CK(cudaDeviceSynchronize());
            const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (rep > 0 && ms < best) best = ms;
        }
        printf("pipeline %2zu MB chunks, %d streams, kernel x%d   %7.3f ms wall  %6.1f GB/s per direction\n", chunkMB, S, reps, best, bytes / best / 1e6);
    }
    // the same pipeline with chunk sizes ramping up at the start and down at the end (first trace starts sooner, last D2H is shorter)
    for (int reps : {0, 8}) for (int S : {4, 8}) for (size_t firstMB : {1, 2, 4}) {
        std::vector<size_t> sizes;
        { size_t left = bytes; size_t c = firstMB << 20; const size_t full = 16u << 20;
          std::vector<size_t> up; for (; c < full; c *= 2) up.push_back(c);
          size_t upSum = 0; for (size_t v : up) upSum += v;
          for (size_t v : up) sizes.push_back(v);
          left -= 2 * upSum;
          while (left > full) { sizes.push_back(full); left -= full; }
          if (left) sizes.push_back(left);
          for (size_t k = up.size(); k-- > 0;) sizes.push_back(up[k]); }
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaDeviceSynchronize());
            const auto t0 = std::chrono::steady_clock::now();
            size_t off = 0; int k = 0;
            for (size_t len : sizes) {
                cudaStream_t st = ps[k++ % S];
                CK(cudaMemcpyAsync(dA + off, hIn + off, len, cudaMemcpyHostToDevice, st));
                if (reps) chunk_kernel<<<sms * 4, 256, 0, st>>>((const uint4*)(dA + off), (uint4*)(dB + off), len / 16, reps);
                CK(cudaMemcpyAsync(hOut + off, dB + off, len, cudaMemcpyDeviceToHost, st));
                off += len;
            }
            CK(cudaDeviceSynchronize());
            const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (rep > 0 && ms < best) best = ms;
        }
        printf("pipeline ramp %zu..16..%zu MB (%zu chunks), %d streams, kernel x%d   %7.3f ms wall  %6.1f GB/s per direction\n", firstMB, firstMB, sizes.size(), S, reps, best, bytes / best / 1e6);
    }
    return 0;
}
