// tools/gather_bench.cu -- microbenchmark: dependent random gathers of 64-byte records, the memory
// pattern of one BVH traversal step per lane. Calibrates the ceiling of the packed trace kernel
// (DESIGN.md section 5). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
//
// Each thread follows a chain: record i holds (in q0.w) the index of the next record (a random
// permutation), so every step is a dependent, incoherent 64-B fetch, like `cur = child ref`.
// Variants: how the 64 bytes are fetched.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void ldg8(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

template <int MODE>
__global__ void __launch_bounds__(256) chase(const float4* __restrict__ tab, uint32_t n, int steps, float* out, int smemPad) {
    extern __shared__ float4 sh[];
    (void)smemPad;
    uint32_t idx = (uint32_t)(((uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 2654435761ull) % n);
    float acc = 0.f;
    const unsigned lane = threadIdx.x & 31u;
    float4* wbuf = sh + (threadIdx.x >> 5) * 128;   // 32 lanes x 4 quarters (MODE 3)
    // MODE 5: every lane has the TMA unit copy its 64-byte record into its own shared-memory slot (80-byte stride: LDS.128
    // conflict-free), one mbarrier per warp counts the 32 x 64 bytes; the lanes then read their record with four LDS.128.
    // The global side never touches the LSU data pipe; the question is how many small bulk copies per clock an SM's TMA takes.
    __shared__ __align__(8) unsigned long long bars[8];
    unsigned char* tbuf = reinterpret_cast<unsigned char*>(sh) + (threadIdx.x >> 5) * (32 * 80);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[threadIdx.x >> 5]);
    if (MODE == 5) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    for (int s = 0; s < steps; ++s) {
        const float4* p = tab + (size_t)idx * 4;
        float4 q0, q1, q2, q3;
        if (MODE == 0) { q0 = __ldg(p); q1 = __ldg(p + 1); q2 = __ldg(p + 2); q3 = __ldg(p + 3); }
        else if (MODE == 1) { ldg8(p, q0, q1); ldg8(p + 2, q2, q3); }
        else if (MODE == 2) { q0 = __ldg(p); q1 = q2 = q3 = make_float4(0, 0, 0, 0); }
        else if (MODE == 4) { ldg8(p, q0, q1); q2 = q3 = make_float4(0, 0, 0, 0); }
        else if (MODE == 5) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tbuf + lane * 80);
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(32u * 64u) : "memory");
            __syncwarp();
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];"
                         :: "r"(dst), "l"(p), "r"(bar) : "memory");
            uint32_t ready = 0;
            while (!ready)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ready) : "r"(bar), "r"((uint32_t)(s & 1)) : "memory");
            const float4* mine = reinterpret_cast<const float4*>(tbuf + lane * 80);
            q0 = mine[0]; q1 = mine[1]; q2 = mine[2]; q3 = mine[3];
            __syncwarp();
        }
        else {   // MODE 3: four lanes fetch one record (one 64-B coalesced request each), transposed through smem
            const unsigned q = lane & 3u;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const unsigned j = r * 8 + (lane >> 2);
                const uint32_t jidx = __shfl_sync(0xffffffffu, idx, j);
                const float4 v = __ldg(tab + (size_t)jidx * 4 + q);
                wbuf[j * 4 + (q ^ ((j >> 1) & 3u))] = v;
            }
            __syncwarp();
            const unsigned sw = (lane >> 1) & 3u;
            q0 = wbuf[lane * 4 + (0 ^ sw)]; q1 = wbuf[lane * 4 + (1 ^ sw)];
            q2 = wbuf[lane * 4 + (2 ^ sw)]; q3 = wbuf[lane * 4 + (3 ^ sw)];
            __syncwarp();
        }
        acc += q0.x + q1.y + q2.z + q3.x + q1.w + q3.w;
        idx = __float_as_uint(q0.w);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv) {
    int steps = argc > 1 ? atoi(argv[1]) : 64;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("# %s, %d SMs, L2 %d MB\n", prop.name, sms, prop.l2CacheSize >> 20);
    const size_t sizesMB[] = {4, 64, 1024};
    for (size_t mb : sizesMB) {
        const uint32_t n = (uint32_t)(mb * 1024 * 1024 / 64);
        std::vector<float4> h((size_t)n * 4);
        std::vector<uint32_t> perm(n);
        for (uint32_t i = 0; i < n; ++i) perm[i] = i;
        uint64_t st = 88172645463325252ull;
        for (uint32_t i = n - 1; i > 0; --i) {      // Sattolo: one big cycle
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            uint32_t j = (uint32_t)(st % i);
            uint32_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
        }
        for (uint32_t i = 0; i < n; ++i) {
            for (int q = 0; q < 4; ++q) h[(size_t)i * 4 + q] = make_float4(1.f, 2.f, 3.f, 0.f);
            uint32_t nx = perm[i];
            h[(size_t)i * 4].w = *reinterpret_cast<float*>(&nx);
        }
        float4* d; CK(cudaMalloc(&d, h.size() * sizeof(float4)));
        CK(cudaMemcpy(d, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice));
        float* out; CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(float)));
        const int occs[] = {2, 4, 8};                // resident 256-thread blocks per SM
        for (int occ : occs) {
            for (int mode = 0; mode < 6; ++mode) {
                // limit occupancy with dynamic smem: 227 KB / occ
                int smem = (int)((200 * 1024) / occ) & ~1023;
                if (smem < 8 * 32 * 80) smem = 8 * 32 * 80;
                auto launch = [&](int m) {
                    dim3 g(sms * occ), b(256);
                    switch (m) {
                        case 0: CK(cudaFuncSetAttribute(chase<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); chase<0><<<g, b, smem>>>(d, n, steps, out, 0); break;
                        case 1: CK(cudaFuncSetAttribute(chase<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); chase<1><<<g, b, smem>>>(d, n, steps, out, 0); break;
                        case 2: CK(cudaFuncSetAttribute(chase<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); chase<2><<<g, b, smem>>>(d, n, steps, out, 0); break;
                        case 3: CK(cudaFuncSetAttribute(chase<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); chase<3><<<g, b, smem>>>(d, n, steps, out, 0); break;
                        case 5: CK(cudaFuncSetAttribute(chase<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); chase<5><<<g, b, smem>>>(d, n, steps, out, 0); break;
                        case 4: CK(cudaFuncSetAttribute(chase<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); chase<4><<<g, b, smem>>>(d, n, steps, out, 0); break;
                    }
                };
                launch(mode); CK(cudaDeviceSynchronize());
                cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                CK(cudaEventRecord(e0)); launch(mode); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                const double gathers = (double)sms * occ * 256 * steps;
                const char* names[] = {"4xLDG128", "2xLDG256", "1xLDG128(16B)", "coop4+smem", "1xLDG256(32B)", "TMA bulk 64B -> smem"};
                const double bytes = (mode == 2 ? 16.0 : mode == 4 ? 32.0 : 64.0);
                printf("{\"table_MB\": %zu, \"threads_per_sm\": %d, \"mode\": \"%s\", \"ms\": %.3f, \"Ggather_s\": %.2f, \"GB_s\": %.0f}\n",
                       mb, occ * 256, names[mode], ms, gathers / ms / 1e6, gathers * bytes / ms / 1e6);
            }
        }
        CK(cudaFree(d)); CK(cudaFree(out));
    }
    return 0;
}
