// Micro-benchmarks that size the scanner design on the actual B200 (random shared-memory lookups, random L2
// gathers, random DSMEM reads).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; return x; }

// ---- 1. shared-memory random word lookups ----------------------------------------------------------------
template <int ILP>
__global__ void smem_lookup(uint32_t* out, int words, int iters) {
    extern __shared__ uint32_t s[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = mix(i);
    __syncthreads();
    uint32_t x = threadIdx.x * 7919u + blockIdx.x * 104729u + 1u, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            x = x * 1664525u + 1013904223u;
            uint32_t idx = __umulhi(x, (uint32_t)words);
            acc += s[idx] >> (x & 31);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- 2. global random gathers (L2 resident table) ----------------------------------------------------------
template <typename T, int ILP>
__global__ void gmem_gather(const T* __restrict__ tab, uint32_t n, uint32_t* out, int iters) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int it = 0; it < iters; ++it) {
        T v[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            x = x * 1664525u + 1013904223u;
            v[k] = tab[__umulhi(x, n)];
        }
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(&v[k]);
            for (int q = 0; q < (int)(sizeof(T) / 4); ++q) acc += p[q];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// sparse gathers: only `active` of 32 lanes load each time (models divergent candidate processing)
template <int ILP>
__global__ void gmem_gather_sparse(const uint4* __restrict__ tab, uint32_t n, uint32_t* out, int iters, int active) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    const bool on = (threadIdx.x & 31) < active;
    for (int it = 0; it < iters; ++it) {
        uint4 v[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            x = x * 1664525u + 1013904223u;
            if (on) v[k] = tab[__umulhi(x, n)]; else v[k] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc += v[k].x + v[k].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- 3. DSMEM random reads -----------------------------------------------------------------------------------
template <int ILP>
__global__ void dsmem_lookup(uint32_t* out, int words, int iters, int local_only) {
    extern __shared__ uint32_t s[];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned cs = cl.num_blocks();
    for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = mix(i + cl.block_rank() * 77);
    cl.sync();
    uint32_t x = threadIdx.x * 7919u + blockIdx.x * 104729u + 1u, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            x = x * 1664525u + 1013904223u;
            uint32_t idx = __umulhi(x, (uint32_t)words);
            unsigned r = local_only ? cl.block_rank() : ((x >> 3) % cs);
            const uint32_t* p = cl.map_shared_rank(s, r);
            acc += p[idx];
        }
    }
    cl.sync();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static float time_it(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

// gathers with a cache operator, run under a large dynamic-smem carve-out (what the scanner CTA looks like)
template <int MODE, int ILP>
__global__ void gmem_gather_mode(const uint4* __restrict__ tab, uint32_t n, uint32_t* out, int iters) {
    extern __shared__ uint4 sbuf[];
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    if (MODE == 5) {
        unsigned long long* bar = reinterpret_cast<unsigned long long*>(sbuf + blockDim.x * ILP) + (threadIdx.x >> 5);
        if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    for (int it = 0; it < iters; ++it) {
        uint4 v[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            x = x * 1664525u + 1013904223u;
            const uint4* p = tab + __umulhi(x, n);
            if (MODE == 0) v[k] = *p;
            else if (MODE == 1) v[k] = __ldg(p);
            else if (MODE == 2) v[k] = __ldcg(p);
            else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(p));
            else if (MODE == 4) {
                uint32_t d = (uint32_t)__cvta_generic_to_shared(&sbuf[threadIdx.x * ILP + k]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(p) : "memory");
            } else if (MODE == 5) {  // per-lane TMA bulk copy, one mbarrier per warp
                unsigned long long* bar = reinterpret_cast<unsigned long long*>(sbuf + blockDim.x * ILP) + (threadIdx.x >> 5);
                uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
                uint32_t d = (uint32_t)__cvta_generic_to_shared(&sbuf[threadIdx.x * ILP + k]);
                if (k == 0) {
                    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(32 * ILP * 16) : "memory");
                    __syncwarp();
                }
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(d), "l"(p), "r"(b) : "memory");
            } else if (MODE == 6) {  // 64-bit atomic OR with 0: served by L2, no L1 line
                unsigned long long r = atomicOr((unsigned long long*)p, 0ull);
                v[k].x = (uint32_t)r; v[k].w = (uint32_t)(r >> 32); v[k].y = v[k].z = 0;
            }
        }
        if (MODE == 5) {
            unsigned long long* bar = reinterpret_cast<unsigned long long*>(sbuf + blockDim.x * ILP) + (threadIdx.x >> 5);
            uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), ok = 0;
            while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b), "r"(it & 1) : "memory");
#pragma unroll
            for (int k = 0; k < ILP; ++k) v[k] = sbuf[threadIdx.x * ILP + k];
            __syncwarp();
        }
        if (MODE == 4) {
            asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int k = 0; k < ILP; ++k) v[k] = sbuf[threadIdx.x * ILP + k];
        }
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc += v[k].x + v[k].w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
static void run_mode(const char* name, const uint4* tab, uint32_t n16, uint32_t* out, int sms, cudaEvent_t e0, cudaEvent_t e1) {
    for (int kb : {100, 132, 164, 196, 224}) for (int threads : {512, 1024}) {
        int iters = 100;
        CK(cudaFuncSetAttribute(gmem_gather_mode<MODE, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
        gmem_gather_mode<MODE, 4><<<sms, threads, kb * 1024>>>(tab, n16, out, 3);
        CK(cudaEventRecord(e0)); gmem_gather_mode<MODE, 4><<<sms, threads, kb * 1024>>>(tab, n16, out, iters); CK(cudaEventRecord(e1));
        float ms = time_it(e0, e1); double n = (double)sms * threads * iters * 4;
        printf("gather16 %-14s smem %3d KB %4d thr ILP4: %.3f ms  %.1f G/s (%.2f /clk/SM)\n", name, kb, threads, ms, n / ms / 1e6, n / ms / 1e6 / sms / 1.965);
    }
}


int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, smem optin %zu\n", pr.name, sms, pr.sharedMemPerBlockOptin);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    uint32_t* out; CK(cudaMalloc(&out, 64 << 20));

    if (getenv("UB_MODES")) {
        uint32_t n16 = (uint32_t)(((size_t)32 << 20) / 16);
        uint4* tab; CK(cudaMalloc(&tab, (size_t)32 << 20)); CK(cudaMemset(tab, 1, (size_t)32 << 20));
        run_mode<0>("ld", tab, n16, out, sms, e0, e1);
        run_mode<4>("cp.async.cg", tab, n16, out, sms, e0, e1);
        run_mode<5>("tma.bulk16", tab, n16, out, sms, e0, e1);
        run_mode<6>("atom.or.b64", tab, n16, out, sms, e0, e1);
        printf("done\n");
        return 0;
    }
    // 1. smem
    for (int kb : {64, 128, 192}) for (int threads : {512, 1024}) {
        int words = kb * 256, iters = 2000;
        CK(cudaFuncSetAttribute(smem_lookup<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
        smem_lookup<8><<<sms, threads, kb * 1024>>>(out, words, 10);
        CK(cudaEventRecord(e0)); smem_lookup<8><<<sms, threads, kb * 1024>>>(out, words, iters); CK(cudaEventRecord(e1));
        float ms = time_it(e0, e1);
        double n = (double)sms * threads * iters * 8;
        printf("smem_lookup  %3d KB %4d thr : %.3f ms  %.1f G lookups/s  (%.2f /clk/SM @1.965GHz)\n", kb, threads, ms, n / ms / 1e6, n / ms / 1e6 / sms / 1.965);
    }
    // 2. gmem gathers
    for (int mb : {4, 16, 64}) {
        uint32_t n16 = (uint32_t)(((size_t)mb << 20) / 16);
        uint4* tab; CK(cudaMalloc(&tab, (size_t)mb << 20)); CK(cudaMemset(tab, 1, (size_t)mb << 20));
        for (int threads : {512, 1024}) {
            int iters = 200;
            gmem_gather<uint4, 4><<<sms, threads>>>(tab, n16, out, 5);
            CK(cudaEventRecord(e0)); gmem_gather<uint4, 4><<<sms, threads>>>(tab, n16, out, iters); CK(cudaEventRecord(e1));
            float ms = time_it(e0, e1); double n = (double)sms * threads * iters * 4;
            printf("gather 16B   %3d MB %4d thr x1 CTA/SM ILP4: %.3f ms  %.1f G/s (%.2f /clk/SM)\n", mb, threads, ms, n / ms / 1e6, n / ms / 1e6 / sms / 1.965);
            CK(cudaEventRecord(e0)); gmem_gather<uint4, 8><<<sms * 2, threads>>>(tab, n16, out, iters); CK(cudaEventRecord(e1));
            ms = time_it(e0, e1); n = (double)sms * 2 * threads * iters * 8;
            printf("gather 16B   %3d MB %4d thr x2 CTA/SM ILP8: %.3f ms  %.1f G/s (%.2f /clk/SM)\n", mb, threads, ms, n / ms / 1e6, n / ms / 1e6 / sms / 1.965);
        }
        {
            int threads = 1024, iters = 200;
            CK(cudaEventRecord(e0)); gmem_gather<uint2, 8><<<sms * 2, threads>>>((const uint2*)tab, n16 * 2, out, iters); CK(cudaEventRecord(e1));
            float ms = time_it(e0, e1); double n = (double)sms * 2 * threads * iters * 8;
            printf("gather  8B   %3d MB %4d thr x2 CTA/SM ILP8: %.3f ms  %.1f G/s (%.2f /clk/SM)\n", mb, threads, ms, n / ms / 1e6, n / ms / 1e6 / sms / 1.965);
            CK(cudaEventRecord(e0)); gmem_gather<uint32_t, 8><<<sms * 2, threads>>>((const uint32_t*)tab, n16 * 4, out, iters); CK(cudaEventRecord(e1));
            ms = time_it(e0, e1);
            printf("gather  4B   %3d MB %4d thr x2 CTA/SM ILP8: %.3f ms  %.1f G/s (%.2f /clk/SM)\n", mb, threads, ms, n / ms / 1e6, n / ms / 1e6 / sms / 1.965);
            for (int active : {1, 4, 8, 16}) {
                CK(cudaEventRecord(e0)); gmem_gather_sparse<8><<<sms * 2, threads>>>(tab, n16, out, iters, active); CK(cudaEventRecord(e1));
                ms = time_it(e0, e1); double na = n * active / 32.0;
                printf("gather 16B sparse %2d/32 lanes %3d MB: %.3f ms  %.1f G/s (%.2f /clk/SM)\n", active, mb, ms, na / ms / 1e6, na / ms / 1e6 / sms / 1.965);
            }
        }
        CK(cudaFree(tab));
    }
    // 3. DSMEM
    for (int cs : {2, 4, 8}) for (int local_only : {1, 0}) {
        int kb = 96, words = kb * 256, threads = 1024, iters = 500;
        CK(cudaFuncSetAttribute(dsmem_lookup<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
        cudaLaunchConfig_t cfg = {};
        int grid = (sms / cs) * cs;
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = kb * 1024;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        cudaOccupancyMaxActiveClusters(&nclusters, dsmem_lookup<8>, &cfg);
        CK(cudaLaunchKernelEx(&cfg, dsmem_lookup<8>, out, words, 5, local_only));
        CK(cudaEventRecord(e0)); CK(cudaLaunchKernelEx(&cfg, dsmem_lookup<8>, out, words, iters, local_only)); CK(cudaEventRecord(e1));
        float ms = time_it(e0, e1); double n = (double)grid * threads * iters * 8;
        printf("dsmem cs=%d local_only=%d grid=%d (max active clusters %d): %.3f ms  %.1f G/s (%.2f /clk/SM)\n", cs, local_only, grid, nclusters, ms, n / ms / 1e6, n / ms / 1e6 / grid / 1.965);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
