// Micro-benchmark: can the TMA unit's row gather (cp.async.bulk.tensor.2d.tile::gather4, sm_100) carry the scanner's
// 16-byte slot gathers instead of LDGSTS?  Measures rows / clk / SM for random rows of a 64 MiB table, alone and next to
// a 160 KB shared-memory carve-out with LDS traffic (the scanner's stage 1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4 tma_gather4.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mix(uint32_t x) { x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; return x; }

// every lane of every warp issues one gather4 (4 random rows of 16 bytes) per round onto its warp's mbarrier
template <int LANES, bool LDS>
__global__ void __launch_bounds__(512, 1) gather4_kernel(const __grid_constant__ CUtensorMap tm, uint32_t n_rows, int iters,
                                                         uint32_t* out, int filter_words) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* land = smem + warp * (32 * 128 + 128);        // 32 lanes x (4 rows x 16 B, in a 128-byte aligned slot)
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(land + 32 * 128);
    uint32_t* filt = reinterpret_cast<uint32_t*>(smem + 16 * (32 * 128 + 128));
    if (LDS) for (int i = threadIdx.x; i < filter_words; i += blockDim.x) filt[i] = mix(i);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0, bad = 0;
    for (int it = 0; it < iters; ++it) {
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(LANES * 64) : "memory");
        __syncwarp();
        int32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { x = x * 1664525u + 1013904223u; r[k] = (int32_t)__umulhi(x, n_rows); }
        if (lane < LANES)
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                         ::"r"(smem_u32(land + lane * 128)), "l"(&tm), "r"(0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                           "r"(smem_u32(bar)) : "memory");
        if (LDS) {   // stage-1-like traffic while the gathers fly: 64 random word lookups per lane
            uint32_t y = x;
#pragma unroll 8
            for (int k = 0; k < 64; ++k) { y = y * 1664525u + 1013904223u; acc += filt[__umulhi(y, (uint32_t)filter_words)] >> (y & 31); }
        }
        uint32_t ok = 0, spins = 0;
        while (!ok && ++spins < (1u << 24))      // a copy that never completes must not hang the box
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(it & 1) : "memory");
        if (!ok) { bad += 1000000u; break; }
        if (lane < LANES) {
            const uint4* rows = reinterpret_cast<const uint4*>(land + lane * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint4 v = rows[k]; acc += v.y; bad += (v.x != (uint32_t)r[k]) || (v.w != ~(uint32_t)r[k]); }
        }
        __syncwarp();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (bad ? 0x80000000u : 0u);
    if (bad) atomicAdd(out + gridDim.x * blockDim.x, bad);
}

__global__ void fill(uint4* t, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) t[i] = make_uint4(i, mix(i), i * 3u, ~i);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int LANES, bool LDS>
static void run(const CUtensorMap& tm, uint32_t n_rows, int sms, int clock_khz, uint32_t* d_out, const char* what, int box_rows) {
    const int iters = LDS ? 300 : 2000, fw = 36864;
    const size_t smem = 16 * (32 * 128 + 128) + (LDS ? (size_t)fw * 4 : 0);
    CK(cudaFuncSetAttribute(gather4_kernel<LANES, LDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemset(d_out, 0, (size_t)(sms * 512 + 1) * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather4_kernel<LANES, LDS><<<sms, 512, smem>>>(tm, n_rows, 10, d_out, fw);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: launch failed: %s\n", what, cudaGetErrorString(err)); exit(2); }
    CK(cudaMemset(d_out, 0, (size_t)(sms * 512 + 1) * 4));
    cudaEventRecord(e0);
    gather4_kernel<LANES, LDS><<<sms, 512, smem>>>(tm, n_rows, iters, d_out, fw);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    uint32_t bad = 0;
    CK(cudaMemcpy(&bad, d_out + sms * 512, 4, cudaMemcpyDeviceToHost));
    const double rows = (double)sms * 16 * LANES * 4 * iters;
    const double clk = ms * 1e-3 * clock_khz * 1e3;
    printf("%-44s box_rows=%d lanes=%2d lds=%d : %8.3f ms  %.3f rows/clk/SM  wrong_rows=%u\n", what, box_rows, LANES, (int)LDS, ms,
           rows / clk / sms, bad);
}

int main(int argc, char** argv) {
    const int only_box = argc > 1 ? atoi(argv[1]) : 0;   // one box shape per process: a faulting shape poisons the context
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    const uint32_t n_rows = 1u << 22;
    uint4* tab;
    CK(cudaMalloc(&tab, (size_t)n_rows * 16));
    fill<<<p.multiProcessorCount * 8, 256>>>(tab, n_rows);
    uint32_t* d_out;
    CK(cudaMalloc(&d_out, (size_t)(p.multiProcessorCount * 512 + 1) * 4));
    CK(cudaDeviceSynchronize());
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    for (int box_rows = 1; box_rows <= 4; box_rows += 3) {
        if (only_box && box_rows != only_box) continue;
        CUtensorMap tm;
        cuuint64_t dims[2] = {4, n_rows};
        cuuint64_t strides[1] = {16};
        cuuint32_t box[2] = {4, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, tab, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode box_rows=%d failed: %d\n", box_rows, (int)r); continue; }
        printf("SMs %d clock %d kHz, table 64 MiB, box {4 x u32, %d}\n", p.multiProcessorCount, clock_khz, box_rows);
        run<32, false>(tm, n_rows, p.multiProcessorCount, clock_khz, d_out, "gather4, every lane", box_rows);
        run<8, false>(tm, n_rows, p.multiProcessorCount, clock_khz, d_out, "gather4, 8 lanes per warp", box_rows);
        run<1, false>(tm, n_rows, p.multiProcessorCount, clock_khz, d_out, "gather4, 1 lane per warp", box_rows);
        run<32, true>(tm, n_rows, p.multiProcessorCount, clock_khz, d_out, "gather4 + 64 LDS/lane (160 KB filter)", box_rows);
        run<8, true>(tm, n_rows, p.multiProcessorCount, clock_khz, d_out, "gather4 8 lanes + 64 LDS/lane", box_rows);
    }
    return 0;
}
