// Which instruction mixes overlap on a B200 SM sub-partition?  The scanner's stage 1 retires one INTEGER warp
// instruction per two clocks whatever the alu / IMAD split; this measures whether FP32 instructions (fma pipes) issue
// beside that stream for free.  Each variant runs CH independent dependency chains per thread (latency hidden), 16 warps
// per SM like the scanner, and reports clocks per loop body and warp scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int CH = 4;   // independent chains per thread

#define SHF(d, a, b, s) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s))
#define LOP(d, a, b, c) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define MAD(d, a, b, c) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define MADHI(d, a, b, c) asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define FFMA_R(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define FFMA_I(d, a, c) asm volatile("fma.rn.f32 %0, %1, 0f3F9E3779, %2;" : "=f"(d) : "f"(a), "f"(c))
#define FADD_I(d, a) asm volatile("add.f32 %0, %1, 0f3F800000;" : "=f"(d) : "f"(a))
#define FFMA2_B(d, a, b, c) asm volatile("{.reg .b64 bb, cc; mov.b64 bb, {%2, %2}; mov.b64 cc, {%3, %3}; fma.rn.f32x2 %0, %1, bb, cc;}" : "=l"(d) : "l"(a), "f"(b), "f"(c))
#define PFADD(d, t1, t2) asm volatile("{.reg .pred p; .reg .b32 t; and.b32 t, %1, %2; and.b32 t, t, 0x80000000; setp.ne.u32 p, t, 0; @p add.f32 %0, %0, 0f44800000;}" : "+f"(d) : "r"(t1), "r"(t2))

template <int V>
__global__ void __launch_bounds__(512, 1) pipes(uint32_t* out, const uint32_t* in, int iters, long long* clk) {
    uint32_t x[CH], y[CH];
    float f[CH], g[CH];
    unsigned long long p2[CH], q2[CH];
    const uint32_t m = in[0], s = in[1];
    const float fs = __int_as_float(in[2]);
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        x[k] = in[4 + k] + threadIdx.x; y[k] = in[8 + k] ^ threadIdx.x;
        f[k] = (float)(threadIdx.x + k); g[k] = 1.0f + k;
        p2[k] = ((unsigned long long)__float_as_uint(f[k]) << 32) | __float_as_uint(g[k]); q2[k] = p2[k] + 12345ull;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (V == 0) {          // 5 alu + 3 imad  (stage 1's integer mix)
                SHF(x[k], x[k], y[k], s); MAD(y[k], x[k], m, y[k]); MADHI(x[k], y[k], m, x[k]); MAD(y[k], x[k], m, y[k]);
                SHF(x[k], x[k], y[k], s); SHF(y[k], y[k], x[k], s); LOP(x[k], x[k], y[k], m); SHF(y[k], y[k], x[k], s);
            } else if (V == 1) {   // 4 alu + 2 imad
                SHF(x[k], x[k], y[k], s); MAD(y[k], x[k], m, y[k]); MADHI(x[k], y[k], m, x[k]);
                SHF(x[k], x[k], y[k], s); SHF(y[k], y[k], x[k], s); LOP(x[k], x[k], y[k], m);
            } else if (V == 2) {   // 4 alu + 2 imad + 2 fp (immediate forms)
                SHF(x[k], x[k], y[k], s); MAD(y[k], x[k], m, y[k]); MADHI(x[k], y[k], m, x[k]);
                SHF(x[k], x[k], y[k], s); SHF(y[k], y[k], x[k], s); LOP(x[k], x[k], y[k], m);
                FFMA_I(f[k], f[k], g[k]); FADD_I(g[k], g[k]);
            } else if (V == 3) {   // 4 alu + 2 imad + 5 fp (3 immediate, 2 register forms)
                SHF(x[k], x[k], y[k], s); MAD(y[k], x[k], m, y[k]); MADHI(x[k], y[k], m, x[k]);
                SHF(x[k], x[k], y[k], s); SHF(y[k], y[k], x[k], s); LOP(x[k], x[k], y[k], m);
                FFMA_I(f[k], f[k], g[k]); FADD_I(g[k], g[k]); FFMA_R(f[k], f[k], fs, g[k]); FFMA_I(g[k], g[k], f[k]);
                FFMA_R(f[k], g[k], fs, f[k]);
            } else if (V == 4) {   // 8 alu
                SHF(x[k], x[k], y[k], s); LOP(y[k], x[k], m, y[k]); SHF(x[k], y[k], x[k], s); LOP(y[k], x[k], m, y[k]);
                SHF(x[k], x[k], y[k], s); SHF(y[k], y[k], x[k], s); LOP(x[k], x[k], y[k], m); SHF(y[k], y[k], x[k], s);
            } else if (V == 5) {   // 8 imad
                MAD(x[k], x[k], m, y[k]); MAD(y[k], x[k], m, y[k]); MADHI(x[k], y[k], m, x[k]); MAD(y[k], x[k], m, y[k]);
                MAD(x[k], x[k], m, y[k]); MADHI(y[k], y[k], m, x[k]); MAD(x[k], x[k], m, y[k]); MAD(y[k], y[k], m, x[k]);
            } else if (V == 6) {   // 4 alu + 4 imad, alternating
                SHF(x[k], x[k], y[k], s); MAD(y[k], x[k], m, y[k]); LOP(x[k], x[k], y[k], m); MADHI(y[k], y[k], m, x[k]);
                SHF(x[k], x[k], y[k], s); MAD(y[k], x[k], m, y[k]); LOP(x[k], x[k], y[k], m); MAD(y[k], y[k], m, x[k]);
            } else if (V == 7) {   // 4 alu + 4 fp (register form)
                SHF(x[k], x[k], y[k], s); FFMA_R(f[k], f[k], fs, g[k]); LOP(y[k], x[k], y[k], m); FFMA_R(g[k], g[k], fs, f[k]);
                SHF(x[k], x[k], y[k], s); FFMA_R(f[k], f[k], fs, g[k]); LOP(y[k], x[k], y[k], m); FFMA_R(g[k], g[k], fs, f[k]);
            } else if (V == 8) {   // 8 fp (register form)
                FFMA_R(f[k], f[k], fs, g[k]); FFMA_R(g[k], g[k], fs, f[k]); FFMA_R(f[k], f[k], fs, g[k]); FFMA_R(g[k], g[k], fs, f[k]);
                FFMA_R(f[k], f[k], fs, g[k]); FFMA_R(g[k], g[k], fs, f[k]); FFMA_R(f[k], f[k], fs, g[k]); FFMA_R(g[k], g[k], fs, f[k]);
            } else if (V == 9) {   // 6 int (the last a LOP3 that writes a predicate) + fp(imm) + predicated fadd: the proposed stage-1 shape
                SHF(x[k], x[k], y[k], s); LOP(y[k], x[k], m, y[k]); MAD(x[k], y[k], m, x[k]);
                SHF(y[k], y[k], x[k], s); SHF(x[k], x[k], y[k], s);
                FFMA_I(f[k], f[k], g[k]); PFADD(g[k], x[k], y[k]);
            } else if (V == 11) {  // 4 alu + 2 imad + 1 packed fp (FFMA2: two fp32 lanes per instruction)
                SHF(x[k], x[k], y[k], s); MAD(y[k], x[k], m, y[k]); MADHI(x[k], y[k], m, x[k]);
                SHF(x[k], x[k], y[k], s); SHF(y[k], y[k], x[k], s); LOP(x[k], x[k], y[k], m);
                FFMA2_B(p2[k], p2[k], fs, g[k]);
            } else if (V == 12) {  // 8 packed fp
                FFMA2_B(p2[k], p2[k], fs, g[k]); FFMA2_B(q2[k], q2[k], fs, g[k]); FFMA2_B(p2[k], p2[k], fs, g[k]); FFMA2_B(q2[k], q2[k], fs, g[k]);
                FFMA2_B(p2[k], p2[k], fs, g[k]); FFMA2_B(q2[k], q2[k], fs, g[k]); FFMA2_B(p2[k], p2[k], fs, g[k]); FFMA2_B(q2[k], q2[k], fs, g[k]);
            } else if (V == 13) {  // 5 int + lop3->pred + fp(imm) + predicated fadd, the fp pair NOT adjacent
                SHF(x[k], x[k], y[k], s); FFMA_I(f[k], f[k], g[k]); LOP(y[k], x[k], m, y[k]); MAD(x[k], y[k], m, x[k]);
                SHF(y[k], y[k], x[k], s); SHF(x[k], x[k], y[k], s);
                PFADD(g[k], x[k], y[k]);
            } else if (V == 10) {  // 4 imad + 4 fp (register form)
                MAD(x[k], x[k], m, y[k]); FFMA_R(f[k], f[k], fs, g[k]); MAD(y[k], x[k], m, y[k]); FFMA_R(g[k], g[k], fs, f[k]);
                MADHI(x[k], y[k], m, x[k]); FFMA_R(f[k], f[k], fs, g[k]); MAD(y[k], x[k], m, y[k]); FFMA_R(g[k], g[k], fs, f[k]);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) acc ^= x[k] ^ y[k] ^ __float_as_uint(f[k]) ^ __float_as_uint(g[k]) ^ (uint32_t)p2[k] ^ (uint32_t)(p2[k] >> 32) ^ (uint32_t)q2[k] ^ (uint32_t)(q2[k] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int V>
static void run(const char* name, int n_instr, uint32_t* out, const uint32_t* in, long long* clk, int sms) {
    const int iters = 4000;
    pipes<V><<<sms, 512>>>(out, in, iters, clk);
    CK(cudaDeviceSynchronize());
    pipes<V><<<sms, 512>>>(out, in, iters, clk);
    CK(cudaDeviceSynchronize());
    long long h[1024];
    CK(cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    mean /= sms;
    // 16 warps per SM = 4 per scheduler; every warp runs CH bodies per iteration
    const double per_body = mean / iters / CH / 4.0;
    printf("%-58s %2d instr/body: %6.2f clk per body and scheduler = %5.2f clk/instr\n", name, n_instr, per_body, per_body / n_instr);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    uint32_t *out, *in;
    long long* clk;
    CK(cudaMalloc(&out, sms * 512 * 4));
    CK(cudaMalloc(&in, 64 * 4));
    CK(cudaMalloc(&clk, 1024 * 8));
    uint32_t h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0x9E3779B1u * (i + 1);
    h[1] = 7;
    const float fs = 0.999f;
    h[2] = *reinterpret_cast<const uint32_t*>(&fs);
    CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    printf("%s, %d SMs; 512 threads per SM, %d chains per thread\n", p.name, sms, CH);
    run<0>("V0  5 alu + 3 imad (stage 1 today)", 8, out, in, clk, sms);
    run<1>("V1  4 alu + 2 imad", 6, out, in, clk, sms);
    run<2>("V2  4 alu + 2 imad + 2 fp(imm)", 8, out, in, clk, sms);
    run<3>("V3  4 alu + 2 imad + 5 fp(3 imm, 2 reg)", 11, out, in, clk, sms);
    run<4>("V4  8 alu", 8, out, in, clk, sms);
    run<5>("V5  8 imad", 8, out, in, clk, sms);
    run<6>("V6  4 alu + 4 imad alternating", 8, out, in, clk, sms);
    run<7>("V7  4 alu + 4 fp(reg) alternating", 8, out, in, clk, sms);
    run<8>("V8  8 fp(reg)", 8, out, in, clk, sms);
    run<9>("V9  5 int + lop3->pred + fp(imm) + predicated fadd", 8, out, in, clk, sms);
    run<10>("V10 4 imad + 4 fp(reg) alternating", 8, out, in, clk, sms);
    run<11>("V11 4 alu + 2 imad + 1 packed fp (FFMA2)", 7, out, in, clk, sms);
    run<12>("V12 8 packed fp (FFMA2)", 8, out, in, clk, sms);
    run<13>("V13 like V9, the two fp instructions apart", 8, out, in, clk, sms);
    return 0;
}
