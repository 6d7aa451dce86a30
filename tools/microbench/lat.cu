// Latency / operand-bandwidth microbenchmarks for sm_100a (one warp per SMSP unless noted).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__device__ __forceinline__ float rsq(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float *out, long long *cyc, float seed, int warps_note)
{
    float a = seed + threadIdx.x * 1e-3f, m = 1.0f + seed * 1e-9f, c = seed * 1e-7f;
    float2 b = make_float2(a, a + 1.f), m2 = make_float2(m, m), c2 = make_float2(c, c);
    float2 r[8];
    unsigned u = threadIdx.x + (unsigned)seed, acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = make_float2(seed + i, seed - i);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0) a = __fmaf_rn(a, m, c);                    // FFMA dependent chain
            if (MODE == 1) b = __ffma2_rn(b, m2, c2);                 // FFMA2 dependent chain
            if (MODE == 2) a = rsq(a) + 1.0f;                          // MUFU.RSQ + FADD chain
            if (MODE == 3) a = fminf(a, c) + m;                        // FMNMX + FADD chain
            if (MODE == 4) { r[j & 7] = __ffma2_rn(r[(j + 1) & 7], r[(j + 3) & 7], r[j & 7]); }   // FFMA2, 3 distinct pairs, throughput
            if (MODE == 5) { r[j & 7] = __ffma2_rn(r[j & 7], m2, c2); }                           // FFMA2, reused operands, throughput
            if (MODE == 6) { r[j & 7].x = __fmaf_rn(r[(j + 1) & 7].x, r[(j + 3) & 7].y, r[j & 7].x); r[j & 7].y = __fmaf_rn(r[(j + 1) & 7].y, r[(j + 3) & 7].x, r[j & 7].y);}  // 2 FFMA distinct
            if (MODE == 7) { u = __byte_perm(u, acc, 0x7440 + (j & 3)); acc += u; }               // PRMT + IADD chain
            if (MODE == 8) { acc = __dp4a(u + j, 0x00000001u, acc); }                            // IDP4A chain on acc
            if (MODE == 9) { u = u * 1664525u + 1013904223u; a += float((u >> 8) & 0xFFu); c += float(u >> 24); }   // IMAD + 2 I2F.U8 + 2 FADD
            if (MODE == 12) { u = u * 1664525u + 1013904223u; a += float(u & 0xFFFFu); c += float(u >> 16); }        // IMAD + 2 I2FP(U16?) + 2 FADD
            if (MODE == 10) b = __fmul2_rn(b, m2);                     // FMUL2 chain
            if (MODE == 11) b = __fadd2_rn(b, c2);                     // FADD2 chain
        }
    }
    long long t1 = clock64();
    float s = a + b.x + b.y + u + acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i].x + r[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int threads, double ops_per_step)
{
    const int blocks = 148;
    float *out; long long *cyc;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f, 0);
    cudaDeviceSynchronize();
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f, 0);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += double(v); avg /= blocks;
    const double warps_per_smsp = threads / 128.0;
    printf("%-44s warps/SMSP %4.1f  %7.2f cycles per step per warp; %6.3f steps/clk/SMSP (%g instr/step)\n", name, warps_per_smsp,
           avg / (ITERS * 16.0), warps_per_smsp * ITERS * 16.0 / avg, ops_per_step);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA chain (latency)", 128, 1);
    run<1>("FFMA2 chain (latency)", 128, 1);
    run<10>("FMUL2 chain (latency)", 128, 1);
    run<11>("FADD2 chain (latency)", 128, 1);
    run<2>("MUFU.RSQ + FADD chain", 128, 2);
    run<3>("FMNMX + FADD chain", 128, 2);
    run<4>("FFMA2 3 distinct pairs, 1 warp", 128, 1);
    run<4>("FFMA2 3 distinct pairs, 8 warps", 1024, 1);
    run<5>("FFMA2 reused operands, 8 warps", 1024, 1);
    run<6>("2x FFMA distinct, 8 warps", 1024, 2);
    run<7>("PRMT + IADD chain, 1 warp", 128, 2);
    run<7>("PRMT + IADD, 8 warps", 1024, 2);
    run<8>("IDP4A chain, 1 warp", 128, 2);
    run<8>("IDP4A (+IADD), 8 warps", 1024, 2);
    run<9>("IMAD + 2 I2F.U8 + 2 FADD, 8 warps", 1024, 5);
    run<12>("IMAD + 2 I2F 16-bit + 2 FADD, 8 warps", 1024, 5);
    return 0;
}
