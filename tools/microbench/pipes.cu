// Issue-rate microbenchmarks for the sm_100a scalar pipes the encoder leans on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Reports warp-instructions per cycle per SM sub-partition (SMSP) for FFMA, the
// packed FFMA2 (fma.rn.f32x2, new on sm_100), and mixes with ALU-pipe integer ops.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 65536;
constexpr int CH = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) k(float *out, long long *cyc, float seed)
{
    float a[CH]; float2 b[CH]; unsigned u[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = seed + i; b[i] = make_float2(seed + i, seed - i); u[i] = threadIdx.x * 7 + i; }
    const float m = 1.0f + threadIdx.x * 1e-9f * seed, c = 1e-7f * seed + threadIdx.x * 1e-12f;   // register operands, not immediates
    const float2 m2 = make_float2(m, m), c2 = make_float2(c, c);
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) a[i] = __fmaf_rn(a[i], m, c);                                    // FFMA
            if (MODE == 1) b[i] = __ffma2_rn(b[i], m2, c2);                                 // FFMA2
            if (MODE == 2) { a[i] = __fmaf_rn(a[i], m, c); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9E37u; }    // FFMA + SHF/LOP/IADD
            if (MODE == 3) { b[i] = __ffma2_rn(b[i], m2, c2); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9E37u; } // FFMA2 + ALU
            if (MODE == 4) u[i] = (u[i] ^ (u[i] >> 3)) + 0x9E37u;                          // ALU only
            if (MODE == 5) a[i] = fminf(fmaxf(a[i], c), m + a[(i + 1) % CH]);              // FMNMX + FADD
            if (MODE == 6) b[i] = __fadd2_rn(__fmul2_rn(b[i], m2), c2);                    // FMUL2 + FADD2
            if (MODE == 7) a[i] = __fadd_rn(__fmul_rn(a[i], m), c);                        // FMUL + FADD
        }
    }
    long long t1 = clock64();
    float s = 0; unsigned su = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { s += a[i] + b[i].x + b[i].y; su += u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + su;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, double instr_per_iter_per_chain)
{
    const int blocks = 148, threads = 1024;             // 1 CTA x 32 warps = 8 warps / SMSP, all co-resident
    float *out; long long *cyc;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += double(v); avg /= blocks;
    // per SMSP: 8 warps each issuing ITERS*CH*instr instructions over `avg` cycles
    const double per_smsp = 8.0 * ITERS * CH * instr_per_iter_per_chain / avg;
    printf("%-28s %8.3f ms  %10.0f cyc (%.3f GHz)  %.3f warp-instr/clk/SMSP (counting %.0f instr per step)\n", name, ms, avg, avg / ms * 1e-6, per_smsp,
           instr_per_iter_per_chain);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA", 1);
    run<1>("FFMA2", 1);
    run<2>("FFMA + 3 ALU (SHF,LOP,IADD)", 4);
    run<3>("FFMA2 + 3 ALU", 4);
    run<4>("3 ALU only", 3);
    run<5>("FMNMX x2 + FADD", 3);
    run<6>("FMUL2 + FADD2", 2);
    run<7>("FMUL + FADD", 2);
    return 0;
}
