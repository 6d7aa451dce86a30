// Operand-bandwidth microbenchmark for the packed FP32 ops of sm_100a: how many cycles an
// FFMA2 costs as a function of how many distinct register operands it reads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o opnd opnd.cu && ./opnd
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 8192;
constexpr int N = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) k(float *out, long long *cyc, float seed)
{
    float2 r[N], x[N], y[N];
    float s[N], a[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        r[i] = make_float2(seed + i, seed - i);
        x[i] = make_float2(1.0f + (seed + i) * 1e-9f, 1.0f - (seed + i) * 1e-9f);
        y[i] = make_float2((seed + i) * 1e-7f, (seed - i) * 1e-7f);
        s[i] = 1.0f + (threadIdx.x + i) * 1e-9f * seed;
        a[i] = seed * i;
    }
    const float m = 1.0f + threadIdx.x * 1e-9f * seed;
    const float2 m2 = make_float2(m, m);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep)
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE == 0) r[i] = __ffma2_rn(r[i], m2, m2);
            if (MODE == 1) r[i] = __ffma2_rn(x[i], m2, r[i]);
            if (MODE == 2) r[i] = __ffma2_rn(x[i], y[i], r[i]);
            if (MODE == 3) r[i] = __ffma2_rn(x[i], make_float2(s[i], s[i]), r[i]);
            if (MODE == 4) a[i] = __fmaf_rn(x[i].x, y[i].x, a[i]);
            if (MODE == 5) a[i] = __fmaf_rn(x[i].x, m, a[i]);
            if (MODE == 6) r[i] = __fmul2_rn(r[i], x[i]);                            // 2 pairs, chain through r
            if (MODE == 7) r[i] = __fadd2_rn(r[i], x[i]);
            if (MODE == 8) r[i] = __ffma2_rn(x[i], make_float2(r[i].x, r[i].x), y[i]);     // acc-free: 2 pairs + scalar from the chain
            if (MODE == 9) r[i] = __fmul2_rn(x[i], make_float2(r[i].x, r[i].x));          // FMUL2 pair * scalar
            if (MODE == 10) r[i] = __ffma2_rn(x[i], make_float2(s[rep], s[rep]), r[i]);     // scalar shared by the 8 consecutive instructions (reuse cache)
            if (MODE == 11) r[i] = __ffma2_rn(r[i], make_float2(255.0f, 255.0f), y[i]);     // immediate multiplier
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) acc += r[i].x + r[i].y + a[i] + x[i].x + y[i].y + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name)
{
    const int blocks = 148, threads = 1024;
    float *out; long long *cyc;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    k<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : h) avg += double(v); avg /= blocks;
    printf("%-58s %6.3f cycles per instruction per SMSP\n", name, avg / (8.0 * ITERS * 2 * N));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA2 r = r*m2+m2          (1 varying pair)");
    run<1>("FFMA2 r = x*m2+r           (2 varying pairs)");
    run<2>("FFMA2 r = x*y+r            (3 varying pairs)");
    run<3>("FFMA2 r = x*bcast(s)+r     (2 pairs + scalar)");
    run<8>("FFMA2 r = x*bcast(r.x)+y   (2 pairs + scalar)");
    run<4>("FFMA  a = x*y+a            (3 varying regs)");
    run<5>("FFMA  a = x*m+a            (2 varying regs)");
    run<6>("FMUL2 r = r*x              (2 varying pairs)");
    run<9>("FMUL2 r = x*bcast(r.x)     (pair + scalar)");
    run<10>("FFMA2 r = x*bcast(s)+r     (2 pairs + scalar in reuse)");
    run<11>("FFMA2 r = r*255+y          (2 pairs + immediate)");
    run<7>("FADD2 r = r+x              (2 varying pairs)");
    return 0;
}
