// Per-phase throughput of the 4x4 block encode (astc_block.cuh) in isolation, on register-resident
// data, as a function of the warps resident per SM sub-partition (SMSP).  Answers: how far is each
// phase from its nominal FP32-pipe occupancy when nothing else competes for the SMSP, and how much
// does residency buy?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I astc_encoder_b200/csrc -I include \
//        -o tools/microbench/phases tools/microbench/phases.cu && tools/microbench/phases
// Output: cycles of SMSP time per warp-block (= per 32 blocks) for each phase and residency.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "astc_block.cuh"

using namespace astc;
using dev::f2;
using dev::Texel;

struct Texels4x4 {
    static constexpr bool kStreamed = false;
    Texel t[16];
    __device__ __forceinline__ Texel raw(int k) const { return t[k]; }
    __device__ __forceinline__ void fence() const {}
};

__device__ const dev::TableImage g_tab = dev::make_table_image<QUANT_12>();

constexpr int ITERS = 128;
constexpr int WAVES = 12;                      // full waves of CTAs per launch: steady-state throughput, not one wave's placement

// PHASE: 0 = block_stats (mean + covariance), 1 = power iteration (8 rounds), 2 = project_block,
//        3 = pack_block, 4 = whole encode_block, 5 = PI with two independent blocks interleaved
template <int PHASE, int MINB>
__global__ void __launch_bounds__(128, MINB) k(float *out, long long *cyc, float seed)
{
    extern __shared__ unsigned char pad[];
    __shared__ dev::SharedTables st;
    for (int i = threadIdx.x; i < int(sizeof(dev::TableImage) / 4); i += blockDim.x) reinterpret_cast<uint32_t *>(&st)[i] = reinterpret_cast<const uint32_t *>(&g_tab)[i];
    __syncthreads();
    const uint32_t s_field = uint32_t(__cvta_generic_to_shared(st.field)), s_trit = uint32_t(__cvta_generic_to_shared(st.trit_scattered));

    uint32_t rng = (blockIdx.x * 128 + threadIdx.x) * 2654435761u + 12345u;
    auto next_byte = [&]() { rng = rng * 1664525u + 1013904223u; return float((rng >> 24) & 255u); };
    Texels4x4 tx;
    f2 sum_lo = dev::bc(0.f), sum_hi = dev::bc(0.f);
    dev::BlockStats bs{};
    f2 alo = dev::mk(0.5f, 0.5f), ahi = dev::mk(0.5f, 0.5f);
    dev::Projected pr{};
    float acc = 0.f;
    if (PHASE == 0 || PHASE == 2 || PHASE == 4) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float c0 = next_byte(), c1 = next_byte(), c2 = next_byte(), c3 = next_byte();
            tx.t[i].lo = dev::mk(c0 / 255.0f, c1 / 255.0f);
            tx.t[i].hi = dev::mk(c2 / 255.0f, c3 / 255.0f);
            asm volatile("" : "+f"(tx.t[i].lo.x), "+f"(tx.t[i].lo.y), "+f"(tx.t[i].hi.x), "+f"(tx.t[i].hi.y));   // opaque: no rematerialisation in the loop
            sum_lo = dev::add2(sum_lo, dev::mk(c0, c1));
            sum_hi = dev::add2(sum_hi, dev::mk(c2, c3));
        }
        bs.mean_lo = dev::mul2(sum_lo, dev::bc(1.0f / 16));
        bs.mean_hi = dev::mul2(sum_hi, dev::bc(1.0f / 16));
    }
    if (PHASE == 1) {                                             // a random Gram matrix (three outer products)
        f2 z = dev::bc(0.f);
        bs.m = dev::Cols{z, z, z, z, z, z, z, z};
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
            const float a = next_byte() - 128.f, b = next_byte() - 128.f, c = next_byte() - 128.f, d = next_byte() - 128.f;
            const f2 lo = dev::mk(a, b), hi = dev::mk(c, d);
            bs.m.c0lo = dev::fma2(lo, dev::bc(a), bs.m.c0lo); bs.m.c0hi = dev::fma2(hi, dev::bc(a), bs.m.c0hi);
            bs.m.c1lo = dev::fma2(lo, dev::bc(b), bs.m.c1lo); bs.m.c1hi = dev::fma2(hi, dev::bc(b), bs.m.c1hi);
            bs.m.c2lo = dev::fma2(lo, dev::bc(c), bs.m.c2lo); bs.m.c2hi = dev::fma2(hi, dev::bc(c), bs.m.c2hi);
            bs.m.c3lo = dev::fma2(lo, dev::bc(d), bs.m.c3lo); bs.m.c3hi = dev::fma2(hi, dev::bc(d), bs.m.c3hi);
        }
    }
    if (PHASE == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) pr.pw[i] = dev::mk(next_byte(), next_byte());
        pr.wlo = 0.f;
        pr.span = 1.0f / 255.0f;
        pr.ep_lo = rng;
        pr.ep_hi = rng * 3u;
    }
    uint32_t acci = 0;

    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (PHASE == 0) {
            bs = dev::block_stats<4, false>(tx, sum_lo, sum_hi);
            const dev::Cols &m = bs.m;                               // consume every output (10 adds of overhead per block)
            const f2 all = dev::add2(dev::add2(dev::add2(m.c0lo, m.c0hi), dev::add2(m.c1lo, m.c1hi)), dev::add2(dev::add2(m.c2lo, m.c2hi), dev::add2(m.c3lo, m.c3hi)));
            tx.t[3].lo.x = dev::ffma(all.x, 1e-12f, tx.t[3].lo.x);
            sum_lo.x = dev::ffma(all.y, 1e-12f, sum_lo.x);                 // both means move: nothing is loop-invariant
            sum_hi.y = dev::ffma(all.x, 1e-12f, sum_hi.y);
            sum_lo.y = dev::ffma(all.x, 1e-12f, sum_lo.y);
            sum_hi.x = dev::ffma(all.y, 1e-12f, sum_hi.x);
        } else if (PHASE == 1) {
            dev::power_iteration<false, true>(bs.m, alo, ahi);
            bs.m.c0lo.x = dev::ffma(dev::fadd(alo.x, alo.y), 1e-9f, bs.m.c0lo.x);
            bs.m.c3hi.y = dev::ffma(dev::fadd(ahi.x, ahi.y), 1e-9f, bs.m.c3hi.y);
        } else if (PHASE == 2) {
            pr = dev::project_block<4, false, false>(tx, bs.mean_lo, bs.mean_hi, alo, ahi);
            f2 all = pr.pw[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) all = dev::add2(all, pr.pw[i]);
            alo.x = dev::ffma(pr.span, 1e-12f, alo.x);
            bs.mean_lo.x = dev::ffma(dev::fadd(all.x, all.y), 1e-12f, bs.mean_lo.x);
            bs.mean_hi.y = dev::ffma(dev::fadd(all.x, all.y), 1e-12f, bs.mean_hi.y);
            alo.y = dev::ffma(pr.wlo, 1e-12f, alo.y);
            ahi.x = dev::ffma(pr.wlo, 1e-12f, ahi.x);
            ahi.y = dev::ffma(pr.span, 1e-12f, ahi.y);
            acci += pr.ep_lo ^ pr.ep_hi;
        } else if (PHASE == 3) {
            const uint4 b = dev::pack_block<false>(pr, s_field, s_trit);
            acci += b.x ^ b.y ^ b.z ^ b.w;
            pr.pw[5].x = dev::ffma(float(b.w & 1u), 1e-3f, pr.pw[5].x);
        } else if (PHASE == 4) {
            const uint4 b = dev::encode_block<4, false, false, false>(tx, sum_lo, sum_hi, s_field, s_trit);
            acci += b.x ^ b.y ^ b.z ^ b.w;
            tx.t[3].lo.x = dev::ffma(float(b.w & 1u), 1e-9f, tx.t[3].lo.x);
        }
    }
    const long long t1 = clock64();
    if (PHASE != 4) acc += bs.m.c0lo.x + bs.m.c3hi.y + alo.x + ahi.y + pr.span + pr.pw[3].x;
    acc += sum_lo.x;
    if (PHASE != 1 && PHASE != 3) acc += tx.t[5].lo.x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + float(acci);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    (void)pad;
}

template <int PHASE, int MINB = 4>
void run(const char *name, int ctas_per_sm)
{
    // one 128-thread CTA = one warp per SMSP; dynamic shared memory caps the CTAs resident per SM
    const int blocks = 148 * ctas_per_sm * WAVES;
    const size_t smem = ctas_per_sm >= 8 ? 0 : (size_t(220) * 1024 / ctas_per_sm - 6 * 1024) & ~size_t(1023);
    float *out; long long *cyc;
    cudaMalloc(&out, size_t(blocks) * 128 * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    cudaFuncSetAttribute(k<PHASE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k<PHASE, MINB>, 128, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<PHASE, MINB><<<blocks, 128, smem>>>(out, cyc, 1.0f);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    long long *h = (long long *)malloc(blocks * sizeof(long long));
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += double(h[i]); avg /= blocks;
    // every SMSP hosts `resident` warps, each doing ITERS warp-blocks in `avg` cycles
    // SMSP time per warp-block from the kernel's duration: 592 SMSPs share blocks * 4 warps * ITERS warp-blocks
    const double smsp_cycles = ms * 1.965e6 * 592.0 / (double(blocks) * 4.0 * ITERS);
    printf("%-32s %d warps/SMSP (occupancy API: %d CTAs/SM): %7.1f cycles of SMSP time per warp-block (kernel %.3f ms); a warp's own latency per block %7.1f cycles\n",
           name, ctas_per_sm, resident, smsp_cycles, ms, avg / ITERS);
    free(h); cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {1, 2, 3, 4}) {
        run<0>("mean+covariance", w);
        run<1>("power iteration x8", w);
        run<2>("minmax+endpoints+proj", w);
        run<3>("quantise+pack", w);
    }
    for (int w : {4, 5, 6, 8}) {
        run<1, 8>("power iteration x8 (<=64 regs)", w);
        run<3, 8>("quantise+pack (<=64 regs)", w);
    }
    return 0;
}
