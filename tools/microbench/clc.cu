// Cluster-launch-control (CLC, sm_100) microbenchmark: what does handing out work tiles cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o clc clc.cu && ./clc
// A "tile" is D cycles of spinning by every warp of a 128-thread CTA; 4 CTAs per SM (capped by
// shared memory), i.e. the shape of the encode kernels.  Modes:
//   0 classic      grid = T CTAs, one tile each (the hardware block scheduler hands them out)
//   1 clc-cta      grid = T CTAs; a running CTA cancels a pending one and takes over its index
//                  (request one tile ahead, one __syncthreads per tile)
//   2 atomic-warp  persistent grid, every warp draws 32-thread tiles from a global counter
//   3 atomic-cta   persistent grid, thread 0 draws for the CTA (smem broadcast + barrier)
// Every wait loop has a clock deadline: a protocol mistake ends the kernel instead of hanging the box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ unsigned int g_counter;
__device__ unsigned int g_error;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void spin(long long cycles)
{
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity)
{
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 400000000ll) { atomicAdd(&g_error, 1u); return false; }
    }
    return true;
}
__device__ __forceinline__ void clc_try_cancel(uint4 *response, uint64_t *bar)
{
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
                 ::"r"(smem_u32(response)), "r"(smem_u32(bar)) : "memory");
}
// returns true and the cancelled CTA's blockIdx.x when the request succeeded
__device__ __forceinline__ bool clc_query(const uint4 *response, uint32_t &ctaid_x)
{
    uint32_t ok = 0, x = 0;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b128 r;\n\tld.shared.b128 r, [%2];\n\t"
                 "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p, r;\n\tselp.u32 %0, 1, 0, p;\n\t"
                 "@p clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %1, r;\n\t}"
                 : "=r"(ok), "=r"(x) : "r"(smem_u32(response)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ctaid_x = x;
    return ok != 0;
}

extern __shared__ unsigned char dyn_smem[];

template <int MODE>
__global__ void __launch_bounds__(128) k(unsigned int tiles, long long work, unsigned int *done_tiles)
{
    __shared__ uint4 s_resp[2];
    __shared__ uint64_t s_bar[2];
    __shared__ unsigned int s_tile;
    unsigned int mine = 0;
    if (MODE == 0) {
        spin(work);
        mine = 1;
    } else if (MODE == 1) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) { mbar_expect_tx(&s_bar[0], 16); clc_try_cancel(&s_resp[0], &s_bar[0]); }
        for (unsigned int i = 0;; ++i) {
            spin(work);                                   // this tile
            ++mine;
            if (!mbar_wait(&s_bar[i & 1], (i >> 1) & 1)) break;
            uint32_t next;
            const bool more = clc_query(&s_resp[i & 1], next);
            __syncthreads();                              // everyone has read slot i&1 ... and slot (i+1)&1 long ago
            if (!more) break;
            if (threadIdx.x == 0) { mbar_expect_tx(&s_bar[(i + 1) & 1], 16); clc_try_cancel(&s_resp[(i + 1) & 1], &s_bar[(i + 1) & 1]); }
        }
    } else if (MODE == 2) {
        const unsigned int lane = threadIdx.x & 31;
        const unsigned int wtiles = tiles * 4;            // warp tiles
        unsigned int t = 0, nxt = 0;
        if (lane == 0) t = atomicAdd(&g_counter, 1u);
        if (lane == 0) nxt = atomicAdd(&g_counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        while (t < wtiles) {
            unsigned int nn = 0;
            if (lane == 0) nn = atomicAdd(&g_counter, 1u);
            spin(work);
            ++mine;
            t = __shfl_sync(0xffffffffu, nxt, 0);
            nxt = nn;
        }
    } else {
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_tile = atomicAdd(&g_counter, 1u);
            __syncthreads();
            if (s_tile >= tiles) break;
            spin(work);
            ++mine;
        }
    }
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(done_tiles, mine);
}

template <int MODE>
static void run(const char *name, unsigned int tiles, long long work, int sms)
{
    unsigned int *done;
    cudaMalloc(&done, 4);
    const size_t smem = 50 * 1024;                        // 4 CTAs per SM
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    const unsigned int grid = (MODE >= 2) ? unsigned(sms * 4) : tiles;
    float best = 1e30f;
    unsigned int got = 0, err = 0;
    for (int rep = 0; rep < 5; ++rep) {
        unsigned int zero = 0;
        cudaMemcpyToSymbol(g_counter, &zero, 4);
        cudaMemcpyToSymbol(g_error, &zero, 4);
        cudaMemset(done, 0, 4);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        k<MODE><<<grid, 128, smem>>>(tiles, work, done);
        cudaEventRecord(b);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
        cudaMemcpy(&got, done, 4, cudaMemcpyDeviceToHost);
        cudaMemcpyFromSymbol(&err, g_error, 4);
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
    const unsigned int want = (MODE == 2) ? tiles * 4 : tiles * 4;   // warp-tiles counted by lane 0 of every warp
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ideal_ms = double(tiles) * double(work) / (double(sms) * 4.0) / (double(clk) * 1e3) * 1e3;
    printf("%-12s tiles %7u work %6lld cyc: %8.3f ms  (ideal %7.3f)  per-tile overhead %6.0f ns  warp-tiles %u/%u err %u\n", name, tiles, work,
           best, ideal_ms, (best - ideal_ms) * 1e6 / (double(tiles) / (sms * 4.0)), got, want, err);
    cudaFree(done);
}

int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long long works[] = {0, 1500, 6000};
    const unsigned int tiles[] = {8192, 131072};
    for (unsigned int t : tiles)
        for (long long w : works) {
            run<0>("classic", t, w, sms);
            run<1>("clc-cta", t, w, sms);
            run<2>("atomic-warp", t, w, sms);
            run<3>("atomic-cta", t, w, sms);
        }
    return 0;
}
