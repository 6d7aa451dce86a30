// TLD (integer-coordinate point fetch) throughput from a pitch-linear RGBA8 texture in
// normalized-float mode, for the access patterns the encoder could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o texrate texrate.cu && ./texrate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NCOMP>
__device__ __forceinline__ float fetch(cudaTextureObject_t t, int x, int y)
{
    float r, g, b, a;
    asm volatile("tex.2d.v4.f32.s32 {%0, %1, %2, %3}, [%4, {%5, %6}];" : "=f"(r), "=f"(g), "=f"(b), "=f"(a) : "l"(t), "r"(x), "r"(y));
    if (NCOMP == 1) return r;
    if (NCOMP == 2) return r + g;
    return (r + g) + (b + a);
}

// MODE 0: thread = 4x4 block, 16 fetches (encoder pattern).  MODE 1: thread = one texel column walk, lanes adjacent in x.
template <int MODE, int NCOMP>
__global__ void k(cudaTextureObject_t t, int w, int h, float *out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    if (MODE == 0) {
        const int bw = w / 4, by = tid / bw, bx = tid - by * bw;
        if (by * 4 >= h) return;
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) acc += fetch<NCOMP>(t, bx * 4 + (k2 & 3), by * 4 + (k2 >> 2));
    } else {
        const int per_row = w, y0 = (tid / per_row) * 16, x = tid % per_row;
        if (y0 >= h) return;
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) acc += fetch<NCOMP>(t, x, y0 + k2);
    }
    if (acc == 123.456f) out[tid] = acc;
}

template <int MODE, int NCOMP>
void run(const char *name, cudaTextureObject_t t, int w, int h, float *out)
{
    const int threads = (w / 4) * (h / 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, NCOMP><<<(threads + 127) / 128, 128>>>(t, w, h, out);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<MODE, NCOMP><<<(threads + 127) / 128, 128>>>(t, w, h, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double texels = double(w) * h;
    printf("%-46s %.3f ms  %.1f Gtexel/s  %.2f texels/clk/SM (1.9 GHz, 148 SMs)\n", name, ms, texels / ms / 1e6, texels / (ms * 1e-3) / 148 / 1.9e9);
}

int main()
{
    const int w = 16384, h = 8192;
    uint8_t *d; size_t pitch = size_t(w) * 4;
    cudaMalloc(&d, pitch * h); cudaMemset(d, 0x55, pitch * h);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = d;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>(); rd.res.pitch2D.width = w; rd.res.pitch2D.height = h; rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat;
    cudaTextureObject_t t; cudaCreateTextureObject(&t, &rd, &td, nullptr);
    float *out; cudaMalloc(&out, size_t(w / 4) * (h / 4) * 4);
    run<0, 4>("block pattern, 4 components", t, w, h, out);
    run<0, 2>("block pattern, 2 components", t, w, h, out);
    run<0, 1>("block pattern, 1 component", t, w, h, out);
    run<1, 4>("dense pattern (lanes adjacent), 4 components", t, w, h, out);
    run<1, 2>("dense pattern, 2 components", t, w, h, out);
    run<1, 1>("dense pattern, 1 component", t, w, h, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
