// Does the texture unit's UNORM8 -> float conversion (cudaReadModeNormalizedFloat) equal the
// correctly rounded c / 255.0f for every byte?  And what does its sRGB decode return?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o texunorm texunorm.cu && ./texunorm
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>

__global__ void fetch(cudaTextureObject_t t, cudaTextureObject_t ts, float4 *out, float4 *outs)
{
    const int i = threadIdx.x;
    out[i] = tex2D<float4>(t, i + 0.5f, 0.5f);
    outs[i] = tex2D<float4>(ts, i + 0.5f, 0.5f);
}

int main()
{
    uint8_t h[256 * 4];
    for (int i = 0; i < 256; ++i) h[4 * i] = h[4 * i + 1] = h[4 * i + 2] = h[4 * i + 3] = uint8_t(i);
    uint8_t *d; size_t pitch;
    cudaMallocPitch(&d, &pitch, 256 * 4, 4);
    for (int r = 0; r < 4; ++r) cudaMemcpy(d + r * pitch, h, sizeof h, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = d; rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
    rd.res.pitch2D.width = 256; rd.res.pitch2D.height = 4; rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
    cudaTextureObject_t t, ts;
    printf("create: %s\n", cudaGetErrorString(cudaCreateTextureObject(&t, &rd, &td, nullptr)));
    td.sRGB = 1;
    printf("create srgb: %s\n", cudaGetErrorString(cudaCreateTextureObject(&ts, &rd, &td, nullptr)));
    float4 *o, *os; cudaMalloc(&o, 256 * 16); cudaMalloc(&os, 256 * 16);
    fetch<<<1, 256>>>(t, ts, o, os);
    printf("run: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    float4 ho[256], hs[256];
    cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost); cudaMemcpy(hs, os, sizeof hs, cudaMemcpyDeviceToHost);
    int bad = 0, bads = 0, bada = 0;
    for (int i = 0; i < 256; ++i) {
        const float want = float(i) / 255.0f;
        if (ho[i].x != want || ho[i].y != want || ho[i].z != want || ho[i].w != want) { if (bad < 8) printf("unorm %d: got %.9g want %.9g\n", i, ho[i].x, want); ++bad; }
        const double c = i / 255.0;
        const double lin = c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4);
        const float wants = float(lin);
        if (hs[i].x != wants) { if (bads < 8) printf("srgb %d: got %.9g want %.9g (rel %.3g)\n", i, hs[i].x, wants, (hs[i].x - wants) / (wants + 1e-30)); ++bads; }
        if (hs[i].w != want) ++bada;
    }
    printf("UNORM8 mismatches vs c/255.0f: %d of 256; sRGB mismatches vs D3D formula: %d of 256; sRGB-mode alpha mismatches: %d\n", bad, bads, bada);
    return 0;
}
