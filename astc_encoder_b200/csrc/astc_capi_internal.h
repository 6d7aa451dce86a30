// astc_capi_internal.h -- helpers shared by the C-ABI translation units (astc_capi.cu,
// astc_context.cu).  Not installed.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "astc_b200.h"
#include "astc_kernels.h"

namespace astc_capi {

// records the CUDA error text for astc_b200_last_cuda_error() and maps it to an astc_b200_status
int cuda_fail(cudaError_t e, const char *what);
void count_launch();

inline int dim_of(const astc_b200_option *o) { return (o->is6x6 || !o->is4x4) ? 6 : 4; }

inline uint32_t align_flags(const void *base, size_t pitch)
{
    const uintptr_t b = reinterpret_cast<uintptr_t>(base);
    uint32_t f = 0;
    if (b % 16 == 0 && pitch % 16 == 0) f |= astc::kFlagAligned16;
    if (b % 8 == 0 && pitch % 8 == 0) f |= astc::kFlagAligned8;
    return f;
}

inline astc::ImageDesc make_desc(const uint8_t *rgba, uint8_t *blocks, size_t pitch, int w, int h, int dim, uint64_t first)
{
    astc::ImageDesc d{};
    d.rgba = rgba; d.blocks = blocks; d.pitch = pitch; d.first_block = first;
    d.width = w; d.height = h;
    d.blocks_x = uint32_t((w + dim - 1) / dim);
    d.flags = align_flags(rgba, pitch);
    return d;
}

// Blocks of one texture.
inline uint64_t block_count(int w, int h, int dim) { return uint64_t((w + dim - 1) / dim) * uint64_t((h + dim - 1) / dim); }

// Layout of the levels below the base in one arena: level l+1 at offsets[l] (multiples of 256 bytes), rows tightly
// packed (4 * widths[l] bytes), followed by a 256-byte scratch area for the fused kernel's ticket.
inline int mip_layout(int width, int height, size_t *offsets, int *widths, int *heights, size_t *total_bytes)
{
    int n = 0, w = width, h = height;
    size_t off = 0;
    while ((w > 1 || h > 1) && n < astc::kMaxMipLevels) {
        w = w > 1 ? w / 2 : 1;
        h = h > 1 ? h / 2 : 1;
        if (offsets) offsets[n] = off;
        if (widths) widths[n] = w;
        if (heights) heights[n] = h;
        off += (size_t(w) * size_t(h) * 4u + 255u) / 256u * 256u;
        ++n;
    }
    if (total_bytes) *total_bytes = off + 256u;
    return n;
}

}  // namespace astc_capi

#define CUDA_TRY(expr)                                                     \
    do {                                                                   \
        cudaError_t e__ = (expr);                                          \
        if (e__ != cudaSuccess) return astc_capi::cuda_fail(e__, #expr);   \
    } while (0)
