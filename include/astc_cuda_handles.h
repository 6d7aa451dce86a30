// astc_cuda_handles.h -- the CUDA-side stand-ins for the D3D11 objects the
// reference passes around (ID3D11Device / ID3D11DeviceContext /
// ID3D11Texture2D / ID3D11Buffer in astc_encode.h:87 and astc_save.h:34).
// Plain structs over the C ABI of astc_b200.h; host code needs no CUDA headers.
#pragma once
#include <cstddef>
#include <cstdint>

#include "astc_b200.h"

// ID3D11Device: which GPU.
struct astc_device {
    int ordinal = 0;
};

// ID3D11DeviceContext: where work is queued (a cudaStream_t, nullptr = default).
struct astc_context {
    void *stream = nullptr;
};

// ID3D11Texture2D with its D3D11_TEXTURE2D_DESC (main.cpp:33-52).
struct astc_texture2d {
    uint8_t *d_rgba = nullptr;     // device memory, RGBA8, row 0 first
    int width = 0, height = 0;
    size_t pitch = 0;
    bool srgb_format = false;      // DXGI_FORMAT_R8G8B8A8_UNORM_SRGB vs _UNORM (main.cpp:38)
};

// ID3D11Buffer with its D3D11_BUFFER_DESC (astc_encode.h:137-146).
struct astc_buffer {
    uint8_t *d_data = nullptr;     // device memory
    uint32_t ByteWidth = 0;        // 16 * TotalBlockNum
    uint32_t StructureByteStride = ASTC_B200_BLOCK_BYTES;
};

inline void release(astc_texture2d *t)
{
    if (t) { astc_b200_free_device(t->d_rgba); delete t; }
}
inline void release(astc_buffer *b)
{
    if (b) { astc_b200_free_device(b->d_data); delete b; }
}
