// astc_encode.h -- the reference's entry point, re-hosted on CUDA.
//
//   encode_option   reference astc_encode.h:14-28  (same fields, order, defaults)
//   encode_astc()   reference astc_encode.h:87-194 (device, context, texture, option
//                   -> device-resident output buffer, nullptr on failure; the
//                   launch is asynchronous, read_gpu() synchronises)
// The D3D11 objects become the plain structs of astc_cuda_handles.h; shader
// compilation, SRV/UAV/constant-buffer binding and Dispatch collapse into one
// C-ABI call, astc_b200_encode_device().
#pragma once
#include <iostream>

#include "astc_cuda_handles.h"

#define BLOCK_BYTES ASTC_B200_BLOCK_BYTES

struct encode_option {
    bool is4x4 = true;
    bool is6x6 = false;
    bool is_normal_map = false;
    bool has_alpha = false;
    bool srgb = false;
    // Extension (after the reference's fields, so aggregate use of the first five is unchanged): the axis
    // heuristic.  false = principal_component_analysis (ASTC_Encode.hlsl:515, what the reference ships);
    // true = max_accumulation_pixel_direction (:170-227, its call is commented out at :514).  CLI: -accum.
    bool max_accumulation_axis = false;
};

// Block edge for an option set.  The reference derives it from is4x4 alone
// (astc_encode.h:124), which its command line can never clear, so `-6x6` has no
// effect there; here is6x6 selects 6x6 (documented deviation).
inline int block_dim_of(const encode_option &option) { return (option.is6x6 || !option.is4x4) ? 6 : 4; }

inline astc_b200_option to_abi(const encode_option &option, bool srgb_texture)
{
    astc_b200_option o;
    astc_b200_option_default(&o);
    o.is4x4 = option.is4x4;
    o.is6x6 = option.is6x6;
    o.is_normal_map = option.is_normal_map;
    o.has_alpha = option.has_alpha;
    o.srgb = srgb_texture;          // the sRGB decode belongs to the texture format (main.cpp:38,214)
    o.axis_method = option.max_accumulation_axis ? 1 : 0;
    return o;
}

inline astc_buffer *encode_astc(astc_device *pDevice, astc_context *pContext, astc_texture2d *pSrcTexture,
                                const encode_option &option)
{
    if (!pSrcTexture) return nullptr;
    if (pDevice && astc_b200_set_device(pDevice->ordinal) != ASTC_B200_OK) return nullptr;
    const astc_b200_option abi = to_abi(option, pSrcTexture->srgb_format);

    astc_buffer *out = new astc_buffer();
    out->ByteWidth = uint32_t(astc_b200_output_size(pSrcTexture->width, pSrcTexture->height, &abi));
    if (astc_b200_malloc_device(reinterpret_cast<void **>(&out->d_data), out->ByteWidth) != ASTC_B200_OK) {
        delete out;
        return nullptr;
    }
    const int rc = astc_b200_encode_device(pSrcTexture->d_rgba, pSrcTexture->width, pSrcTexture->height,
                                           pSrcTexture->pitch, &abi, out->d_data, pContext ? pContext->stream : nullptr);
    if (rc != ASTC_B200_OK) {
        std::cout << "encode kernel failed: " << astc_b200_strerror(rc) << " " << astc_b200_last_cuda_error() << std::endl;
        release(out);
        return nullptr;
    }
    return out;
}
