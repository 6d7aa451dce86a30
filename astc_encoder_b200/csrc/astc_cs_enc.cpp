// astc_cs_enc -- command-line front end with the reference's surface
// (main.cpp:140-258, README.md:27-40):
//
//     astc_cs_enc <input image> [-4x4] [-6x6] [-alpha] [-norm] [-srgb] [-accum]
//
// (-accum is an extension: max_accumulation_pixel_direction, ASTC_Encode.hlsl:170-227, instead of the PCA.)
//
// Same flag spelling and semantics (flags are read from argv[2] on, unknown
// flags are ignored), same stdout lines, same output naming (<input minus its
// last extension>.astc) and exit codes (0 / -1).  The D3D11 device, swap chain,
// texture and UAV of the reference are replaced by the C ABI in astc_b200.h.
// Deviation: -6x6 really selects 6x6 blocks (it is inert in the reference).
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "astc_encode.h"
#include "astc_save.h"

// load_tex (main.cpp:19-56): decode, flip vertically, force RGBA8, upload.
static astc_texture2d *load_tex(astc_device *dev, const char *tex_path, bool bSRGB)
{
    int xsize = 0, ysize = 0, components = 0;
    uint8_t *image = nullptr;
    if (astc_b200_load_image(tex_path, /*flip_vertically=*/1, &xsize, &ysize, &components, &image) != ASTC_B200_OK) {
        std::printf("Failed to load image %s\nReason: %s\n", tex_path, astc_b200_image_failure_reason());
        return nullptr;
    }
    astc_texture2d *tex = new astc_texture2d();
    tex->width = xsize;
    tex->height = ysize;
    tex->pitch = size_t(xsize) * 4;
    tex->srgb_format = bSRGB;
    int rc = astc_b200_set_device(dev->ordinal);
    if (rc == ASTC_B200_OK) rc = astc_b200_malloc_device(reinterpret_cast<void **>(&tex->d_rgba), tex->pitch * size_t(ysize));
    if (rc == ASTC_B200_OK) rc = astc_b200_memcpy_h2d(tex->d_rgba, image, tex->pitch * size_t(ysize), nullptr);
    if (rc == ASTC_B200_OK) rc = astc_b200_stream_synchronize(nullptr);
    astc_b200_free_host_buffer(image);
    if (rc != ASTC_B200_OK) {
        release(tex);
        return nullptr;
    }
    return tex;
}

// create_device_swapchain (main.cpp:58-119) reduces to picking a GPU.
static int create_device(astc_device &dev, astc_context &ctx)
{
    int count = 0;
    int rc = astc_b200_device_count(&count);
    if (rc != ASTC_B200_OK) return rc;
    if (count <= 0) return ASTC_B200_ERR_NO_DEVICE;
    dev.ordinal = 0;
    ctx.stream = nullptr;
    return astc_b200_set_device(dev.ordinal);
}

static void strip_file_extension(std::string &file_path)
{
    const std::string::size_type dot = file_path.rfind('.');
    if (dot != std::string::npos) file_path.erase(dot);
}

static bool parse_cmd(int argc, char **argv, encode_option &option)
{
    for (int i = 2; i < argc; ++i) {
        const std::string arg = argv[i];
        bool *target = nullptr;
        if (arg == "-4x4") target = &option.is4x4;
        else if (arg == "-6x6") target = &option.is6x6;
        else if (arg == "-norm") target = &option.is_normal_map;
        else if (arg == "-srgb") target = &option.srgb;
        else if (arg == "-alpha") target = &option.has_alpha;
        else if (arg == "-accum") target = &option.max_accumulation_axis;   // extension: the reference's commented-out axis heuristic
        if (target) *target = true;
    }
    return true;
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::cout << "wrong args count" << std::endl;
        return -1;
    }

    encode_option option;
    if (!parse_cmd(argc, argv, option)) {
        std::cout << "wrong args options" << std::endl;
        return -1;
    }
    const int DimSize = block_dim_of(option);

    std::cout << "encode option setting:\n"
              << "has_alpha\t" << std::boolalpha << option.has_alpha << std::endl
              << "is 4x4 block\t" << (DimSize == 4) << std::endl
              << "normal map\t" << option.is_normal_map << std::endl
              << "encode in gamma color space\t" << option.srgb << std::endl;

    astc_device device;
    astc_context context;
    const int hr = create_device(device, context);
    if (hr != ASTC_B200_OK) {
        std::cout << "init cuda failed! (" << astc_b200_strerror(hr) << ")" << std::endl;
        return hr;
    }

    const std::string src_tex = argv[1];
    astc_texture2d *pSrcTexture = load_tex(&device, src_tex.c_str(), option.srgb && !option.is_normal_map);
    if (pSrcTexture == nullptr) {
        std::cout << "load source texture failed! [" << src_tex << "]" << std::endl;
        return -1;
    }

    astc_buffer *pOutBuf = encode_astc(&device, &context, pSrcTexture, option);
    if (pOutBuf == nullptr) {
        std::cout << "encode astc failed!" << std::endl;
        return -1;
    }

    const uint32_t bufLen = pOutBuf->ByteWidth;
    uint8_t *pMemBuf = new uint8_t[bufLen ? bufLen : 1]();
    if (read_gpu(&device, &context, pOutBuf, pMemBuf, bufLen) != ASTC_B200_OK) {
        std::cout << "save astc failed!" << std::endl;
        return -1;
    }

    std::string dst_tex(src_tex);
    strip_file_extension(dst_tex);
    dst_tex += ".astc";
    save_astc(dst_tex.c_str(), DimSize, DimSize, pSrcTexture->width, pSrcTexture->height, pMemBuf, int(bufLen));

    delete[] pMemBuf;
    release(pOutBuf);
    release(pSrcTexture);

    std::cout << "save astc to:" << dst_tex << std::endl;
    return 0;
}
