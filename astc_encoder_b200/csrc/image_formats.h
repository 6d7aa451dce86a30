// image_formats.h -- internal interface between astc_b200_load_image (image_io.cpp) and the per-format
// decoders.  Every decoder fills an 8-bit RGBA image, top row first, and reports the channel count the
// file held the way stbi_load's `comp` does (the reference calls stbi_load(..., STBI_rgb_alpha), main.cpp:25).
#pragma once
#include <cstdint>
#include <vector>

namespace astc_image {

struct Image {
    int w = 0, h = 0, comp = 0;
    std::vector<uint8_t> rgba;
};

bool fail(const char *why);            // records the reason (astc_b200_image_failure_reason) and returns false

bool decode_jpeg(const std::vector<uint8_t> &file, Image &img);      // jpeg_io.cpp
bool decode_gif(const std::vector<uint8_t> &file, Image &img);       // image_formats.cpp
bool decode_psd(const std::vector<uint8_t> &file, Image &img);
bool decode_hdr(const std::vector<uint8_t> &file, Image &img);
bool decode_pic(const std::vector<uint8_t> &file, Image &img);
bool decode_bmp(const std::vector<uint8_t> &file, Image &img);
bool decode_tga(const std::vector<uint8_t> &file, Image &img);
bool looks_like_tga(const std::vector<uint8_t> &file);               // TGA has no magic: header plausibility, tried last

}  // namespace astc_image
