// image_io.cpp -- host image ingest: the role stb_image plays in the reference
// (main.cpp:19-30: stbi_set_flip_vertically_on_load(1); stbi_load(..., STBI_rgb_alpha)).
// Own decoders: PNG (every colour type / bit depth, tRNS, Adam7; zlib) and PNM (P5 / P6) here, JPEG in
// jpeg_io.cpp, GIF / PSD / HDR / PIC / BMP / TGA in image_formats.cpp -- every format the reference's
// stb_image v2.22 reads, with its results (tests/test_image_formats.py checks against stb itself).
// Output is always 8-bit RGBA; `components_in_file` reports what the file held, as stbi_load's `comp` does.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "astc_b200.h"
#include "image_formats.h"

namespace {
thread_local const char *g_reason = "";
}

namespace astc_image {
bool fail(const char *why) { g_reason = why; return false; }
}  // namespace astc_image

namespace {

using astc_image::fail;
using astc_image::Image;

bool read_file(const char *path, std::vector<uint8_t> &out)
{
    std::FILE *f = std::fopen(path, "rb");
    if (!f) return fail("can't fopen");
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return fail("can't read"); }
    out.resize(size_t(n));
    const size_t got = n ? std::fread(out.data(), 1, size_t(n), f) : 0;
    std::fclose(f);
    return got == size_t(n) ? true : fail("short read");
}

uint32_t be32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
uint32_t le32(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }
uint32_t le16(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8); }

// ---------------------------------------------------------------- PNG -----
int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Reverse the per-scanline filters of one (sub)image in place; returns rows
// without their filter byte.
bool png_unfilter(const uint8_t *src, size_t src_len, int w, int h, int bits_per_pixel, std::vector<uint8_t> &out)
{
    const size_t stride = (size_t(w) * size_t(bits_per_pixel) + 7) / 8;
    const size_t bpp = size_t(bits_per_pixel + 7) / 8;
    if (src_len < (stride + 1) * size_t(h)) return fail("not enough pixels");
    out.assign(stride * size_t(h), 0);
    for (int y = 0; y < h; ++y) {
        const uint8_t *in = src + (stride + 1) * size_t(y);
        uint8_t *cur = out.data() + stride * size_t(y);
        const uint8_t *up = y ? cur - stride : nullptr;
        const int ft = in[0];
        ++in;
        if (ft > 4) return fail("invalid filter");
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0;
            const int b = up ? up[i] : 0;
            const int c = (up && i >= bpp) ? up[i - bpp] : 0;
            int v = in[i];
            switch (ft) {
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: break;
            }
            cur[i] = uint8_t(v);
        }
    }
    return true;
}

struct PngInfo {
    int w, h, depth, ctype, interlace;
    uint8_t palette[256][4];
    int palette_len = 0;
    bool has_key = false;
    uint16_t key[3] = {0, 0, 0};
};

int png_channels(int ctype) { return ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4; }

// Expand unfiltered rows of a (sub)image to RGBA8 pixels at (x0 + x*dx, y0 + y*dy).
void png_expand(const PngInfo &pi, const std::vector<uint8_t> &rows, int w, int h, int x0, int y0, int dx, int dy, Image &img)
{
    const int ch = png_channels(pi.ctype);
    const size_t stride = (size_t(w) * size_t(ch * pi.depth) + 7) / 8;
    const int scale = pi.depth == 1 ? 255 : pi.depth == 2 ? 85 : pi.depth == 4 ? 17 : 1;
    for (int y = 0; y < h; ++y) {
        const uint8_t *row = rows.data() + stride * size_t(y);
        for (int x = 0; x < w; ++x) {
            uint16_t s[4] = {0, 0, 0, 0};
            for (int c = 0; c < ch; ++c) {
                const size_t idx = size_t(x) * size_t(ch) + size_t(c);
                if (pi.depth == 8) s[c] = row[idx];
                else if (pi.depth == 16) s[c] = uint16_t((row[2 * idx] << 8) | row[2 * idx + 1]);
                else {
                    const size_t bit = idx * size_t(pi.depth);
                    s[c] = uint16_t((row[bit >> 3] >> (8 - pi.depth - int(bit & 7))) & ((1 << pi.depth) - 1));
                }
            }
            uint8_t px[4] = {0, 0, 0, 255};
            auto to8 = [&](uint16_t v) -> uint8_t { return pi.depth == 16 ? uint8_t(v >> 8) : uint8_t(v * scale); };
            switch (pi.ctype) {
            case 0:
                px[0] = px[1] = px[2] = to8(s[0]);
                if (pi.has_key && s[0] == pi.key[0]) px[3] = 0;
                break;
            case 2:
                px[0] = to8(s[0]); px[1] = to8(s[1]); px[2] = to8(s[2]);
                if (pi.has_key && s[0] == pi.key[0] && s[1] == pi.key[1] && s[2] == pi.key[2]) px[3] = 0;
                break;
            case 3: {
                const int i = s[0] < pi.palette_len ? s[0] : 0;
                std::memcpy(px, pi.palette[i], 4);
                break;
            }
            case 4:
                px[0] = px[1] = px[2] = to8(s[0]); px[3] = to8(s[1]);
                break;
            default:
                px[0] = to8(s[0]); px[1] = to8(s[1]); px[2] = to8(s[2]); px[3] = to8(s[3]);
                break;
            }
            std::memcpy(&img.rgba[(size_t(y0 + y * dy) * size_t(img.w) + size_t(x0 + x * dx)) * 4], px, 4);
        }
    }
}

bool decode_png(const std::vector<uint8_t> &file, Image &img)
{
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) return fail("bad png sig");
    PngInfo pi{};
    std::vector<uint8_t> idat;
    bool have_ihdr = false, have_trns = false;
    size_t pos = 8;
    for (;;) {
        if (pos + 12 > file.size()) return fail("truncated png");
        const uint32_t len = be32(&file[pos]);
        const uint8_t *type = &file[pos + 4], *data = &file[pos + 8];
        if (len > file.size() || pos + 12 + size_t(len) > file.size()) return fail("truncated png chunk");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) return fail("bad IHDR len");
            pi.w = int(be32(data)); pi.h = int(be32(data + 4));
            pi.depth = data[8]; pi.ctype = data[9]; pi.interlace = data[12];
            if (pi.w <= 0 || pi.h <= 0 || pi.w > (1 << 24) || pi.h > (1 << 24)) return fail("too large");
            if (data[10] || data[11] || pi.interlace > 1) return fail("bad png method");
            const bool depth_ok = pi.depth == 1 || pi.depth == 2 || pi.depth == 4 || pi.depth == 8 || pi.depth == 16;
            const bool ctype_ok = pi.ctype == 0 || pi.ctype == 2 || pi.ctype == 3 || pi.ctype == 4 || pi.ctype == 6;
            if (!depth_ok || !ctype_ok) return fail("bad ctype");
            if ((pi.ctype == 3 && pi.depth == 16) || ((pi.ctype == 2 || pi.ctype == 4 || pi.ctype == 6) && pi.depth < 8))
                return fail("bad depth");
            have_ihdr = true;
        } else if (!have_ihdr) {
            return fail("first not IHDR");
        } else if (!std::memcmp(type, "PLTE", 4)) {
            if (len > 768 || len % 3) return fail("invalid PLTE");
            pi.palette_len = int(len / 3);
            for (int i = 0; i < pi.palette_len; ++i) {
                pi.palette[i][0] = data[3 * i]; pi.palette[i][1] = data[3 * i + 1];
                pi.palette[i][2] = data[3 * i + 2]; pi.palette[i][3] = 255;
            }
        } else if (!std::memcmp(type, "tRNS", 4)) {
            have_trns = true;
            if (pi.ctype == 3) {
                for (uint32_t i = 0; i < len && int(i) < pi.palette_len; ++i) pi.palette[i][3] = data[i];
            } else if (pi.ctype == 0 && len >= 2) {
                pi.has_key = true; pi.key[0] = uint16_t((data[0] << 8) | data[1]);
            } else if (pi.ctype == 2 && len >= 6) {
                pi.has_key = true;
                for (int c = 0; c < 3; ++c) pi.key[c] = uint16_t((data[2 * c] << 8) | data[2 * c + 1]);
            }
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + size_t(len);
    }
    if (!have_ihdr || idat.empty()) return fail("no IDAT");

    const int ch = png_channels(pi.ctype), bpp_bits = ch * pi.depth;
    // inflated size: every (sub)image row carries a filter byte
    size_t raw_len = 0;
    static const int ax0[7] = {0, 4, 0, 2, 0, 1, 0}, ay0[7] = {0, 0, 4, 0, 2, 0, 1};
    static const int adx[7] = {8, 8, 4, 4, 2, 2, 1}, ady[7] = {8, 8, 8, 4, 4, 2, 2};
    if (pi.interlace) {
        for (int p = 0; p < 7; ++p) {
            const int pw = (pi.w - ax0[p] + adx[p] - 1) / adx[p], ph = (pi.h - ay0[p] + ady[p] - 1) / ady[p];
            if (pw > 0 && ph > 0) raw_len += ((size_t(pw) * size_t(bpp_bits) + 7) / 8 + 1) * size_t(ph);
        }
    } else {
        raw_len = ((size_t(pi.w) * size_t(bpp_bits) + 7) / 8 + 1) * size_t(pi.h);
    }
    // deflate expands at most ~1032 : 1, so a header that promises more pixels than the IDAT data can hold is
    // rejected before anything of that size is allocated (a 30 000 x 30 000 IHDR on a 4 KB file)
    if (raw_len / 1040u > idat.size() + 64u) return fail("not enough pixels");
    std::vector<uint8_t> raw(raw_len);
    uLongf got = uLongf(raw_len);
    const int zrc = uncompress(raw.data(), &got, idat.data(), uLong(idat.size()));
    if (zrc != Z_OK && zrc != Z_BUF_ERROR) return fail("bad zlib stream");
    if (size_t(got) < raw_len) return fail("not enough pixels");

    img.w = pi.w; img.h = pi.h;
    img.comp = pi.ctype == 3 ? (have_trns ? 4 : 3) : ch + ((have_trns && (pi.ctype == 0 || pi.ctype == 2)) ? 1 : 0);
    img.rgba.assign(size_t(pi.w) * size_t(pi.h) * 4, 0);
    std::vector<uint8_t> rows;
    if (!pi.interlace) {
        if (!png_unfilter(raw.data(), raw.size(), pi.w, pi.h, bpp_bits, rows)) return false;
        png_expand(pi, rows, pi.w, pi.h, 0, 0, 1, 1, img);
    } else {
        size_t off = 0;
        for (int p = 0; p < 7; ++p) {
            const int pw = (pi.w - ax0[p] + adx[p] - 1) / adx[p], ph = (pi.h - ay0[p] + ady[p] - 1) / ady[p];
            if (pw <= 0 || ph <= 0) continue;
            if (!png_unfilter(raw.data() + off, raw.size() - off, pw, ph, bpp_bits, rows)) return false;
            png_expand(pi, rows, pw, ph, ax0[p], ay0[p], adx[p], ady[p], img);
            off += ((size_t(pw) * size_t(bpp_bits) + 7) / 8 + 1) * size_t(ph);
        }
    }
    return true;
}

// ---------------------------------------------------------------- PNM -----
bool decode_pnm(const std::vector<uint8_t> &f, Image &img)
{
    if (f.size() < 3 || f[0] != 'P' || (f[1] != '5' && f[1] != '6')) return fail("not PNM");
    size_t pos = 2;
    int vals[3];
    for (int k = 0; k < 3; ++k) {
        for (;;) {
            while (pos < f.size() && (f[pos] == ' ' || f[pos] == '\n' || f[pos] == '\r' || f[pos] == '\t')) ++pos;
            if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') ++pos; continue; }
            break;
        }
        int v = 0, digits = 0;
        while (pos < f.size() && f[pos] >= '0' && f[pos] <= '9') { v = v * 10 + (f[pos++] - '0'); ++digits; }
        if (!digits) return fail("bad PNM header");
        vals[k] = v;
    }
    ++pos;
    const int ch = f[1] == '6' ? 3 : 1;
    if (vals[0] <= 0 || vals[1] <= 0 || vals[2] != 255) return fail("unsupported PNM");
    if (pos + size_t(vals[0]) * size_t(vals[1]) * size_t(ch) > f.size()) return fail("truncated PNM");
    img.w = vals[0]; img.h = vals[1]; img.comp = ch;
    img.rgba.resize(size_t(img.w) * size_t(img.h) * 4);
    for (size_t i = 0; i < size_t(img.w) * size_t(img.h); ++i) {
        const uint8_t *p = &f[pos + i * size_t(ch)];
        uint8_t *o = &img.rgba[i * 4];
        o[0] = p[0]; o[1] = ch == 3 ? p[1] : p[0]; o[2] = ch == 3 ? p[2] : p[0]; o[3] = 255;
    }
    return true;
}

bool ends_with_ci(const std::string &s, const char *ext)
{
    const size_t n = std::strlen(ext);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (std::tolower((unsigned char)s[s.size() - n + i]) != ext[i]) return false;
    return true;
}

}  // namespace

extern "C" {

int astc_b200_load_image(const char *path, int flip_vertically, int *width, int *height, int *components_in_file,
                         uint8_t **rgba)
{
    if (!path || !width || !height || !rgba) return ASTC_B200_ERR_INVALID_ARGUMENT;
    *rgba = nullptr;
    g_reason = "";
    std::vector<uint8_t> file;
    if (!read_file(path, file)) return ASTC_B200_ERR_IO;
    Image img;
    bool ok;
    // format sniffing in stb_image's order (stbi__load_main): JPEG, PNG, BMP, GIF, PSD, PIC, PNM, HDR; TGA, which has
    // no magic number, last
    auto starts = [&](const char *magic, size_t n, size_t at = 0) { return file.size() >= at + n && std::memcmp(file.data() + at, magic, n) == 0; };
    try {
        if (file.size() >= 3 && file[0] == 0xFF && file[1] == 0xD8) ok = astc_image::decode_jpeg(file, img);
        else if (file.size() >= 8 && file[0] == 137 && file[1] == 'P') ok = decode_png(file, img);
        else if (starts("BM", 2)) ok = astc_image::decode_bmp(file, img);
        else if (starts("GIF8", 4)) ok = astc_image::decode_gif(file, img);
        else if (starts("8BPS", 4)) ok = astc_image::decode_psd(file, img);
        else if (starts("\x53\x80\xF6\x34", 4) && starts("PICT", 4, 88)) ok = astc_image::decode_pic(file, img);
        else if (file.size() >= 2 && file[0] == 'P' && (file[1] == '5' || file[1] == '6')) ok = decode_pnm(file, img);
        else if (starts("#?RADIANCE\n", 11) || starts("#?RGBE\n", 7)) ok = astc_image::decode_hdr(file, img);
        else if (astc_image::looks_like_tga(file)) ok = astc_image::decode_tga(file, img);
        else ok = fail("unknown image type");
    } catch (const std::bad_alloc &) {                       // a crafted header must not throw through the C ABI
        g_reason = "outofmem";
        return ASTC_B200_ERR_OUT_OF_MEMORY;
    }
    if (!ok) return ASTC_B200_ERR_BAD_IMAGE;

    const size_t row = size_t(img.w) * 4;
    uint8_t *out = static_cast<uint8_t *>(std::malloc(row * size_t(img.h)));
    if (!out) { g_reason = "outofmem"; return ASTC_B200_ERR_OUT_OF_MEMORY; }
    for (int y = 0; y < img.h; ++y)
        std::memcpy(out + row * size_t(y), &img.rgba[row * size_t(flip_vertically ? img.h - 1 - y : y)], row);
    *width = img.w; *height = img.h;
    if (components_in_file) *components_in_file = img.comp;
    *rgba = out;
    return ASTC_B200_OK;
}

const char *astc_b200_image_failure_reason(void) { return g_reason; }

}  // extern "C"
