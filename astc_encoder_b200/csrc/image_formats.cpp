// image_formats.cpp -- the remaining input formats of the reference's loader (stb_image v2.22 behind
// stbi_load(..., STBI_rgb_alpha), main.cpp:24-25): GIF (first frame), Photoshop PSD (RGB, 8 / 16 bit,
// raw or RLE), Radiance HDR (tone-mapped to 8 bit the way stbi_load does: gamma 2.2, scale 1),
// Softimage PIC, BMP (1 / 4 / 8-bit palettes, 16 / 24 / 32 bit, bit-field masks) and TGA (true colour,
// 15 / 16-bit, grey, grey + alpha, colour-mapped, raw or RLE, either origin).
//
// Independent implementations that keep stb_image's OBSERVABLE results, quirks included, because the
// texels the encoder sees must be the texels the reference would have seen:
//   * GIF: pixels the first frame does not cover take the background colour with red and blue swapped
//     (stb copies its BGR palette entry verbatim), transparent pixels stay (0, 0, 0, 0);
//   * PSD: colour is un-multiplied from a white matte where 0 < alpha < 255;
//   * BMP: a 32-bit file whose alpha bytes are all zero comes back opaque;
//   * TGA: 5-bit channels expand as (v * 255) / 31.
// Checked against stb_image itself in the build container (oracle/_ref/libstb_ref.so) and against
// committed fixtures made with it (tests/test_image_formats.py, tools/make_image_fixtures.py).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "image_formats.h"

namespace astc_image {
namespace {

struct Reader {
    const uint8_t *p, *end;
    explicit Reader(const std::vector<uint8_t> &f) : p(f.data()), end(f.data() + f.size()) {}
    bool eof() const { return p >= end; }
    int u8() { return p < end ? *p++ : 0; }
    int le16() { const int a = u8(); return a | (u8() << 8); }
    uint32_t le32() { const uint32_t a = uint32_t(le16()); return a | (uint32_t(le16()) << 16); }
    int be16() { const int a = u8(); return (a << 8) | u8(); }
    uint32_t be32() { const uint32_t a = uint32_t(be16()); return (a << 16) | uint32_t(be16()); }
    void skip(int64_t n) { p = (n < 0 || n > end - p) ? end : p + n; }
    bool getn(uint8_t *dst, size_t n)
    {
        if (size_t(end - p) < n) { const size_t have = size_t(end - p); std::memcpy(dst, p, have); std::memset(dst + have, 0, n - have); p = end; return false; }
        std::memcpy(dst, p, n);
        p += n;
        return true;
    }
};

bool size_ok(int64_t w, int64_t h) { return w > 0 && h > 0 && w * h <= (int64_t(1) << 28); }

inline uint8_t luma(int r, int g, int b) { return uint8_t((r * 77 + g * 150 + 29 * b) >> 8); }

// n-channel pixels -> RGBA the way stbi__convert_format(..., 4) does
void expand_to_rgba(const uint8_t *src, int comp, size_t pixels, uint8_t *dst)
{
    for (size_t i = 0; i < pixels; ++i, src += comp, dst += 4) {
        switch (comp) {
        case 1: dst[0] = dst[1] = dst[2] = src[0]; dst[3] = 255; break;
        case 2: dst[0] = dst[1] = dst[2] = src[0]; dst[3] = src[1]; break;
        case 3: dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = 255; break;
        default: std::memcpy(dst, src, 4); break;
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// GIF (first frame)
// ---------------------------------------------------------------------------------------------
bool decode_gif(const std::vector<uint8_t> &file, Image &img)
{
    Reader s(file);
    if (s.u8() != 'G' || s.u8() != 'I' || s.u8() != 'F' || s.u8() != '8') return fail("not GIF");
    const int version = s.u8();
    if ((version != '7' && version != '9') || s.u8() != 'a') return fail("not GIF");
    const int W = s.le16(), H = s.le16(), flags = s.u8(), bgindex = s.u8();
    s.u8();                                                  // aspect ratio
    if (!size_ok(W, H)) return fail("too large");
    uint8_t pal[256][4] = {}, lpal[256][4] = {};             // stored B, G, R, A like the loader being matched
    auto read_palette = [&](uint8_t (*t)[4], int n, int transparent) {
        for (int i = 0; i < n; ++i) {
            t[i][2] = uint8_t(s.u8()); t[i][1] = uint8_t(s.u8()); t[i][0] = uint8_t(s.u8());
            t[i][3] = transparent == i ? 0 : 255;
        }
    };
    if (flags & 0x80) read_palette(pal, 2 << (flags & 7), -1);
    int transparent = -1, eflags = 0;
    const size_t pcount = size_t(W) * size_t(H);
    img.w = W; img.h = H; img.comp = 4;
    img.rgba.assign(pcount * 4, 0);
    std::vector<uint8_t> touched(pcount, 0);
    for (;;) {
        const int tag = s.u8();
        if (tag == 0x2C) {                                   // image descriptor
            const int x = s.le16(), y = s.le16(), w = s.le16(), h = s.le16();
            if (x + w > W || y + h > H) return fail("bad Image Descriptor");
            const int lflags = s.u8();
            const uint8_t(*table)[4];
            if (lflags & 0x80) {
                read_palette(lpal, 2 << (lflags & 7), (eflags & 1) ? transparent : -1);
                table = lpal;
            } else if (flags & 0x80) {
                table = pal;
            } else {
                return fail("missing color table");
            }
            // raster walk: rows of the frame rectangle, interlaced in the four GIF passes (8, 8, 4, 2)
            const int line = W * 4, start_x = x * 4, start_y = y * line, max_x = start_x + w * 4, max_y = start_y + h * line;
            int cur_x = start_x, cur_y = w == 0 ? max_y : start_y, step = (lflags & 0x40) ? 8 * line : line, parse = (lflags & 0x40) ? 3 : 0;
            auto emit = [&](int index) {
                if (cur_y >= max_y) return;
                const int idx = cur_x + cur_y;
                touched[size_t(idx) / 4] = 1;
                const uint8_t *c = table[index];
                if (c[3] > 128) {                            // transparent pixels are not drawn
                    uint8_t *o = &img.rgba[size_t(idx)];
                    o[0] = c[2]; o[1] = c[1]; o[2] = c[0]; o[3] = c[3];
                }
                cur_x += 4;
                if (cur_x >= max_x) {
                    cur_x = start_x;
                    cur_y += step;
                    while (cur_y >= max_y && parse > 0) {
                        step = (1 << parse) * line;
                        cur_y = start_y + (step >> 1);
                        --parse;
                    }
                }
            };
            // LZW
            const int lzw_cs = s.u8();
            if (lzw_cs > 12) return fail("bad LZW code size");
            struct Code { int16_t prefix; uint8_t first, suffix; };
            std::vector<Code> codes(8192);
            const int clear = 1 << lzw_cs;
            for (int i = 0; i < clear; ++i) codes[i] = Code{-1, uint8_t(i), uint8_t(i)};
            int codesize = lzw_cs + 1, codemask = (1 << codesize) - 1, avail = clear + 2, oldcode = -1, len = 0;
            uint32_t bits = 0;
            int valid = 0;
            bool first = true, done = false;
            std::vector<uint8_t> stack;
            while (!done) {
                if (valid < codesize) {
                    if (len == 0) {
                        len = s.u8();
                        if (len == 0) break;                 // block terminator: end of the raster
                    }
                    --len;
                    bits |= uint32_t(s.u8()) << valid;
                    valid += 8;
                    continue;
                }
                const int code = int(bits & uint32_t(codemask));
                bits >>= codesize;
                valid -= codesize;
                if (code == clear) {
                    codesize = lzw_cs + 1; codemask = (1 << codesize) - 1; avail = clear + 2; oldcode = -1; first = false;
                } else if (code == clear + 1) {              // end of information
                    s.skip(len);
                    while ((len = s.u8()) > 0) s.skip(len);
                    done = true;
                } else if (code <= avail) {
                    if (first) return fail("no clear code");
                    if (oldcode >= 0) {
                        if (avail + 1 > 8192) return fail("too many codes");
                        Code &n = codes[size_t(avail++)];
                        n.prefix = int16_t(oldcode);
                        n.first = codes[size_t(oldcode)].first;
                        n.suffix = code == avail ? n.first : codes[size_t(code)].first;
                    } else if (code == avail) {
                        return fail("illegal code in raster");
                    }
                    stack.clear();
                    for (int c = code; c >= 0; c = codes[size_t(c)].prefix) stack.push_back(codes[size_t(c)].suffix);
                    for (size_t i = stack.size(); i-- > 0;) emit(stack[i]);
                    if ((avail & codemask) == 0 && avail <= 0x0FFF) { ++codesize; codemask = (1 << codesize) - 1; }
                    oldcode = code;
                } else {
                    return fail("illegal code in raster");
                }
            }
            if (bgindex > 0) {                               // first frame: untouched pixels take the background entry, bytes as stored (B, G, R)
                for (size_t i = 0; i < pcount; ++i)
                    if (!touched[i]) { uint8_t *o = &img.rgba[i * 4]; o[0] = pal[bgindex][0]; o[1] = pal[bgindex][1]; o[2] = pal[bgindex][2]; o[3] = 255; }
            }
            return true;
        }
        if (tag == 0x21) {                                   // extension
            const int ext = s.u8();
            int len;
            if (ext == 0xF9) {                               // graphic control
                len = s.u8();
                if (len == 4) {
                    eflags = s.u8();
                    s.le16();                                // delay
                    if (transparent >= 0) pal[transparent][3] = 255;
                    if (eflags & 1) {
                        transparent = s.u8();
                        pal[transparent][3] = 0;
                    } else {
                        s.skip(1);
                        transparent = -1;
                    }
                } else {
                    s.skip(len);
                    continue;
                }
            }
            while ((len = s.u8()) != 0) s.skip(len);
            continue;
        }
        if (tag == 0x3B) return fail("no image in GIF");     // trailer before any frame
        return fail("unknown code");
    }
}

// ---------------------------------------------------------------------------------------------
// PSD
// ---------------------------------------------------------------------------------------------
bool decode_psd(const std::vector<uint8_t> &file, Image &img)
{
    Reader s(file);
    if (s.be32() != 0x38425053u) return fail("not PSD");
    if (s.be16() != 1) return fail("wrong version");
    s.skip(6);
    const int channels = s.be16();
    if (channels < 0 || channels > 16) return fail("wrong channel count");
    const int h = int(s.be32()), w = int(s.be32());
    const int depth = s.be16();
    if (depth != 8 && depth != 16) return fail("unsupported bit depth");
    if (s.be16() != 3) return fail("wrong color format");
    s.skip(s.be32());                                        // colour mode data
    s.skip(s.be32());                                        // image resources
    s.skip(s.be32());                                        // layers and masks
    const int compression = s.be16();
    if (compression > 1) return fail("bad compression");
    if (!size_ok(w, h)) return fail("too large");
    const size_t n = size_t(w) * size_t(h);
    img.w = w; img.h = h; img.comp = 4;
    img.rgba.assign(n * 4, 0);
    uint8_t *out = img.rgba.data();
    if (compression) {
        s.skip(int64_t(h) * channels * 2);                   // per-row byte counts: the PackBits stream is self-delimiting
        for (int c = 0; c < 4; ++c) {
            uint8_t *p = out + c;
            if (c >= channels) {
                for (size_t i = 0; i < n; ++i, p += 4) *p = c == 3 ? 255 : 0;
                continue;
            }
            size_t count = 0;
            while (count < n) {
                int len = s.u8();
                if (len == 128) continue;
                if (len < 128) {
                    ++len;
                    if (size_t(len) > n - count) return fail("bad RLE data");
                    count += size_t(len);
                    for (; len; --len, p += 4) *p = uint8_t(s.u8());
                } else {
                    len = 257 - len;
                    if (size_t(len) > n - count) return fail("bad RLE data");
                    const uint8_t v = uint8_t(s.u8());
                    count += size_t(len);
                    for (; len; --len, p += 4) *p = v;
                }
                if (s.eof() && count < n) return fail("bad RLE data");
            }
        }
    } else {
        for (int c = 0; c < 4; ++c) {
            uint8_t *p = out + c;
            if (c >= channels) {
                for (size_t i = 0; i < n; ++i, p += 4) *p = c == 3 ? 255 : 0;
            } else if (depth == 16) {
                for (size_t i = 0; i < n; ++i, p += 4) *p = uint8_t(s.be16() >> 8);
            } else {
                for (size_t i = 0; i < n; ++i, p += 4) *p = uint8_t(s.u8());
            }
        }
    }
    if (channels >= 4) {                                     // remove the white matte
        for (size_t i = 0; i < n; ++i) {
            uint8_t *px = out + 4 * i;
            if (px[3] != 0 && px[3] != 255) {
                const float a = px[3] / 255.0f, ra = 1.0f / a, inv_a = 255.0f * (1 - ra);
                px[0] = (unsigned char)(px[0] * ra + inv_a);
                px[1] = (unsigned char)(px[1] * ra + inv_a);
                px[2] = (unsigned char)(px[2] * ra + inv_a);
            }
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Radiance HDR -> 8 bit
// ---------------------------------------------------------------------------------------------
bool decode_hdr(const std::vector<uint8_t> &file, Image &img)
{
    Reader s(file);
    auto token = [&]() {
        std::string t;
        int c = s.u8();
        while (!s.eof() && c != '\n') {
            t.push_back(char(c));
            if (t.size() == 1023) {
                while (!s.eof() && s.u8() != '\n') {}
                break;
            }
            c = s.u8();
        }
        return t;
    };
    const std::string head = token();
    if (head != "#?RADIANCE" && head != "#?RGBE") return fail("not HDR");
    bool valid = false;
    for (;;) {
        const std::string t = token();
        if (t.empty()) break;
        if (t == "FORMAT=32-bit_rle_rgbe") valid = true;
    }
    if (!valid) return fail("unsupported format");
    const std::string dims = token();
    if (dims.compare(0, 3, "-Y ") != 0) return fail("unsupported data layout");
    char *next = nullptr;
    const int height = int(std::strtol(dims.c_str() + 3, &next, 10));
    while (*next == ' ') ++next;
    if (std::strncmp(next, "+X ", 3) != 0) return fail("unsupported data layout");
    const int width = int(std::strtol(next + 3, nullptr, 10));
    if (!size_ok(width, height)) return fail("too large");
    const size_t n = size_t(width) * size_t(height);
    std::vector<float> rgb(n * 3);
    auto convert = [](float *o, const uint8_t *rgbe) {
        if (rgbe[3] != 0) {
            const float f = float(std::ldexp(1.0f, int(rgbe[3]) - (128 + 8)));
            o[0] = rgbe[0] * f; o[1] = rgbe[1] * f; o[2] = rgbe[2] * f;
        } else {
            o[0] = o[1] = o[2] = 0.0f;
        }
    };
    bool flat = width < 8 || width >= 32768;
    size_t done = 0;                                         // pixels already converted when falling back to flat
    if (!flat) {
        std::vector<uint8_t> scan(size_t(width) * 4);
        for (int j = 0; j < height && !flat; ++j) {
            const int c1 = s.u8(), c2 = s.u8();
            int len = s.u8();
            if (c1 != 2 || c2 != 2 || (len & 0x80)) {
                // not run-length encoded after all: these four bytes are the first pixel, the rest of the file is
                // flat RGBE (the loader being matched restarts at the top of the image here, whatever row it is on)
                const uint8_t first[4] = {uint8_t(c1), uint8_t(c2), uint8_t(len), uint8_t(s.u8())};
                convert(&rgb[0], first);
                done = 1;
                flat = true;
                break;
            }
            len = (len << 8) | s.u8();
            if (len != width) return fail("invalid decoded scanline length");
            for (int k = 0; k < 4; ++k) {
                int i = 0;
                while (i < width) {
                    int count = s.u8();
                    const int left = width - i;
                    if (count > 128) {
                        const uint8_t v = uint8_t(s.u8());
                        count -= 128;
                        if (count > left) return fail("bad RLE data in HDR");
                        for (int z = 0; z < count; ++z) scan[size_t(i++) * 4 + size_t(k)] = v;
                    } else {
                        if (count > left) return fail("bad RLE data in HDR");
                        if (count == 0 && s.eof()) return fail("bad RLE data in HDR");
                        for (int z = 0; z < count; ++z) scan[size_t(i++) * 4 + size_t(k)] = uint8_t(s.u8());
                    }
                }
            }
            for (int i = 0; i < width; ++i) convert(&rgb[(size_t(j) * size_t(width) + size_t(i)) * 3], &scan[size_t(i) * 4]);
        }
    }
    if (flat) {
        for (size_t i = done; i < n; ++i) {
            uint8_t rgbe[4];
            s.getn(rgbe, 4);
            convert(&rgb[i * 3], rgbe);
        }
    }
    // stbi_load on an HDR file: gamma 1 / 2.2, scale 1, alpha = 1 -> 255
    img.w = width; img.h = height; img.comp = 3;
    img.rgba.resize(n * 4);
    const float gamma = 1.0f / 2.2f, scale = 1.0f;
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) {
            float z = float(std::pow(double(rgb[i * 3 + size_t(k)] * scale), double(gamma))) * 255 + 0.5f;
            if (z < 0) z = 0;
            if (z > 255) z = 255;
            img.rgba[i * 4 + size_t(k)] = uint8_t(int(z));
        }
        img.rgba[i * 4 + 3] = 255;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Softimage PIC
// ---------------------------------------------------------------------------------------------
bool decode_pic(const std::vector<uint8_t> &file, Image &img)
{
    Reader s(file);
    s.skip(92);
    const int w = s.be16(), h = s.be16();
    if (s.eof()) return fail("file too short (pic header)");
    if (!size_ok(w, h)) return fail("too large");
    s.be32(); s.be16(); s.be16();                            // ratio, fields, pad
    struct Packet { int size, type, channel; } packets[10];
    int count = 0, act = 0, chained;
    do {
        if (count == 10) return fail("too many packets");
        Packet &p = packets[count++];
        chained = s.u8();
        p.size = s.u8(); p.type = s.u8(); p.channel = s.u8();
        act |= p.channel;
        if (s.eof()) return fail("file too short (reading packets)");
        if (p.size != 8) return fail("packet isn't 8bpp");
    } while (chained);
    img.w = w; img.h = h; img.comp = (act & 0x10) ? 4 : 3;
    img.rgba.assign(size_t(w) * size_t(h) * 4, 0xFF);
    auto readval = [&](int channel, uint8_t *dst) {
        for (int i = 0, mask = 0x80; i < 4; ++i, mask >>= 1)
            if (channel & mask) {
                if (s.eof()) return fail("PIC file too short");
                dst[i] = uint8_t(s.u8());
            }
        return true;
    };
    auto copyval = [](int channel, uint8_t *dst, const uint8_t *src) {
        for (int i = 0, mask = 0x80; i < 4; ++i, mask >>= 1)
            if (channel & mask) dst[i] = src[i];
    };
    for (int y = 0; y < h; ++y)
        for (int k = 0; k < count; ++k) {
            const Packet &p = packets[k];
            uint8_t *dst = &img.rgba[size_t(y) * size_t(w) * 4];
            if (p.type == 0) {
                for (int x = 0; x < w; ++x, dst += 4)
                    if (!readval(p.channel, dst)) return false;
            } else if (p.type == 1) {                        // pure run-length
                int left = w;
                while (left > 0) {
                    int n = s.u8();
                    uint8_t v[4] = {};
                    if (s.eof()) return fail("file too short (pure read count)");
                    if (n > left) n = left;
                    if (!readval(p.channel, v)) return false;
                    for (int i = 0; i < n; ++i, dst += 4) copyval(p.channel, dst, v);
                    left -= n;
                }
            } else if (p.type == 2) {                        // mixed
                int left = w;
                while (left > 0) {
                    int n = s.u8();
                    if (s.eof()) return fail("file too short (mixed read count)");
                    if (n >= 128) {
                        uint8_t v[4] = {};
                        n = n == 128 ? s.be16() : n - 127;
                        if (n > left) return fail("scanline overrun");
                        if (!readval(p.channel, v)) return false;
                        for (int i = 0; i < n; ++i, dst += 4) copyval(p.channel, dst, v);
                    } else {
                        ++n;
                        if (n > left) return fail("scanline overrun");
                        for (int i = 0; i < n; ++i, dst += 4)
                            if (!readval(p.channel, dst)) return false;
                    }
                    left -= n;
                }
            } else {
                return fail("packet has bad compression type");
            }
        }
    return true;
}

// ---------------------------------------------------------------------------------------------
// BMP
// ---------------------------------------------------------------------------------------------
namespace {
int high_bit(uint32_t z)
{
    if (!z) return -1;
    int n = 0;
    while (z >>= 1) ++n;
    return n;
}
int bit_count(uint32_t a)
{
    int n = 0;
    for (; a; a &= a - 1) ++n;
    return n;
}
// a masked field of `bits` bits -> 0..255 by bit replication
int field_to_byte(uint32_t v, int shift, int bits)
{
    static const uint32_t mul[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01};
    static const uint32_t sh[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
    v = shift < 0 ? v << -shift : v >> shift;
    v >>= (8 - bits);
    return int(v * mul[bits]) >> sh[bits];
}
}  // namespace

bool decode_bmp(const std::vector<uint8_t> &file, Image &img)
{
    Reader s(file);
    if (s.u8() != 'B' || s.u8() != 'M') return fail("not BMP");
    s.le32(); s.le16(); s.le16();
    const int offset = int(s.le32()), hsz = int(s.le32());
    if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) return fail("unknown BMP");
    int w, hraw;
    if (hsz == 12) { w = s.le16(); hraw = s.le16(); } else { w = int(s.le32()); hraw = int(s.le32()); }
    if (s.le16() != 1) return fail("bad BMP");
    const int bpp = s.le16();
    uint32_t mr = 0, mg = 0, mb = 0, ma = 0, all_a = 255;
    if (hsz != 12) {
        const int compress = int(s.le32());
        if (compress == 1 || compress == 2) return fail("BMP RLE");
        s.le32(); s.le32(); s.le32(); s.le32(); s.le32();
        if (hsz == 40 || hsz == 56) {
            if (hsz == 56) { s.le32(); s.le32(); s.le32(); s.le32(); }
            if (bpp == 16 || bpp == 32) {
                if (compress == 0) {
                    if (bpp == 32) { mr = 0xffu << 16; mg = 0xffu << 8; mb = 0xffu; ma = 0xffu << 24; all_a = 0; }
                    else { mr = 31u << 10; mg = 31u << 5; mb = 31u; }
                } else if (compress == 3) {
                    mr = s.le32(); mg = s.le32(); mb = s.le32();
                    if (mr == mg && mg == mb) return fail("bad BMP");
                } else {
                    return fail("bad BMP");
                }
            }
        } else {
            mr = s.le32(); mg = s.le32(); mb = s.le32(); ma = s.le32();
            s.le32();
            for (int i = 0; i < 12; ++i) s.le32();
            if (hsz == 124) { s.le32(); s.le32(); s.le32(); s.le32(); }
        }
    }
    const bool bottom_up = hraw > 0;
    const int h = std::abs(hraw);
    if (!size_ok(w, h)) return fail("too large");
    int psize = 0;
    if (hsz == 12) { if (bpp < 24) psize = (offset - 14 - 24) / 3; }
    else if (bpp < 16) psize = (offset - 14 - hsz) >> 2;
    img.w = w; img.h = h; img.comp = ma ? 4 : 3;
    img.rgba.assign(size_t(w) * size_t(h) * 4, 0);
    uint8_t *out = img.rgba.data();
    size_t z = 0;
    if (bpp < 16) {
        if (psize <= 0 || psize > 256) return fail("invalid");       // (a data offset inside the header makes it negative)
        uint8_t pal[256][4];
        for (int i = 0; i < psize; ++i) {
            pal[i][2] = uint8_t(s.u8()); pal[i][1] = uint8_t(s.u8()); pal[i][0] = uint8_t(s.u8());
            if (hsz != 12) s.u8();
            pal[i][3] = 255;
        }
        for (int i = psize; i < 256; ++i) pal[i][0] = pal[i][1] = pal[i][2] = 0, pal[i][3] = 255;
        s.skip(offset - 14 - hsz - psize * (hsz == 12 ? 3 : 4));
        int width;
        if (bpp == 1) width = (w + 7) >> 3;
        else if (bpp == 4) width = (w + 1) >> 1;
        else if (bpp == 8) width = w;
        else return fail("bad bpp");
        const int pad = (-width) & 3;
        auto put = [&](int c) { out[z++] = pal[c][0]; out[z++] = pal[c][1]; out[z++] = pal[c][2]; out[z++] = 255; };
        for (int j = 0; j < h; ++j) {
            if (bpp == 1) {
                int bit = 7, v = s.u8();
                for (int i = 0; i < w; ++i) {
                    put((v >> bit) & 1);
                    if (i + 1 == w) break;
                    if (--bit < 0) { bit = 7; v = s.u8(); }
                }
            } else {
                for (int i = 0; i < w; i += 2) {
                    int v = s.u8(), v2 = 0;
                    if (bpp == 4) { v2 = v & 15; v >>= 4; }
                    put(v);
                    if (i + 1 == w) break;
                    put(bpp == 8 ? s.u8() : v2);
                }
            }
            s.skip(pad);
        }
    } else {
        s.skip(offset - 14 - hsz);
        const int width = bpp == 24 ? 3 * w : bpp == 16 ? 2 * w : 0, pad = (-width) & 3;
        int easy = 0;
        if (bpp == 24) easy = 1;
        else if (bpp == 32 && mb == 0xff && mg == 0xff00 && mr == 0x00ff0000 && ma == 0xff000000) easy = 2;
        int rs = 0, gs = 0, bs = 0, as = 0, rc = 0, gc = 0, bc = 0, ac = 0;
        if (!easy) {
            if (bpp != 16 && bpp != 32) return fail("bad bpp");
            if (!mr || !mg || !mb) return fail("bad masks");
            rs = high_bit(mr) - 7; rc = bit_count(mr);
            gs = high_bit(mg) - 7; gc = bit_count(mg);
            bs = high_bit(mb) - 7; bc = bit_count(mb);
            as = high_bit(ma) - 7; ac = bit_count(ma);
            if (rc > 8 || gc > 8 || bc > 8 || ac > 8) return fail("bad masks");
        }
        for (int j = 0; j < h; ++j) {
            for (int i = 0; i < w; ++i) {
                if (easy) {
                    out[z + 2] = uint8_t(s.u8()); out[z + 1] = uint8_t(s.u8()); out[z] = uint8_t(s.u8());
                    const uint8_t a = easy == 2 ? uint8_t(s.u8()) : 255;
                    all_a |= a;
                    out[z + 3] = a;
                    z += 4;
                } else {
                    const uint32_t v = bpp == 16 ? uint32_t(s.le16()) : s.le32();
                    out[z++] = uint8_t(field_to_byte(v & mr, rs, rc));
                    out[z++] = uint8_t(field_to_byte(v & mg, gs, gc));
                    out[z++] = uint8_t(field_to_byte(v & mb, bs, bc));
                    const uint32_t a = ma ? uint32_t(field_to_byte(v & ma, as, ac)) : 255u;
                    all_a |= a;
                    out[z++] = uint8_t(a);
                }
            }
            s.skip(pad);
        }
    }
    if (all_a == 0)
        for (size_t i = 3; i < img.rgba.size(); i += 4) out[i] = 255;
    if (bottom_up) {
        const size_t row = size_t(w) * 4;
        std::vector<uint8_t> tmp(row);
        for (int j = 0; j < h / 2; ++j) {
            uint8_t *a = out + row * size_t(j), *b = out + row * size_t(h - 1 - j);
            std::memcpy(tmp.data(), a, row); std::memcpy(a, b, row); std::memcpy(b, tmp.data(), row);
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// TGA
// ---------------------------------------------------------------------------------------------
bool looks_like_tga(const std::vector<uint8_t> &file)
{
    if (file.size() < 18) return false;
    const uint8_t *f = file.data();
    const int color_type = f[1], type = f[2];
    if (color_type > 1) return false;
    if (color_type == 1) {
        if (type != 1 && type != 9) return false;
        const int pb = f[7];
        if (pb != 8 && pb != 15 && pb != 16 && pb != 24 && pb != 32) return false;
    } else if (type != 2 && type != 3 && type != 10 && type != 11) {
        return false;
    }
    if ((f[12] | (f[13] << 8)) < 1 || (f[14] | (f[15] << 8)) < 1) return false;
    const int bpp = f[16];
    if (color_type == 1 && bpp != 8 && bpp != 16) return false;
    return bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32;
}

bool decode_tga(const std::vector<uint8_t> &file, Image &img)
{
    Reader s(file);
    const int id_len = s.u8(), indexed = s.u8();
    int type = s.u8();
    const int pal_start = s.le16(), pal_len = s.le16(), pal_bits = s.u8();
    s.le16(); s.le16();                                      // origin
    const int w = s.le16(), h = s.le16(), bpp = s.u8();
    const bool top_down = (s.u8() >> 5) & 1;
    const bool rle = type >= 8;
    if (rle) type -= 8;
    auto comp_of = [](int bits, bool grey, bool &rgb16) {
        rgb16 = false;
        switch (bits) {
        case 8: return 1;
        case 16: if (grey) return 2;   // fall through
        case 15: rgb16 = true; return 3;
        case 24: return 3;
        case 32: return 4;
        default: return 0;
        }
    };
    bool rgb16 = false;
    const int comp = indexed ? comp_of(pal_bits, false, rgb16) : comp_of(bpp, type == 3, rgb16);
    if (!comp) return fail("bad format");
    if (!size_ok(w, h)) return fail("too large");
    const size_t n = size_t(w) * size_t(h);
    std::vector<uint8_t> data(n * size_t(comp));
    s.skip(id_len);
    auto read555 = [&](uint8_t *o) {
        const int px = s.le16();
        o[0] = uint8_t((((px >> 10) & 31) * 255) / 31);
        o[1] = uint8_t((((px >> 5) & 31) * 255) / 31);
        o[2] = uint8_t(((px & 31) * 255) / 31);
    };
    if (!indexed && !rle && !rgb16) {
        for (int i = 0; i < h; ++i) {
            const int row = top_down ? i : h - 1 - i;
            s.getn(&data[size_t(row) * size_t(w) * size_t(comp)], size_t(w) * size_t(comp));
        }
    } else {
        std::vector<uint8_t> palette;
        if (indexed) {
            s.skip(pal_start);
            palette.assign(size_t(pal_len) * size_t(comp) + 4, 0);
            if (rgb16) {
                for (int i = 0; i < pal_len; ++i) read555(&palette[size_t(i) * size_t(comp)]);
            } else if (!s.getn(palette.data(), size_t(pal_len) * size_t(comp))) {
                return fail("bad palette");
            }
        }
        uint8_t raw[4] = {};
        int run = 0;
        bool repeating = false, read_next = true;
        for (size_t i = 0; i < n; ++i) {
            if (rle) {
                if (run == 0) {
                    const int cmd = s.u8();
                    run = 1 + (cmd & 127);
                    repeating = (cmd >> 7) != 0;
                    read_next = true;
                } else if (!repeating) {
                    read_next = true;
                }
            } else {
                read_next = true;
            }
            if (read_next) {
                if (indexed) {
                    int idx = bpp == 8 ? s.u8() : s.le16();
                    if (idx >= pal_len) idx = 0;
                    for (int j = 0; j < comp; ++j) raw[j] = palette[size_t(idx) * size_t(comp) + size_t(j)];
                } else if (rgb16) {
                    read555(raw);
                } else {
                    for (int j = 0; j < comp; ++j) raw[j] = uint8_t(s.u8());
                }
                read_next = false;
            }
            for (int j = 0; j < comp; ++j) data[i * size_t(comp) + size_t(j)] = raw[j];
            --run;
        }
        if (!top_down) {
            const size_t row = size_t(w) * size_t(comp);
            std::vector<uint8_t> tmp(row);
            for (int j = 0; j * 2 < h; ++j) {
                uint8_t *a = &data[row * size_t(j)], *b = &data[row * size_t(h - 1 - j)];
                if (a == b) continue;
                std::memcpy(tmp.data(), a, row); std::memcpy(a, b, row); std::memcpy(b, tmp.data(), row);
            }
        }
    }
    if (comp >= 3 && !rgb16)
        for (size_t i = 0; i < n; ++i) { uint8_t *p = &data[i * size_t(comp)]; const uint8_t t = p[0]; p[0] = p[2]; p[2] = t; }
    img.w = w; img.h = h; img.comp = comp;
    img.rgba.resize(n * 4);
    expand_to_rgba(data.data(), comp, n, img.rgba.data());
    return true;
}

}  // namespace astc_image
