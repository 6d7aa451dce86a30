// jpeg_io.cpp -- JPEG ingest for astc_b200_load_image: baseline, extended-sequential and progressive
// Huffman JPEG, 8-bit, 1 / 3 / 4 components (grey, YCbCr, RGB, Adobe CMYK / YCCK), any sampling
// factors, restart intervals, 8- and 16-bit quantisation tables.
//
// The reference loads its input with the vendored stb_image v2.22 (main.cpp:24-25,
// stbi_load(..., STBI_rgb_alpha)), so "the same texels as the reference" means stb's arithmetic, not
// libjpeg's: a JPEG decoder is only specified up to IDCT accuracy and the up-sampling filter is a
// free choice.  This is an independent implementation of the arithmetic stb_image publishes
// (public domain, Sean Barrett et al.):
//   * inverse DCT: the 12-bit fixed-point Loeffler-Ligtenberg-Moschytz form ("islow"), column pass
//     keeping two extra bits, row pass rounding at 1<<17 with the +128 level shift folded in;
//   * chroma up-sampling: 3:1 "triangle" filters -- (3n + f + 2) >> 2 in one direction,
//     (3a + b + 8) >> 4 on the 3n + f sums in both -- nearest neighbour for other factors;
//   * YCbCr -> RGB in 20-bit fixed point with the constants rounded to 12 bits first;
//   * CMYK / YCCK through the Adobe APP14 transform flag with the (t + (t >> 8)) >> 8 multiply.
// tests/test_image_formats.py compares the output with stb_image itself (compiled in the build
// container from the reference checkout, oracle/_ref) and with committed fixtures made by it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "image_formats.h"

namespace astc_image {
namespace {

constexpr uint8_t kZigzag[64 + 15] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
    // a corrupt run may step past 63: land on the last coefficient instead of outside the block
    63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huffman {
    uint8_t fast[512];                 // 9-bit prefix -> symbol index, 255 = longer code
    uint16_t code[256];
    uint8_t values[256], size[257];
    uint32_t maxcode[18];
    int delta[17];
    bool valid = false;

    bool build(const int counts[16])
    {
        int k = 0;
        for (int i = 0; i < 16; ++i)
            for (int j = 0; j < counts[i]; ++j) {
                if (k >= 256) return false;
                size[k++] = uint8_t(i + 1);
            }
        size[k] = 0;
        uint32_t c = 0;
        k = 0;
        for (int len = 1; len <= 16; ++len) {
            delta[len] = k - int(c);
            if (size[k] == len) {
                while (size[k] == len) code[k++] = uint16_t(c++);
                if (c - 1 >= (1u << len)) return false;
            }
            maxcode[len] = c << (16 - len);            // first code of this length that is too large, left-aligned
            c <<= 1;
        }
        maxcode[17] = 0xFFFFFFFFu;
        std::memset(fast, 255, sizeof fast);
        for (int i = 0; i < k; ++i) {
            const int s = size[i];
            if (s <= 9) {
                const int c0 = code[i] << (9 - s), n = 1 << (9 - s);
                for (int j = 0; j < n; ++j) fast[c0 + j] = uint8_t(i);
            }
        }
        valid = true;
        return true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dc_pred = 0;
    int x = 0, y = 0, w2 = 0, h2 = 0;          // true size and size padded to whole MCUs
    std::vector<uint8_t> data;                 // w2 * h2 samples after the inverse DCT
    std::vector<int16_t> coeff;                // progressive: all coefficients, 64 per block
    int coeff_w = 0, coeff_h = 0;
};

class Decoder {
  public:
    Decoder(const uint8_t *p, size_t n) : cur_(p), end_(p + n) {}
    bool run(Image &img);

  private:
    // ---- byte stream ----
    int get8() { return cur_ < end_ ? *cur_++ : 0; }
    int get16() { const int hi = get8(); return (hi << 8) | get8(); }
    void skip(int n) { cur_ = (n < 0 || size_t(n) > size_t(end_ - cur_)) ? end_ : cur_ + n; }
    bool eof() const { return cur_ >= end_; }

    // ---- entropy-coded segment: MSB-first bit buffer, FF00 un-stuffing, stops at a marker ----
    void refill()
    {
        do {
            uint32_t b = nomore_ ? 0u : uint32_t(get8());
            if (b == 0xFF) {
                int c = get8();
                while (c == 0xFF) c = get8();          // fill bytes
                if (c != 0) {
                    marker_ = uint8_t(c);
                    nomore_ = true;
                    return;
                }
            }
            bits_ |= b << (24 - nbits_);
            nbits_ += 8;
        } while (nbits_ <= 24);
    }
    int decode(const Huffman &h)
    {
        if (nbits_ < 16) refill();
        const int f = h.fast[bits_ >> 23];
        if (f < 255) {
            const int s = h.size[f];
            if (s > nbits_) return -1;
            bits_ <<= s;
            nbits_ -= s;
            return h.values[f];
        }
        const uint32_t top = bits_ >> 16;
        int len = 10;
        while (top >= h.maxcode[len]) ++len;
        if (len == 17) { nbits_ -= 16; return -1; }
        if (len > nbits_) return -1;
        const int idx = int((bits_ >> (32 - len)) & ((1u << len) - 1u)) + h.delta[len];
        if (idx < 0 || idx > 255) return -1;
        bits_ <<= len;
        nbits_ -= len;
        return h.values[idx];
    }
    int receive(int n)                                    // n raw bits
    {
        if (n == 0) return 0;
        if (nbits_ < n) refill();
        const int v = int(bits_ >> (32 - n));
        bits_ <<= n;
        nbits_ -= n;
        return v;
    }
    int extend(int n)                                     // n bits, JPEG sign extension (F.12)
    {
        const int v = receive(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
    void reset_entropy()
    {
        nbits_ = 0;
        bits_ = 0;
        nomore_ = false;
        marker_ = 0xFF;
        for (auto &c : comp_) c.dc_pred = 0;
        todo_ = restart_interval_ ? restart_interval_ : 0x7FFFFFFF;
        eob_run_ = 0;
    }

    bool marker_segment(int m);
    bool frame_header(int m);
    bool scan_header();
    bool scan_data();
    bool block_baseline(int16_t d[64], Component &c);
    bool block_dc_progressive(int16_t d[64], Component &c);
    bool block_ac_progressive(int16_t d[64], const Huffman &h);
    bool restart_if_due();
    void finish_progressive();
    void output(Image &img);
    int next_marker()
    {
        if (marker_ != 0xFF) { const int m = marker_; marker_ = 0xFF; return m; }
        int x = get8();
        if (x != 0xFF) return 0xFF;                       // "none"
        while (x == 0xFF) x = get8();
        return x;
    }

    const uint8_t *cur_, *end_;
    uint32_t bits_ = 0;
    int nbits_ = 0;
    bool nomore_ = false;
    uint8_t marker_ = 0xFF;
    Huffman dc_[4], ac_[4];
    uint16_t dequant_[4][64] = {};
    std::vector<Component> comp_;
    int width_ = 0, height_ = 0, hmax_ = 1, vmax_ = 1, mcus_x_ = 0, mcus_y_ = 0;
    bool progressive_ = false, jfif_ = false;
    int adobe_transform_ = -1, rgb_ids_ = 0;
    int restart_interval_ = 0, todo_ = 0;
    int scan_n_ = 0, order_[4] = {}, spec_start_ = 0, spec_end_ = 0, succ_high_ = 0, succ_low_ = 0, eob_run_ = 0;
};

// ---------------------------------------------------------------------------------------------
// inverse DCT
// ---------------------------------------------------------------------------------------------
constexpr int fix12(float x) { return int(double(x * 4096.0f) + 0.5); }      // constant -> 12-bit fixed point, truncating like a C cast

struct Idct1D {
    int x0, x1, x2, x3, t0, t1, t2, t3;
};

inline Idct1D idct1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7)
{
    Idct1D r;
    // even part
    int p1 = (s2 + s6) * fix12(0.5411961f);
    const int e2 = p1 + s6 * fix12(-1.847759065f);
    const int e3 = p1 + s2 * fix12(0.765366865f);
    const int e0 = (s0 + s4) * 4096, e1 = (s0 - s4) * 4096;
    r.x0 = e0 + e3;
    r.x3 = e0 - e3;
    r.x1 = e1 + e2;
    r.x2 = e1 - e2;
    // odd part
    int t0 = s7, t1 = s5, t2 = s3, t3 = s1;
    int p3 = t0 + t2, p4 = t1 + t3;
    p1 = t0 + t3;
    int p2 = t1 + t2;
    const int p5 = (p3 + p4) * fix12(1.175875602f);
    t0 *= fix12(0.298631336f);
    t1 *= fix12(2.053119869f);
    t2 *= fix12(3.072711026f);
    t3 *= fix12(1.501321110f);
    p1 = p5 + p1 * fix12(-0.899976223f);
    p2 = p5 + p2 * fix12(-2.562915447f);
    p3 *= fix12(-1.961570560f);
    p4 *= fix12(-0.390180644f);
    r.t3 = t3 + p1 + p4;
    r.t2 = t2 + p2 + p3;
    r.t1 = t1 + p2 + p4;
    r.t0 = t0 + p1 + p3;
    return r;
}

inline uint8_t clamp8(int v) { return uint8_t(v < 0 ? 0 : v > 255 ? 255 : v); }

void idct_block(uint8_t *out, int stride, const int16_t d[64])
{
    int tmp[64];
    for (int c = 0; c < 8; ++c) {                          // columns: keep 2 extra bits
        const int16_t *s = d + c;
        int *v = tmp + c;
        if (!(s[8] | s[16] | s[24] | s[32] | s[40] | s[48] | s[56])) {
            const int dc = s[0] * 4;
            for (int r = 0; r < 8; ++r) v[8 * r] = dc;
            continue;
        }
        Idct1D k = idct1d(s[0], s[8], s[16], s[24], s[32], s[40], s[48], s[56]);
        k.x0 += 512; k.x1 += 512; k.x2 += 512; k.x3 += 512;
        v[0] = (k.x0 + k.t3) >> 10;  v[56] = (k.x0 - k.t3) >> 10;
        v[8] = (k.x1 + k.t2) >> 10;  v[48] = (k.x1 - k.t2) >> 10;
        v[16] = (k.x2 + k.t1) >> 10; v[40] = (k.x2 - k.t1) >> 10;
        v[24] = (k.x3 + k.t0) >> 10; v[32] = (k.x3 - k.t0) >> 10;
    }
    for (int r = 0; r < 8; ++r, out += stride) {           // rows: remove 1 << 17, round, level shift
        const int *v = tmp + 8 * r;
        Idct1D k = idct1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        constexpr int bias = 65536 + (128 << 17);
        k.x0 += bias; k.x1 += bias; k.x2 += bias; k.x3 += bias;
        out[0] = clamp8((k.x0 + k.t3) >> 17); out[7] = clamp8((k.x0 - k.t3) >> 17);
        out[1] = clamp8((k.x1 + k.t2) >> 17); out[6] = clamp8((k.x1 - k.t2) >> 17);
        out[2] = clamp8((k.x2 + k.t1) >> 17); out[5] = clamp8((k.x2 - k.t1) >> 17);
        out[3] = clamp8((k.x3 + k.t0) >> 17); out[4] = clamp8((k.x3 - k.t0) >> 17);
    }
}

// ---------------------------------------------------------------------------------------------
// headers
// ---------------------------------------------------------------------------------------------
bool Decoder::marker_segment(int m)
{
    if (m == 0xFF) return fail("expected marker");
    if (m == 0xDD) {                                        // DRI
        if (get16() != 4) return fail("bad DRI len");
        restart_interval_ = get16();
        return true;
    }
    if (m == 0xDB) {                                        // DQT
        int len = get16() - 2;
        while (len > 0) {
            const int q = get8(), prec = q >> 4, t = q & 15;
            if (prec > 1) return fail("bad DQT type");
            if (t > 3) return fail("bad DQT table");
            for (int i = 0; i < 64; ++i) dequant_[t][kZigzag[i]] = uint16_t(prec ? get16() : get8());
            len -= prec ? 129 : 65;
        }
        return len == 0 ? true : fail("bad DQT len");
    }
    if (m == 0xC4) {                                        // DHT
        int len = get16() - 2;
        while (len > 0) {
            const int q = get8(), tc = q >> 4, th = q & 15;
            if (tc > 1 || th > 3) return fail("bad DHT header");
            int counts[16], n = 0;
            for (int &c : counts) { c = get8(); n += c; }
            if (n > 256) return fail("bad DHT counts");
            Huffman &h = tc ? ac_[th] : dc_[th];
            if (!h.build(counts)) return fail("bad code lengths");
            for (int i = 0; i < n; ++i) h.values[i] = uint8_t(get8());
            len -= 17 + n;
        }
        return len == 0 ? true : fail("bad DHT len");
    }
    if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) {            // APPn / COM
        int len = get16();
        if (len < 2) return fail(m == 0xFE ? "bad COM len" : "bad APP len");
        len -= 2;
        if (m == 0xE0 && len >= 5) {
            static const char tag[5] = {'J', 'F', 'I', 'F', 0};
            bool ok = true;
            for (char c : tag) ok &= get8() == uint8_t(c);
            len -= 5;
            if (ok) jfif_ = true;
        } else if (m == 0xEE && len >= 12) {
            static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
            bool ok = true;
            for (char c : tag) ok &= get8() == uint8_t(c);
            len -= 6;
            if (ok) {
                get8(); get16(); get16();                   // version, flags0, flags1
                adobe_transform_ = get8();
                len -= 6;
            }
        }
        skip(len);
        return true;
    }
    return fail("unknown marker");
}

bool Decoder::frame_header(int m)
{
    progressive_ = m == 0xC2;
    const int len = get16();
    if (len < 11) return fail("bad SOF len");
    if (get8() != 8) return fail("only 8-bit");
    height_ = get16();
    width_ = get16();
    if (height_ == 0) return fail("no header height");
    if (width_ == 0) return fail("0 width");
    const int n = get8();
    if (n != 1 && n != 3 && n != 4) return fail("bad component count");
    if (len != 8 + 3 * n) return fail("bad SOF len");
    if (uint64_t(width_) * uint64_t(height_) > (1ull << 28)) return fail("too large");
    comp_.assign(size_t(n), Component());
    static const char rgb[3] = {'R', 'G', 'B'};
    rgb_ids_ = 0;
    for (int i = 0; i < n; ++i) {
        Component &c = comp_[i];
        c.id = get8();
        if (n == 3 && c.id == rgb[i]) ++rgb_ids_;
        const int q = get8();
        c.h = q >> 4;
        c.v = q & 15;
        if (c.h < 1 || c.h > 4) return fail("bad H");
        if (c.v < 1 || c.v > 4) return fail("bad V");
        c.tq = get8();
        if (c.tq > 3) return fail("bad TQ");
        hmax_ = i ? (c.h > hmax_ ? c.h : hmax_) : c.h;
        vmax_ = i ? (c.v > vmax_ ? c.v : vmax_) : c.v;
    }
    const int mcu_w = hmax_ * 8, mcu_h = vmax_ * 8;
    mcus_x_ = (width_ + mcu_w - 1) / mcu_w;
    mcus_y_ = (height_ + mcu_h - 1) / mcu_h;
    for (Component &c : comp_) {
        c.x = (width_ * c.h + hmax_ - 1) / hmax_;
        c.y = (height_ * c.v + vmax_ - 1) / vmax_;
        c.w2 = mcus_x_ * c.h * 8;
        c.h2 = mcus_y_ * c.v * 8;
        c.data.assign(size_t(c.w2) * size_t(c.h2), 0);
        if (progressive_) {
            c.coeff_w = c.w2 / 8;
            c.coeff_h = c.h2 / 8;
            c.coeff.assign(size_t(c.w2) * size_t(c.h2), 0);
        }
    }
    return true;
}

bool Decoder::scan_header()
{
    const int len = get16();
    scan_n_ = get8();
    if (scan_n_ < 1 || scan_n_ > 4 || scan_n_ > int(comp_.size())) return fail("bad SOS component count");
    if (len != 6 + 2 * scan_n_) return fail("bad SOS len");
    for (int i = 0; i < scan_n_; ++i) {
        const int id = get8(), q = get8();
        int which = 0;
        while (which < int(comp_.size()) && comp_[which].id != id) ++which;
        if (which == int(comp_.size())) return false;       // ignore-able garbage in stb too: it just stops
        comp_[which].hd = q >> 4;
        comp_[which].ha = q & 15;
        if (comp_[which].hd > 3 || comp_[which].ha > 3) return fail("bad huff table index");
        order_[i] = which;
    }
    spec_start_ = get8();
    spec_end_ = get8();
    const int a = get8();
    succ_high_ = a >> 4;
    succ_low_ = a & 15;
    if (progressive_) {
        if (spec_start_ > 63 || spec_end_ > 63 || spec_start_ > spec_end_ || succ_high_ > 13 || succ_low_ > 13) return fail("bad SOS");
    } else {
        if (spec_start_ != 0 || succ_high_ != 0 || succ_low_ != 0) return fail("bad SOS");
        spec_end_ = 63;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// blocks
// ---------------------------------------------------------------------------------------------
bool Decoder::block_baseline(int16_t d[64], Component &c)
{
    const Huffman &hd = dc_[c.hd], &ha = ac_[c.ha];
    const uint16_t *q = dequant_[c.tq];
    if (!hd.valid || !ha.valid) return fail("missing huffman table");
    const int t = decode(hd);
    if (t < 0 || t > 15) return fail("bad huffman code");
    std::memset(d, 0, 64 * sizeof(int16_t));
    const int dc = c.dc_pred + (t ? extend(t) : 0);
    c.dc_pred = dc;
    d[0] = int16_t(dc * q[0]);
    int k = 1;
    do {
        const int rs = decode(ha);
        if (rs < 0) return fail("bad huffman code");
        const int s = rs & 15, r = rs >> 4;
        if (s == 0) {
            if (rs != 0xF0) break;                          // end of block
            k += 16;
        } else {
            k += r;
            const int zig = kZigzag[k++];
            d[zig] = int16_t(extend(s) * q[zig]);
        }
    } while (k < 64);
    return true;
}

bool Decoder::block_dc_progressive(int16_t d[64], Component &c)
{
    if (spec_end_ != 0) return fail("can't merge dc and ac");
    if (succ_high_ == 0) {                                  // first pass: the value, shifted
        const Huffman &hd = dc_[c.hd];
        if (!hd.valid) return fail("missing huffman table");
        std::memset(d, 0, 64 * sizeof(int16_t));
        const int t = decode(hd);
        if (t < 0 || t > 15) return fail("bad huffman code");
        const int dc = c.dc_pred + (t ? extend(t) : 0);
        c.dc_pred = dc;
        d[0] = int16_t(dc * (1 << succ_low_));
    } else if (receive(1)) {                                // refinement: one more bit
        d[0] = int16_t(d[0] + (1 << succ_low_));
    }
    return true;
}

bool Decoder::block_ac_progressive(int16_t d[64], const Huffman &h)
{
    if (spec_start_ == 0) return fail("can't merge dc and ac");
    if (!h.valid) return fail("missing huffman table");
    if (succ_high_ == 0) {                                  // first pass over this band
        if (eob_run_) { --eob_run_; return true; }
        int k = spec_start_;
        do {
            const int rs = decode(h);
            if (rs < 0) return fail("bad huffman code");
            const int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (r < 15) {
                    eob_run_ = (1 << r) + (r ? receive(r) : 0) - 1;
                    break;
                }
                k += 16;
            } else {
                k += r;
                const int zig = kZigzag[k++];
                d[zig] = int16_t(extend(s) * (1 << succ_low_));
            }
        } while (k <= spec_end_);
        return true;
    }
    // refinement pass: one more bit for the coefficients already non-zero, new +-1 coefficients in between
    const int16_t bit = int16_t(1 << succ_low_);
    auto refine = [&](int16_t &p) {
        if (receive(1) && (p & bit) == 0) p = int16_t(p > 0 ? p + bit : p - bit);
    };
    if (eob_run_) {
        --eob_run_;
        for (int k = spec_start_; k <= spec_end_; ++k) {
            int16_t &p = d[kZigzag[k]];
            if (p != 0) refine(p);
        }
        return true;
    }
    int k = spec_start_;
    do {
        const int rs = decode(h);
        if (rs < 0) return fail("bad huffman code");
        int s = rs & 15, r = rs >> 4;
        if (s == 0) {
            if (r < 15) {
                eob_run_ = (1 << r) - 1 + (r ? receive(r) : 0);
                r = 64;                                     // run to the end of the band
            }                                               // r == 15: sixteen zeros = a run of 15 and then a zero "value"
        } else {
            if (s != 1) return fail("bad huffman code");
            s = receive(1) ? bit : -bit;
        }
        while (k <= spec_end_) {
            int16_t &p = d[kZigzag[k++]];
            if (p != 0) {
                refine(p);
            } else {
                if (r == 0) { p = int16_t(s); break; }
                --r;
            }
        }
    } while (k <= spec_end_);
    return true;
}

bool Decoder::restart_if_due()
{
    if (--todo_ > 0) return true;
    if (nbits_ < 24) refill();
    if (!(marker_ >= 0xD0 && marker_ <= 0xD7)) return false;   // no restart marker: the scan ends here
    reset_entropy();
    return true;
}

bool Decoder::scan_data()
{
    reset_entropy();
    int16_t block[64];
    if (scan_n_ == 1) {                                     // non-interleaved: the component's own blocks, row by row
        Component &c = comp_[order_[0]];
        const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
        for (int j = 0; j < bh; ++j)
            for (int i = 0; i < bw; ++i) {
                if (progressive_) {
                    int16_t *d = &c.coeff[64 * (size_t(i) + size_t(j) * size_t(c.coeff_w))];
                    if (!(spec_start_ == 0 ? block_dc_progressive(d, c) : block_ac_progressive(d, ac_[c.ha]))) return false;
                } else {
                    if (!block_baseline(block, c)) return false;
                    idct_block(&c.data[size_t(c.w2) * size_t(j) * 8 + size_t(i) * 8], c.w2, block);
                }
                if (!restart_if_due()) return true;
            }
        return true;
    }
    for (int my = 0; my < mcus_y_; ++my)                    // interleaved: MCU by MCU
        for (int mx = 0; mx < mcus_x_; ++mx) {
            for (int k = 0; k < scan_n_; ++k) {
                Component &c = comp_[order_[k]];
                for (int y = 0; y < c.v; ++y)
                    for (int x = 0; x < c.h; ++x) {
                        const int bx = mx * c.h + x, by = my * c.v + y;
                        if (progressive_) {
                            int16_t *d = &c.coeff[64 * (size_t(bx) + size_t(by) * size_t(c.coeff_w))];
                            if (!block_dc_progressive(d, c)) return false;      // interleaved progressive scans are DC scans
                        } else {
                            if (!block_baseline(block, c)) return false;
                            idct_block(&c.data[size_t(c.w2) * size_t(by) * 8 + size_t(bx) * 8], c.w2, block);
                        }
                    }
            }
            if (!restart_if_due()) return true;
        }
    return true;
}

void Decoder::finish_progressive()
{
    for (Component &c : comp_) {
        const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
        const uint16_t *q = dequant_[c.tq];
        for (int j = 0; j < bh; ++j)
            for (int i = 0; i < bw; ++i) {
                int16_t *d = &c.coeff[64 * (size_t(i) + size_t(j) * size_t(c.coeff_w))];
                for (int k = 0; k < 64; ++k) d[k] = int16_t(d[k] * q[k]);
                idct_block(&c.data[size_t(c.w2) * size_t(j) * 8 + size_t(i) * 8], c.w2, d);
            }
    }
}

// ---------------------------------------------------------------------------------------------
// up-sampling + colour conversion -> RGBA8
// ---------------------------------------------------------------------------------------------
const uint8_t *upsample(uint8_t *out, const uint8_t *near_row, const uint8_t *far_row, int w, int hs, int vs)
{
    if (hs == 1 && vs == 1) return near_row;
    if (hs == 1 && vs == 2) {
        for (int i = 0; i < w; ++i) out[i] = uint8_t((3 * near_row[i] + far_row[i] + 2) >> 2);
        return out;
    }
    if (hs == 2 && vs == 1) {
        const uint8_t *in = near_row;
        if (w == 1) { out[0] = out[1] = in[0]; return out; }
        out[0] = in[0];
        out[1] = uint8_t((in[0] * 3 + in[1] + 2) >> 2);
        int i = 1;
        for (; i < w - 1; ++i) {
            const int n = 3 * in[i] + 2;
            out[2 * i] = uint8_t((n + in[i - 1]) >> 2);
            out[2 * i + 1] = uint8_t((n + in[i + 1]) >> 2);
        }
        out[2 * i] = uint8_t((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
        out[2 * i + 1] = in[w - 1];
        return out;
    }
    if (hs == 2 && vs == 2) {
        if (w == 1) { out[0] = out[1] = uint8_t((3 * near_row[0] + far_row[0] + 2) >> 2); return out; }
        int t1 = 3 * near_row[0] + far_row[0];
        out[0] = uint8_t((t1 + 2) >> 2);
        for (int i = 1; i < w; ++i) {
            const int t0 = t1;
            t1 = 3 * near_row[i] + far_row[i];
            out[2 * i - 1] = uint8_t((3 * t0 + t1 + 8) >> 4);
            out[2 * i] = uint8_t((3 * t1 + t0 + 8) >> 4);
        }
        out[2 * w - 1] = uint8_t((t1 + 2) >> 2);
        return out;
    }
    for (int i = 0; i < w; ++i)                             // anything else: nearest neighbour horizontally
        for (int j = 0; j < hs; ++j) out[i * hs + j] = near_row[i];
    return out;
}

constexpr int fix20(float x) { return int(x * 4096.0f + 0.5f) << 8; }

void ycc_to_rgba(uint8_t *out, const uint8_t *y, const uint8_t *cb, const uint8_t *cr, int n)
{
    for (int i = 0; i < n; ++i, out += 4) {
        const int yf = (y[i] << 20) + (1 << 19);
        const int r_ = cr[i] - 128, b_ = cb[i] - 128;
        const int r = (yf + r_ * fix20(1.40200f)) >> 20;
        const int g = (yf + r_ * -fix20(0.71414f) + int(uint32_t(b_ * -fix20(0.34414f)) & 0xffff0000u)) >> 20;
        const int b = (yf + b_ * fix20(1.77200f)) >> 20;
        out[0] = clamp8(r); out[1] = clamp8(g); out[2] = clamp8(b); out[3] = 255;
    }
}

inline uint8_t mul8(uint8_t x, uint8_t y)
{
    const unsigned t = unsigned(x) * y + 128;
    return uint8_t((t + (t >> 8)) >> 8);
}

void Decoder::output(Image &img)
{
    const int n = int(comp_.size());
    img.w = width_; img.h = height_; img.comp = n >= 3 ? 3 : 1;
    img.rgba.assign(size_t(width_) * size_t(height_) * 4, 0);
    const bool is_rgb = n == 3 && (rgb_ids_ == 3 || (adobe_transform_ == 0 && !jfif_));
    struct Row {
        int hs, vs, w_lores, ystep, ypos;
        const uint8_t *line0, *line1;
        std::vector<uint8_t> buf;
    } rows[4];
    for (int k = 0; k < n; ++k) {
        Row &r = rows[k];
        r.hs = hmax_ / comp_[k].h;
        r.vs = vmax_ / comp_[k].v;
        r.ystep = r.vs >> 1;
        r.w_lores = (width_ + r.hs - 1) / r.hs;
        r.ypos = 0;
        r.line0 = r.line1 = comp_[k].data.data();
        r.buf.assign(size_t(width_) + 8 + size_t(r.hs) * 2, 0);
    }
    const uint8_t *c[4] = {};
    for (int j = 0; j < height_; ++j) {
        uint8_t *out = &img.rgba[size_t(width_) * 4 * size_t(j)];
        for (int k = 0; k < n; ++k) {
            Row &r = rows[k];
            const bool bottom = r.ystep >= (r.vs >> 1);
            c[k] = upsample(r.buf.data(), bottom ? r.line1 : r.line0, bottom ? r.line0 : r.line1, r.w_lores, r.hs, r.vs);
            if (++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < comp_[k].y) r.line1 += comp_[k].w2;
            }
        }
        if (n == 3) {
            if (is_rgb) {
                for (int i = 0; i < width_; ++i, out += 4) { out[0] = c[0][i]; out[1] = c[1][i]; out[2] = c[2][i]; out[3] = 255; }
            } else {
                ycc_to_rgba(out, c[0], c[1], c[2], width_);
            }
        } else if (n == 4) {
            if (adobe_transform_ == 0) {                    // CMYK
                for (int i = 0; i < width_; ++i, out += 4) {
                    const uint8_t m = c[3][i];
                    out[0] = mul8(c[0][i], m); out[1] = mul8(c[1][i], m); out[2] = mul8(c[2][i], m); out[3] = 255;
                }
            } else {
                ycc_to_rgba(out, c[0], c[1], c[2], width_);
                if (adobe_transform_ == 2) {                // YCCK
                    for (int i = 0; i < width_; ++i, out += 4) {
                        const uint8_t m = c[3][i];
                        out[0] = mul8(uint8_t(255 - out[0]), m); out[1] = mul8(uint8_t(255 - out[1]), m); out[2] = mul8(uint8_t(255 - out[2]), m);
                    }
                }
            }
        } else {
            for (int i = 0; i < width_; ++i, out += 4) { out[0] = out[1] = out[2] = c[0][i]; out[3] = 255; }
        }
    }
}

bool Decoder::run(Image &img)
{
    if (next_marker() != 0xD8) return fail("no SOI");
    int m = next_marker();
    while (!(m == 0xC0 || m == 0xC1 || m == 0xC2)) {
        if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) return fail("unsupported JPEG process (lossless / hierarchical / arithmetic)");
        if (!marker_segment(m)) return false;
        m = next_marker();
        while (m == 0xFF) {
            if (eof()) return fail("no SOF");
            m = next_marker();
        }
    }
    if (!frame_header(m)) return false;
    m = next_marker();
    while (m != 0xD9) {                                     // until EOI
        if (m == 0xDA) {
            if (!scan_header()) return false;
            if (!scan_data()) return false;
            if (marker_ == 0xFF) {
                // bytes after the entropy-coded data that are not a marker (seen from some cameras): skip to the next one
                while (!eof()) {
                    const int x = get8();
                    if (x == 0xFF) { marker_ = uint8_t(get8()); break; }
                }
                // if nothing was found marker_ stays "none": the next marker_segment() reports it
            }
        } else if (m == 0xDC) {                             // DNL
            const int len = get16();
            const int lines = get16();
            if (len != 4) return fail("bad DNL len");
            if (lines != height_) return fail("bad DNL height");
        } else if (m == 0xFF && eof()) {
            break;                                          // truncated after the last scan: use what was decoded
        } else {
            if (!marker_segment(m)) return false;
        }
        m = next_marker();
    }
    if (progressive_) finish_progressive();
    output(img);
    return true;
}

}  // namespace

bool decode_jpeg(const std::vector<uint8_t> &file, Image &img)
{
    Decoder d(file.data(), file.size());
    return d.run(img);
}

}  // namespace astc_image
