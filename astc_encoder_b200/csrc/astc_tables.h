// astc_tables.h -- ASTC lookup tables generated at compile time from the
// specification's decode rules (host + device, C++17 constexpr).
//
// Replaces the literal tables of the reference:
//   bits_trits_quints_table  ASTC_IntegerSequenceEncoding.hlsl:5-28
//   integer_from_trits       ASTC_IntegerSequenceEncoding.hlsl:30-62
//   integer_from_quints      ASTC_IntegerSequenceEncoding.hlsl:64-71
//   scramble_table           ASTC_Table.hlsl:3-66
// tests/test_tables.py checks every entry against vectors extracted from the
// reference (tests/golden/ref_tables.json).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ASTC_HD __host__ __device__
#else
#define ASTC_HD
#endif

namespace astc {

enum : int {
    QUANT_2 = 0, QUANT_3, QUANT_4, QUANT_5, QUANT_6, QUANT_8, QUANT_10, QUANT_12,
    QUANT_16, QUANT_20, QUANT_24, QUANT_32, QUANT_40, QUANT_48, QUANT_64, QUANT_80,
    QUANT_96, QUANT_128, QUANT_160, QUANT_192, QUANT_256, QUANT_MAX
};
enum : int { CEM_LDR_RGB_DIRECT = 8, CEM_LDR_RGBA_DIRECT = 12 };
constexpr int kWeightMethods = 12;      // QUANT_2..QUANT_32 are legal weight ranges
constexpr int kScrambleStride = 32;     // WEIGHT_QUANTIZE_NUM (ASTC_Table.hlsl:2)

struct QuantLayout { int bits, trits, quints; };

// Levels run 2,3,4 | 5,6,8 | 10,12,16 | ... : from level 3 on a
// (quint, trit, plain) period that gains one plain bit each time round.
ASTC_HD constexpr QuantLayout quant_layout(int q)
{
    if (q <= 0) return {1, 0, 0};
    if (q == 1) return {0, 1, 0};
    if (q == 2) return {2, 0, 0};
    const int period = (q - 3) / 3, phase = (q - 3) % 3;
    if (phase == 0) return {period, 0, 1};
    if (phase == 1) return {period + 1, 1, 0};
    return {period + 3, 0, 0};
}

ASTC_HD constexpr int quant_levels(int q)
{
    const QuantLayout l = quant_layout(q);
    return (1 << l.bits) * (l.trits ? 3 : l.quints ? 5 : 1);
}

// compute_ise_bitcount (ASTC_IntegerSequenceEncoding.hlsl:76-93)
ASTC_HD constexpr uint32_t ise_bitcount(uint32_t items, int q)
{
    const QuantLayout l = quant_layout(q);
    if (l.trits) return ((8u + 5u * uint32_t(l.bits)) * items + 4u) / 5u;
    if (l.quints) return ((7u + 3u * uint32_t(l.bits)) * items + 2u) / 3u;
    return items * uint32_t(l.bits);
}

// ---- ASTC spec C.2.12: packed trit / quint block -> digits -------------
struct Trits { int t[5]; };
struct Quints { int q[3]; };

ASTC_HD constexpr Trits trits_from_integer(int T)
{
    Trits r{};
    int C = 0;
    if (((T >> 2) & 7) == 7) {
        C = (((T >> 5) & 7) << 2) | (T & 3);
        r.t[4] = 2; r.t[3] = 2;
    } else {
        C = T & 31;
        if (((T >> 5) & 3) == 3) { r.t[4] = 2; r.t[3] = (T >> 7) & 1; }
        else { r.t[4] = (T >> 7) & 1; r.t[3] = (T >> 5) & 3; }
    }
    if ((C & 3) == 3) {
        r.t[2] = 2; r.t[1] = (C >> 4) & 1;
        r.t[0] = (((C >> 3) & 1) << 1) | (((C >> 2) & 1) & ~((C >> 3) & 1));
    } else if (((C >> 2) & 3) == 3) {
        r.t[2] = 2; r.t[1] = 2; r.t[0] = C & 3;
    } else {
        r.t[2] = (C >> 4) & 1; r.t[1] = (C >> 2) & 3;
        r.t[0] = (((C >> 1) & 1) << 1) | ((C & 1) & ~((C >> 1) & 1));
    }
    return r;
}

ASTC_HD constexpr Quints quints_from_integer(int Q)
{
    Quints r{};
    int C = 0;
    if (((Q >> 1) & 3) == 3 && ((Q >> 5) & 3) == 0) {
        const int b0 = Q & 1;
        r.q[2] = (b0 << 2) | ((((Q >> 4) & 1) & ~b0) << 1) | (((Q >> 3) & 1) & ~b0);
        r.q[1] = 4; r.q[0] = 4;
        return r;
    }
    if (((Q >> 1) & 3) == 3) {
        r.q[2] = 4;
        C = (((Q >> 3) & 3) << 3) | ((~(Q >> 5) & 3) << 1) | (Q & 1);
    } else {
        r.q[2] = (Q >> 5) & 3;
        C = Q & 31;
    }
    if ((C & 7) == 5) { r.q[1] = 4; r.q[0] = (C >> 3) & 3; }
    else { r.q[1] = (C >> 3) & 3; r.q[0] = C & 7; }
    return r;
}

// Encoder-side inverse.  Where several packed values decode to one tuple the
// reference tables hold the largest; ascending overwrite reproduces that.
struct TritPack { uint8_t v[243]; };
struct QuintPack { uint8_t v[125]; };

constexpr TritPack make_trit_pack()
{
    TritPack p{};
    for (int T = 0; T < 256; ++T) {
        const Trits d = trits_from_integer(T);
        p.v[d.t[4] * 81 + d.t[3] * 27 + d.t[2] * 9 + d.t[1] * 3 + d.t[0]] = uint8_t(T);
    }
    return p;
}

constexpr QuintPack make_quint_pack()
{
    QuintPack p{};
    for (int Q = 0; Q < 128; ++Q) {
        const Quints d = quints_from_integer(Q);
        if (d.q[0] < 5 && d.q[1] < 5 && d.q[2] < 5)
            p.v[d.q[2] * 25 + d.q[1] * 5 + d.q[0]] = uint8_t(Q);
    }
    return p;
}

// ---- ASTC spec C.2.17: weight unquantisation of an ENCODED index -------
ASTC_HD constexpr int unquant_weight(int method, int v)
{
    const QuantLayout l = quant_layout(method);
    int r = 0;
    if (!l.trits && !l.quints) {
        int acc = 0, have = 0;
        while (have < 6) { acc = (acc << l.bits) | v; have += l.bits; }
        r = acc >> (have - 6);
    } else if (l.bits == 0) {
        if (l.trits) r = v == 0 ? 0 : v == 1 ? 32 : 63;
        else r = v == 0 ? 0 : v == 1 ? 16 : v == 2 ? 32 : v == 3 ? 47 : 63;
    } else {
        const int m = v & ((1 << l.bits) - 1), d = v >> l.bits;
        const int a = m & 1, b = (m >> 1) & 1, c = (m >> 2) & 1;
        const int A = a ? 0x7F : 0;
        int B = 0, C = 0;
        if (l.trits) {
            if (l.bits == 1) { B = 0; C = 50; }
            else if (l.bits == 2) { B = (b << 6) | (b << 2) | b; C = 23; }
            else { B = (c << 6) | (b << 5) | (c << 1) | b; C = 11; }
        } else {
            if (l.bits == 1) { B = 0; C = 28; }
            else { B = (b << 6) | (b << 1); C = 13; }
        }
        int T = d * C + B;
        T ^= A;
        r = (A & 0x20) | (T >> 2);
    }
    return r > 32 ? r + 1 : r;
}

// scramble[method][rank] = encoded index whose reconstruction is the
// rank-th smallest; unscramble is its inverse; unq the reconstruction.
struct WeightTables {
    uint8_t scramble[kWeightMethods][kScrambleStride];
    uint8_t unscramble[kWeightMethods][kScrambleStride];
    uint8_t unq[kWeightMethods][kScrambleStride];
};

constexpr WeightTables make_weight_tables()
{
    WeightTables w{};
    for (int m = 0; m < kWeightMethods; ++m) {
        const int n = quant_levels(m);
        for (int v = 0; v < n; ++v) w.unq[m][v] = uint8_t(unquant_weight(m, v));
        for (int v = 0; v < n; ++v) {
            int rank = 0;
            for (int u = 0; u < n; ++u)
                if (w.unq[m][u] < w.unq[m][v]) ++rank;
            w.scramble[m][rank] = uint8_t(v);
            w.unscramble[m][v] = uint8_t(rank);
        }
    }
    return w;
}

// assemble_blockmode (ASTC_Encode.hlsl:446-473) for the fixed 4x4 weight grid.
ASTC_HD constexpr uint32_t blockmode_4x4grid(int weight_quant)
{
    const uint32_t a = 2u, b = 0u;
    const uint32_t h = weight_quant < 6 ? 0u : 1u;
    const uint32_t r = uint32_t(weight_quant % 6) + 2u;
    return ((r >> 1) & 3u) | ((r & 1u) << 4) | (a << 5) | (b << 7) | (h << 9);
}

}  // namespace astc
