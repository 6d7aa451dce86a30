/*
 * astc_oracle.c -- scalar CPU restatement of the reference's ASTC block
 * encoder (see astc_oracle.h for scope, citations and the "test
 * infrastructure only" rule).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp ...
 * -ffp-contract=off is REQUIRED: every fused multiply-add below is written
 * explicitly with fmaf(); the compiler must not add or remove any.
 */
#include "astc_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__FAST_MATH__)
#error "the oracle must not be compiled with -ffast-math"
#endif

#define SMALL_VALUE 1e-5f          /* ASTC_Encode.hlsl:34 */
#define QUANT_6   4                /* ASTC_Encode.hlsl:51 */
#define QUANT_12  7                /* ASTC_Encode.hlsl:54 */
#define QUANT_256 20               /* ASTC_Encode.hlsl:67 */
#define CEM_LDR_RGB_DIRECT  8      /* ASTC_Encode.hlsl:39 */
#define CEM_LDR_RGBA_DIRECT 12     /* ASTC_Encode.hlsl:40 */

/* ------------------------------------------------------------------ */
/* Tables, derived from the ASTC specification rather than stored.     */
/* tests/test_oracle_tables.py checks them against vectors extracted   */
/* from the reference's ASTC_Table.hlsl / IntegerSequenceEncoding.hlsl */
/* ------------------------------------------------------------------ */

/* bits / trits / quints per quant level
 * (ASTC_IntegerSequenceEncoding.hlsl:5-28).  The ranges come in triples
 * {2^n, 3*2^(n-1)... } -- written out via the generating rule. */
void astc_oracle_quant_layout(int quant, int *bits, int *trits, int *quints)
{
    /* levels 0..20: 2,3,4,5,6,8,10,12,16,20,24,32,40,48,64,80,96,128,160,192,256 */
    int b = 0, t = 0, q = 0;
    if (quant <= 0) { b = 1; }
    else if (quant == 1) { t = 1; }
    else if (quant == 2) { b = 2; }
    else {
        /* from level 3 on the pattern (quint, trit, pure) repeats with
         * one more plain bit per period: 5,6,8 | 10,12,16 | 20,24,32 ... */
        int period = (quant - 3) / 3, phase = (quant - 3) % 3;
        if (phase == 0) { q = 1; b = period; }
        else if (phase == 1) { t = 1; b = period + 1; }
        else { b = period + 3; }
    }
    *bits = b; *trits = t; *quints = q;
}

/* ASTC_IntegerSequenceEncoding.hlsl:76-93 */
uint32_t astc_oracle_ise_bitcount(uint32_t items, int quant)
{
    int bits, trits, quints;
    astc_oracle_quant_layout(quant, &bits, &trits, &quints);
    if (trits)  return ((8u + 5u * (uint32_t)bits) * items + 4u) / 5u;
    if (quints) return ((7u + 3u * (uint32_t)bits) * items + 2u) / 3u;
    return items * (uint32_t)bits;
}

/* ASTC spec C.2.12 trit-block decode: T (8 bits) -> five trits. */
static void spec_trits_from_integer(int T, int t[5])
{
    int C;
    if (((T >> 2) & 7) == 7) {
        C = (((T >> 5) & 7) << 2) | (T & 3);
        t[4] = 2; t[3] = 2;
    } else {
        C = T & 31;
        if (((T >> 5) & 3) == 3) { t[4] = 2; t[3] = (T >> 7) & 1; }
        else { t[4] = (T >> 7) & 1; t[3] = (T >> 5) & 3; }
    }
    if ((C & 3) == 3) {
        t[2] = 2; t[1] = (C >> 4) & 1;
        t[0] = (((C >> 3) & 1) << 1) | (((C >> 2) & 1) & ~((C >> 3) & 1));
    } else if (((C >> 2) & 3) == 3) {
        t[2] = 2; t[1] = 2; t[0] = C & 3;
    } else {
        t[2] = (C >> 4) & 1; t[1] = (C >> 2) & 3;
        t[0] = (((C >> 1) & 1) << 1) | ((C & 1) & ~((C >> 1) & 1));
    }
}

/* ASTC spec C.2.12 quint-block decode: Q (7 bits) -> three quints. */
static void spec_quints_from_integer(int Q, int q[3])
{
    int C;
    if (((Q >> 1) & 3) == 3 && ((Q >> 5) & 3) == 0) {
        int b0 = Q & 1;
        q[2] = (b0 << 2) | ((((Q >> 4) & 1) & ~b0) << 1) | (((Q >> 3) & 1) & ~b0);
        q[1] = 4; q[0] = 4;
        return;
    }
    if (((Q >> 1) & 3) == 3) {
        q[2] = 4;
        C = (((Q >> 3) & 3) << 3) | ((~(Q >> 5) & 3) << 1) | (Q & 1);
    } else {
        q[2] = (Q >> 5) & 3;
        C = Q & 31;
    }
    if ((C & 7) == 5) { q[1] = 4; q[0] = (C >> 3) & 3; }
    else { q[1] = (C >> 3) & 3; q[0] = C & 7; }
}

/* Inverse tables.  Several packed integers decode to the same tuple; the
 * reference's tables (integer_from_trits, :30-62; integer_from_quints,
 * :64-71) hold the LARGEST such integer, which ascending overwrite gives. */
static uint8_t g_trit_pack[243];
static uint8_t g_quint_pack[125];
static uint8_t g_scramble[12][32];
static uint8_t g_unscramble[12][32];
static uint8_t g_weight_unq[12][32];   /* encoded index -> 0..64 */
static int g_tables_ready;

/* ASTC spec C.2.17 weight unquantisation of one ENCODED weight index. */
static int spec_unquant_weight(int method, int v)
{
    int bits, trits, quints, r;
    astc_oracle_quant_layout(method, &bits, &trits, &quints);
    if (!trits && !quints) {
        /* bit replication to 6 bits */
        int acc = 0, have = 0;
        while (have < 6) { acc = (acc << bits) | v; have += bits; }
        r = acc >> (have - 6);
    } else if (bits == 0) {
        static const int t3[3] = {0, 32, 63};
        static const int q5[5] = {0, 16, 32, 47, 63};
        r = trits ? t3[v] : q5[v];
    } else {
        int m = v & ((1 << bits) - 1), d = v >> bits;
        int a = m & 1, b = (m >> 1) & 1, c = (m >> 2) & 1;
        int A = a ? 0x7F : 0, B = 0, C = 0, T;
        if (trits) {
            if (bits == 1) { B = 0; C = 50; }
            else if (bits == 2) { B = (b << 6) | (b << 2) | b; C = 23; }
            else { B = (c << 6) | (b << 5) | (c << 1) | b; C = 11; }
        } else {
            if (bits == 1) { B = 0; C = 28; }
            else { B = (b << 6) | (b << 1); C = 13; }
        }
        T = d * C + B;
        T ^= A;
        r = (A & 0x20) | (T >> 2);
    }
    if (r > 32) r += 1;
    return r;
}

static void build_tables(void)
{
    int T, Q, m;
    for (T = 0; T < 256; ++T) {
        int t[5];
        spec_trits_from_integer(T, t);
        g_trit_pack[t[4] * 81 + t[3] * 27 + t[2] * 9 + t[1] * 3 + t[0]] = (uint8_t)T;
    }
    for (Q = 0; Q < 128; ++Q) {
        int q[3];
        spec_quints_from_integer(Q, q);
        if (q[0] < 5 && q[1] < 5 && q[2] < 5)
            g_quint_pack[q[2] * 25 + q[1] * 5 + q[0]] = (uint8_t)Q;
    }
    /* scramble: natural (sorted by reconstructed value) rank -> encoded index
     * (ASTC_Table.hlsl:3-66 is that permutation for methods 0..11). */
    memset(g_scramble, 0, sizeof g_scramble);
    memset(g_unscramble, 0, sizeof g_unscramble);
    for (m = 0; m < 12; ++m) {
        int bits, trits, quints, n, v, rank;
        astc_oracle_quant_layout(m, &bits, &trits, &quints);
        n = (1 << bits) * (trits ? 3 : quints ? 5 : 1);
        for (v = 0; v < n; ++v) g_weight_unq[m][v] = (uint8_t)spec_unquant_weight(m, v);
        for (v = 0; v < n; ++v) {
            int u;
            rank = 0;
            for (u = 0; u < n; ++u)
                if (g_weight_unq[m][u] < g_weight_unq[m][v]) ++rank;
            g_scramble[m][rank] = (uint8_t)v;
            g_unscramble[m][v] = (uint8_t)rank;
        }
    }
    g_tables_ready = 1;
}

static void ensure_tables(void)
{
    if (!g_tables_ready) {
#ifdef _OPENMP
#pragma omp critical(astc_oracle_tables)
#endif
        { if (!g_tables_ready) build_tables(); }
    }
}

uint8_t astc_oracle_integer_from_trits(int t0, int t1, int t2, int t3, int t4)
{
    ensure_tables();
    return g_trit_pack[t4 * 81 + t3 * 27 + t2 * 9 + t1 * 3 + t0];
}

uint8_t astc_oracle_integer_from_quints(int q0, int q1, int q2)
{
    ensure_tables();
    return g_quint_pack[q2 * 25 + q1 * 5 + q0];
}

uint8_t astc_oracle_scramble(int method, int q)
{
    ensure_tables();
    return g_scramble[method][q & 31];
}

/* ------------------------------------------------------------------ */
/* Bit stream + BISE (ASTC_IntegerSequenceEncoding.hlsl:96-276)        */
/* ------------------------------------------------------------------ */

/* orbits8_ptr (:98-119): OR the low `count` bits of value at bit `*pos`. */
static void put_bits(uint8_t stream[16], uint32_t *pos, uint32_t value, uint32_t count)
{
    uint32_t i;
    for (i = 0; i < count; ++i) {
        uint32_t p = *pos + i;
        if (p < 128u && ((value >> i) & 1u)) stream[p >> 3] |= (uint8_t)(1u << (p & 7u));
    }
    *pos += count;
}

/* encode_trits (:142-176): m0 T[1:0] m1 T[3:2] m2 T[4] m3 T[6:5] m4 T[7]. */
static void put_trit_group(uint8_t stream[16], uint32_t *pos, int bits, const uint8_t v[5])
{
    static const int tshift[5] = {0, 2, 4, 5, 7};
    static const int tcount[5] = {2, 2, 1, 2, 1};
    uint32_t mask = (1u << bits) - 1u, T;
    int i;
    T = astc_oracle_integer_from_trits(v[0] >> bits, v[1] >> bits, v[2] >> bits,
                                       v[3] >> bits, v[4] >> bits);
    for (i = 0; i < 5; ++i) {
        put_bits(stream, pos, v[i] & mask, (uint32_t)bits);
        put_bits(stream, pos, (T >> tshift[i]) & ((1u << tcount[i]) - 1u), (uint32_t)tcount[i]);
    }
}

/* encode_quints (:181-205): m0 Q[2:0] m1 Q[4:3] m2 Q[6:5]. */
static void put_quint_group(uint8_t stream[16], uint32_t *pos, int bits, const uint8_t v[3])
{
    static const int qshift[3] = {0, 3, 5};
    static const int qcount[3] = {3, 2, 2};
    uint32_t mask = (1u << bits) - 1u, Q;
    int i;
    Q = astc_oracle_integer_from_quints(v[0] >> bits, v[1] >> bits, v[2] >> bits);
    for (i = 0; i < 3; ++i) {
        put_bits(stream, pos, v[i] & mask, (uint32_t)bits);
        put_bits(stream, pos, (Q >> qshift[i]) & ((1u << qcount[i]) - 1u), (uint32_t)qcount[i]);
    }
}

/* bise_endpoints / bise_weights (:207-276) for any value count. */
uint32_t astc_oracle_bise_encode(const uint8_t *values, int count, int quant, uint8_t stream[16])
{
    int bits, trits, quints, i, j;
    uint32_t pos = 0;
    astc_oracle_quant_layout(quant, &bits, &trits, &quints);
    if (trits) {
        for (i = 0; i < count; i += 5) {
            uint8_t g[5] = {0, 0, 0, 0, 0};
            for (j = 0; j < 5 && i + j < count; ++j) g[j] = values[i + j];
            put_trit_group(stream, &pos, bits, g);
        }
    } else if (quints) {
        for (i = 0; i < count; i += 3) {
            uint8_t g[3] = {0, 0, 0};
            for (j = 0; j < 3 && i + j < count; ++j) g[j] = values[i + j];
            put_quint_group(stream, &pos, bits, g);
        }
    } else {
        for (i = 0; i < count; ++i) put_bits(stream, &pos, values[i], (uint32_t)bits);
    }
    return pos;
}

/* assemble_blockmode (ASTC_Encode.hlsl:446-473): 4x4 weight grid, single
 * plane; R = method%6+2 split over bits {4,1,0}, H = method>=6 at bit 9. */
uint32_t astc_oracle_blockmode(int weight_quant)
{
    uint32_t a = (4u - 2u) & 3u, b = (4u - 4u) & 3u;
    uint32_t h = weight_quant < 6 ? 0u : 1u;
    uint32_t r = (uint32_t)(weight_quant % 6) + 2u;
    return ((r >> 1) & 3u) | ((r & 1u) << 4) | (a << 5) | (b << 7) | (h << 9);
}

static uint32_t load_le32(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static void store_le32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
static uint32_t bitrev32(uint32_t v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
}

/* assemble_block (ASTC_Encode.hlsl:400-444).  The per-byte reversal written
 * there, applied to the four bytes of a word in swapped positions, is a
 * 32-bit bit reversal.  Note :438 ASSIGNS word y, discarding the reversed
 * third weight word; kept for fidelity (that word is 0 in shipped modes). */
static void assemble(uint32_t blockmode, uint32_t cem, const uint8_t ep_ise[16],
                     const uint8_t wt_ise[16], uint8_t out[16])
{
    uint32_t ex = load_le32(ep_ise), ey = load_le32(ep_ise + 4);
    uint32_t x, y, z, w;
    w = bitrev32(load_le32(wt_ise));
    z = bitrev32(load_le32(wt_ise + 4));
    x = blockmode | ((cem & 0xFu) << 13) | ((ex & 0x7FFFu) << 17);
    y = ((ex >> 15) & 0x1FFFFu) | ((ey & 0x7FFFu) << 17);
    z |= (ey >> 15) & 0x1FFFFu;
    store_le32(out, x); store_le32(out + 4, y); store_le32(out + 8, z); store_le32(out + 12, w);
}

/* ------------------------------------------------------------------ */
/* Float stages                                                        */
/* ------------------------------------------------------------------ */

/* HLSL dot() on float4, as a contracted multiply-add chain x,y,z,w. */
static inline float dot4(const float a[4], const float b[4])
{
    return fmaf(a[3], b[3], fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])));
}

static inline float clamp255(float v)
{
    return fminf(fmaxf(v, 0.0f), 255.0f);
}

/* ---------------------------------------------------------------------------
 * rcp / rsq as the reference's hardware evaluates them.
 *
 * `1.0f / x` (ASTC_Encode.hlsl:366) and normalize() (:103,332) compile to the DXBC rcp / rsq
 * instructions, which D3D11 only specifies to ~1 ulp; the GPU that produced textures/leaf.astc
 * evaluates them with NVIDIA's MUFU.RCP / MUFU.RSQ units.  Evidence: with correctly rounded
 * 1/sqrt and 1/x this restatement reproduces 99.63 % of the golden's blocks, with the MUFU
 * results 99.94 % (tests/test_oracle_golden.py).  The units are emulated exactly from delta
 * tables captured on a B200 (tools/gen_mufu_tables.py -> oracle/tables/): for a positive normal x
 *     rcp(x) = bits(float(1.0 / (double)x))        + rcp_delta[mantissa(x)]
 *     rsq(x) = bits(float(1.0 / sqrt((double)x)))  + rsq_delta[exponent parity(x) : mantissa(x)]
 * (both units scale exactly with the exponent: 0 mismatches on 2^26 random inputs over the whole
 * normal range, checked by the generator and again by tests/test_gpu_edges.py on the device).
 * Every argument the encoder feeds them is positive and normal (>= 1e-20). */
static const int8_t *g_rcp_delta = NULL, *g_rsq_delta = NULL;
static unsigned g_variant = 0;           /* sensitivity switches, tools/pin_sensitivity.py only */

void astc_oracle_set_variant(unsigned flags) { g_variant = flags; }
static int g_rcp_bias = 0, g_rsq_bias = 0;  /* ulps added to every rcp / rsq result, tools/golden_residual.py only */
void astc_oracle_set_mufu_bias(int rcp_ulps, int rsq_ulps) { g_rcp_bias = rcp_ulps; g_rsq_bias = rsq_ulps; }

void astc_oracle_set_mufu_tables(const int8_t *rcp_delta, const int8_t *rsq_delta)
{
    g_rcp_delta = rcp_delta;
    g_rsq_delta = rsq_delta;
}

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

float astc_oracle_mufu_rcp(float x)
{
    const uint32_t b = f2u(x);
    const float base = (float)(1.0 / (double)x);
    if (g_variant & ASTC_ORACLE_VAR_EXACT_RCP_RSQ) return base;
    if (!g_rcp_delta) return u2f(0x7FC00000u);                 /* tables not loaded: poison, never a silent fallback */
    return u2f(f2u(base) + (uint32_t)(int32_t)g_rcp_delta[b & 0x7FFFFFu] + (uint32_t)g_rcp_bias);
}

float astc_oracle_mufu_rsq(float x)
{
    const uint32_t b = f2u(x), parity = ((b >> 23) + 1u) & 1u;  /* biased exponent 127 ([1,2)) -> 0 */
    const float base = (float)(1.0 / sqrt((double)x));
    if (g_variant & ASTC_ORACLE_VAR_EXACT_RCP_RSQ) return base;
    if (!g_rsq_delta) return u2f(0x7FC00000u);
    return u2f(f2u(base) + (uint32_t)(int32_t)g_rsq_delta[(parity << 23) | (b & 0x7FFFFFu)] + (uint32_t)g_rsq_bias);
}

/* eigen_vector (ASTC_Encode.hlsl:93-106): power iteration, two mat-vecs per
 * round, early return of the un-normalised vector when it collapses. */
static void power_iteration(const float m[16], float v[4])
{
    int it, r;
    v[0] = 0.26726f; v[1] = 0.80178f; v[2] = 0.53452f; v[3] = 0.0f;
    for (it = 0; it < 8; ++it) {
        float u[4], w[4], inv;
        for (r = 0; r < 4; ++r) u[r] = dot4(&m[4 * r], v);
        if (sqrtf(dot4(u, u)) < SMALL_VALUE) {
            memcpy(v, u, sizeof u);
            return;
        }
        for (r = 0; r < 4; ++r) w[r] = dot4(&m[4 * r], u);
        inv = astc_oracle_mufu_rsq(dot4(w, w));                 /* normalize(): dp4, rsq, mul */
        for (r = 0; r < 4; ++r) v[r] = w[r] * inv;
    }
}

/* texel*255 - base: one FMA in the canonical arithmetic (the golden prefers it: DESIGN.md 2). */
static inline float deviation(float raw, float base)
{
    if (g_variant & ASTC_ORACLE_VAR_UNFUSED_DEV) return raw * 255.0f - base;
    return fmaf(raw, 255.0f, -base);
}

/* pt_mean of principal_component_analysis / max_accumulation_pixel_direction (:142-147, :172-178). */
static void block_mean(const float (*raw)[4], int bs, float mean[4])
{
    float sum[4] = {0, 0, 0, 0};
    const float inv_n = 1.0f / (float)bs;
    int k, i;
    for (k = 0; k < bs; ++k)
        for (i = 0; i < 4; ++i) sum[i] = sum[i] + raw[k][i] * 255.0f;
    for (i = 0; i < 4; ++i)
        mean[i] = (g_variant & ASTC_ORACLE_VAR_TRUE_DIVISION) ? sum[i] / (float)bs : sum[i] * inv_n;
}

/* principal_component_analysis up to the axis (ASTC_Encode.hlsl:149-164). */
static void pca_axis(const float (*raw)[4], int bs, const float mean[4], float cov[16], float axis[4])
{
    const float inv_n1 = 1.0f / (float)(bs - 1);
    int k, i, j;
    for (i = 0; i < 16; ++i) cov[i] = 0.0f;
    for (k = 0; k < bs; ++k) {
        float d[4];
        for (i = 0; i < 4; ++i) d[i] = deviation(raw[k][i], mean[i]);
        for (i = 0; i < 4; ++i)
            for (j = 0; j < 4; ++j) cov[4 * i + j] = fmaf(d[i], d[j], cov[4 * i + j]);
    }
    for (i = 0; i < 16; ++i)
        cov[i] = (g_variant & ASTC_ORACLE_VAR_TRUE_DIVISION) ? cov[i] / (float)(bs - 1) : cov[i] * inv_n1;
    power_iteration(cov, axis);
}

/* max_accumulation_pixel_direction up to the axis (ASTC_Encode.hlsl:180-223): for each channel c the
 * sum of the deviations of the texels that lie above the mean in c; the longest of the four sums
 * (three without alpha; strict >, so ties keep the earlier channel) is the direction, normalised
 * unless shorter than SMALL_VALUE.  `sum += cond ? dt : 0` adds an exact zero when the condition
 * fails, so it is restated as a conditional add. */
static void accumulation_axis(const float (*raw)[4], int bs, int has_alpha, const float mean[4], float axis[4])
{
    float sum[4][4] = {{0}}, best;
    int k, i, c, pick = 0;
    for (k = 0; k < bs; ++k) {
        float d[4];
        for (i = 0; i < 4; ++i) d[i] = deviation(raw[k][i], mean[i]);
        for (c = 0; c < 4; ++c)
            if (d[c] > 0.0f)
                for (i = 0; i < 4; ++i) sum[c][i] = sum[c][i] + d[i];
    }
    best = dot4(sum[0], sum[0]);
    for (c = 1; c < (has_alpha ? 4 : 3); ++c) {
        const float dc = dot4(sum[c], sum[c]);
        if (dc > best) { best = dc; pick = c; }
    }
    memcpy(axis, sum[pick], sizeof sum[pick]);
    if (!(sqrtf(best) < SMALL_VALUE)) {                           /* safe normalize (:219-221) */
        const float inv = astc_oracle_mufu_rsq(best);
        for (i = 0; i < 4; ++i) axis[i] = axis[i] * inv;
    }
}

/* principal_component_analysis / max_accumulation_pixel_direction + find_min_max
 * (ASTC_Encode.hlsl:108-227). */
static void pca_endpoints(const float (*raw)[4], int bs, int has_alpha, int axis_method,
                          float e0[4], float e1[4], astc_oracle_trace *tr)
{
    float mean[4], cov[16], axis[4];
    float lo = 1e31f, hi = -1e31f, s0, s1;
    int k, i;

    block_mean(raw, bs, mean);
    memset(cov, 0, sizeof cov);
    if (axis_method == 1) accumulation_axis(raw, bs, has_alpha, mean, axis);
    else pca_axis(raw, bs, mean, cov, axis);

    for (k = 0; k < bs; ++k) {
        float d[4], t;
        for (i = 0; i < 4; ++i) d[i] = deviation(raw[k][i], mean[i]);
        t = dot4(d, axis);
        lo = fminf(lo, t);
        hi = fmaxf(hi, t);
    }
    for (i = 0; i < 4; ++i) {
        e0[i] = clamp255(fmaf(axis[i], lo, mean[i]));
        e1[i] = clamp255(fmaf(axis[i], hi, mean[i]));
    }
    /* darkest endpoint first, compared on the ROUNDED rgb sums (:125-130) */
    s0 = rintf(e0[0]) + rintf(e0[1]) + rintf(e0[2]);
    s1 = rintf(e1[0]) + rintf(e1[1]) + rintf(e1[2]);
    if (s0 > s1) {
        for (i = 0; i < 4; ++i) { float t = e0[i]; e0[i] = e1[i]; e1[i] = t; }
    }
    if (!has_alpha) { e0[3] = 255.0f; e1[3] = 255.0f; }

    if (tr) {
        memcpy(tr->mean, mean, sizeof mean);
        memcpy(tr->cov, cov, sizeof cov);
        memcpy(tr->axis, axis, sizeof axis);
    }
}

/* 6x6 block -> 4x4 weight grid taps (ASTC_Encode.hlsl:268-304).  Grid cell g
 * covers texel columns {x0,x0+1}, x0 = 3*(gx/2) + (gx&1), the heavier tap on
 * the outer side; rows alike.  Tap weights are the literals 0.444/0.222/0.111. */
static void grid_taps(int g, int idx[4], float wt[4])
{
    int gx = g & 3, gy = g >> 2, t;
    int x0 = 3 * (gx >> 1) + (gx & 1), y0 = 3 * (gy >> 1) + (gy & 1);
    for (t = 0; t < 4; ++t) {
        int dx = t & 1, dy = t >> 1;
        int heavy = ((dx == (gx & 1)) ? 1 : 0) + ((dy == (gy & 1)) ? 1 : 0);
        idx[t] = (y0 + dy) * 6 + (x0 + dx);
        wt[t] = heavy == 2 ? 0.444f : heavy == 1 ? 0.222f : 0.111f;
    }
}

/* calculate_normal_weights (ASTC_Encode.hlsl:316-372), on the UNROUNDED
 * endpoints; output always spans [0,1] unless the endpoints coincide. */
static void project_weights(const float (*raw)[4], int dim, const float e0[4],
                            const float e1[4], float projw[16])
{
    float vk[4], k[4], len, inv, lo = 1e31f, hi = -1e31f, span;
    int i, c;
    for (c = 0; c < 4; ++c) vk[c] = e1[c] - e0[c];
    len = sqrtf(dot4(vk, vk));
    if (len < SMALL_VALUE) {
        for (i = 0; i < 16; ++i) projw[i] = 0.0f;
        return;
    }
    inv = astc_oracle_mufu_rsq(dot4(vk, vk));
    for (c = 0; c < 4; ++c) k[c] = vk[c] * inv;

    for (i = 0; i < 16; ++i) {
        float d[4], w;
        if (dim == 4) {
            for (c = 0; c < 4; ++c) d[c] = deviation(raw[i][c], e0[c]);
        } else {
            int idx[4];
            float wt[4];
            grid_taps(i, idx, wt);
            for (c = 0; c < 4; ++c) {
                float s = (raw[idx[0]][c] * 255.0f) * wt[0];
                if (g_variant & ASTC_ORACLE_VAR_UNFUSED_SAMPLE) {
                    s = s + (raw[idx[1]][c] * 255.0f) * wt[1];
                    s = s + (raw[idx[2]][c] * 255.0f) * wt[2];
                    s = s + (raw[idx[3]][c] * 255.0f) * wt[3];
                } else {
                    s = fmaf(raw[idx[1]][c] * 255.0f, wt[1], s);
                    s = fmaf(raw[idx[2]][c] * 255.0f, wt[2], s);
                    s = fmaf(raw[idx[3]][c] * 255.0f, wt[3], s);
                }
                d[c] = s - e0[c];
            }
        }
        w = dot4(k, d);
        lo = fminf(w, lo);
        hi = fmaxf(w, hi);
        projw[i] = w;
    }
    span = fmaxf(SMALL_VALUE, hi - lo);
    span = astc_oracle_mufu_rcp(span);                          /* 1.0f / invlen (:366) */
    for (i = 0; i < 16; ++i) projw[i] = (projw[i] - lo) * span;
}

/* encode_block (ASTC_Encode.hlsl:510-550). */
void astc_oracle_encode_block(const float (*raw)[4], const astc_oracle_opt *opt,
                              uint8_t out[16], astc_oracle_trace *tr)
{
    const int dim = opt->block_dim == 6 ? 6 : 4, bs = dim * dim;
    const int has_alpha = opt->has_alpha ? 1 : 0;
    const int wq = has_alpha ? QUANT_6 : QUANT_12;          /* :518-522 */
    const uint32_t range1 = has_alpha ? 5u : 11u;           /* :540 weight_range-1 */
    float e0[4], e1[4], projw[16];
    uint8_t ep[8], q[16], qs[16], ep_ise[16], wt_ise[16];
    int i;

    ensure_tables();
    pca_endpoints(raw, bs, has_alpha, opt->axis_method == 1 ? 1 : 0, e0, e1, tr);

    /* encode_color (:233-245) + endpoint_ise (:475-489) */
    for (i = 0; i < 4; ++i) {
        ep[2 * i]     = (uint8_t)(uint32_t)rintf(e0[i]);
        ep[2 * i + 1] = (uint8_t)(uint32_t)rintf(e1[i]);
    }
    if (!has_alpha) { ep[6] = 0; ep[7] = 0; }
    memset(ep_ise, 0, sizeof ep_ise);
    astc_oracle_bise_encode(ep, has_alpha ? 8 : 6, QUANT_256, ep_ise);

    /* weight_ise (:491-508) */
    project_weights(raw, dim, e0, e1, projw);
    for (i = 0; i < 16; ++i) {
        float f = rintf(projw[i] * (float)range1);          /* :256-260 */
        uint32_t u = (uint32_t)f;
        if (u > range1) u = range1;
        q[i] = (uint8_t)u;
        qs[i] = g_scramble[wq][u];
    }
    memset(wt_ise, 0, sizeof wt_ise);
    astc_oracle_bise_encode(qs, 16, wq, wt_ise);

    assemble(astc_oracle_blockmode(wq),
             has_alpha ? CEM_LDR_RGBA_DIRECT : CEM_LDR_RGB_DIRECT, ep_ise, wt_ise, out);

    if (tr) {
        memcpy(tr->e0, e0, sizeof e0);
        memcpy(tr->e1, e1, sizeof e1);
        memcpy(tr->ep, ep, sizeof ep);
        memcpy(tr->projw, projw, sizeof projw);
        memcpy(tr->q, q, sizeof q);
        memcpy(tr->qs, qs, sizeof qs);
    }
}

/* UNORM8 -> float as the D3D11 texture unit defines it (main.cpp:38): c/255
 * for UNORM; for UNORM_SRGB the rgb channels go through the sRGB transfer
 * function (D3D11 functional spec 3.2.x), alpha stays linear. */
void astc_oracle_unorm_lut(int srgb, float out[256])
{
    int c;
    for (c = 0; c < 256; ++c) {
        if (!srgb) {
            out[c] = (float)c / 255.0f;
        } else {
            double x = (double)c / 255.0;
            double y = x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4);
            out[c] = (float)y;
            if (g_variant & ASTC_ORACLE_VAR_SRGB_POWF) {
                const float xf = (float)c / 255.0f;
                out[c] = xf <= 0.04045f ? xf / 12.92f : powf((xf + 0.055f) / 1.055f, 2.4f);
            }
        }
    }
}

/* MainCS (ASTC_Encode.hlsl:553-582): block b -> (bx,by) row-major; texel k at
 * (bx*D + k%D, by*D + k/D); out-of-range Load() returns 0; normal maps force
 * b = a = 1 even on padded texels (:575-578). */
int astc_oracle_encode_rows(const uint8_t *rgba, int width, int height, size_t pitch,
                            const astc_oracle_opt *opt, int row0, int row1,
                            uint8_t *blocks, int threads)
{
    const int dim = opt->block_dim == 6 ? 6 : 4;
    const int bw = (width + dim - 1) / dim;
    const int use_srgb = opt->srgb && !opt->is_normal_map;     /* main.cpp:214 */
    float lut_rgb[256], lut_a[256];
    int used = 1, by;

    ensure_tables();
    astc_oracle_unorm_lut(use_srgb, lut_rgb);
    astc_oracle_unorm_lut(0, lut_a);
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
    used = threads;
#else
    (void)threads;
#endif

#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (by = row0; by < row1; ++by) {
        int bx, k;
        for (bx = 0; bx < bw; ++bx) {
            float raw[36][4];
            for (k = 0; k < dim * dim; ++k) {
                int x = bx * dim + k % dim, y = by * dim + k / dim;
                if (x < width && y < height) {
                    const uint8_t *p = rgba + (size_t)y * pitch + (size_t)x * 4u;
                    raw[k][0] = lut_rgb[p[0]]; raw[k][1] = lut_rgb[p[1]];
                    raw[k][2] = lut_rgb[p[2]]; raw[k][3] = lut_a[p[3]];
                } else {
                    raw[k][0] = raw[k][1] = raw[k][2] = raw[k][3] = 0.0f;
                }
                if (opt->is_normal_map) { raw[k][2] = 1.0f; raw[k][3] = 1.0f; }
            }
            astc_oracle_encode_block((const float (*)[4])raw, opt,
                                     blocks + 16u * ((size_t)(by - row0) * (size_t)bw + (size_t)bx), NULL);
        }
    }
    return used;
}

int astc_oracle_encode_image(const uint8_t *rgba, int width, int height, size_t pitch,
                             const astc_oracle_opt *opt, uint8_t *blocks, int threads)
{
    const int dim = opt->block_dim == 6 ? 6 : 4;
    return astc_oracle_encode_rows(rgba, width, height, pitch, opt, 0,
                                   (height + dim - 1) / dim, blocks, threads);
}

/* ------------------------------------------------------------------ */
/* Subset decoder (ASTC specification C.2; not part of the reference)  */
/* ------------------------------------------------------------------ */

static uint32_t get_bits(const uint8_t s[16], uint32_t pos, uint32_t count)
{
    uint32_t v = 0, i;
    for (i = 0; i < count; ++i) {
        uint32_t p = pos + i;
        if (p < 128u && ((s[p >> 3] >> (p & 7u)) & 1u)) v |= 1u << i;
    }
    return v;
}

static void bise_decode(const uint8_t s[16], int count, int quant, uint8_t *values)
{
    int bits, trits, quints, i, j;
    uint32_t pos = 0;
    astc_oracle_quant_layout(quant, &bits, &trits, &quints);
    if (trits) {
        static const int tshift[5] = {0, 2, 4, 5, 7}, tcount[5] = {2, 2, 1, 2, 1};
        for (i = 0; i < count; i += 5) {
            uint32_t m[5], T = 0;
            int t[5];
            for (j = 0; j < 5; ++j) {
                m[j] = get_bits(s, pos, (uint32_t)bits); pos += (uint32_t)bits;
                T |= get_bits(s, pos, (uint32_t)tcount[j]) << tshift[j]; pos += (uint32_t)tcount[j];
            }
            spec_trits_from_integer((int)T, t);
            for (j = 0; j < 5 && i + j < count; ++j)
                values[i + j] = (uint8_t)(((uint32_t)t[j] << bits) | m[j]);
        }
    } else if (quints) {
        static const int qshift[3] = {0, 3, 5}, qcount[3] = {3, 2, 2};
        for (i = 0; i < count; i += 3) {
            uint32_t m[3], Q = 0;
            int q[3];
            for (j = 0; j < 3; ++j) {
                m[j] = get_bits(s, pos, (uint32_t)bits); pos += (uint32_t)bits;
                Q |= get_bits(s, pos, (uint32_t)qcount[j]) << qshift[j]; pos += (uint32_t)qcount[j];
            }
            spec_quints_from_integer((int)Q, q);
            for (j = 0; j < 3 && i + j < count; ++j)
                values[i + j] = (uint8_t)(((uint32_t)q[j] << bits) | m[j]);
        }
    } else {
        for (i = 0; i < count; ++i) {
            values[i] = (uint8_t)get_bits(s, pos, (uint32_t)bits); pos += (uint32_t)bits;
        }
    }
}

/* ASTC spec table C.2.8 (2D block modes).  Returns 0 when decodable here. */
static int parse_blockmode(uint32_t mode, uint32_t *gw, uint32_t *gh, uint32_t *wq, uint32_t *dual)
{
    uint32_t R, H = (mode >> 9) & 1u, D = (mode >> 10) & 1u;
    uint32_t A = (mode >> 5) & 3u, B = (mode >> 7) & 3u;
    if ((mode & 3u) != 0u) {
        R = ((mode & 3u) << 1) | ((mode >> 4) & 1u);
        switch ((mode >> 2) & 3u) {
        case 0: *gw = B + 4; *gh = A + 2; break;
        case 1: *gw = B + 8; *gh = A + 2; break;
        case 2: *gw = A + 2; *gh = B + 8; break;
        default:
            if ((B & 2u) == 0u) { *gw = A + 2; *gh = (B & 1u) + 6; }
            else { *gw = (B & 1u) + 2; *gh = A + 2; }
            break;
        }
    } else {
        if (((mode >> 2) & 3u) == 0u) return -1;            /* reserved / void extent */
        R = (((mode >> 2) & 3u) << 1) | ((mode >> 4) & 1u);
        switch (B) {
        case 0: *gw = 12; *gh = A + 2; break;
        case 1: *gw = A + 2; *gh = 12; break;
        case 3:
            if (A == 0) { *gw = 6; *gh = 10; }
            else if (A == 1) { *gw = 10; *gh = 6; }
            else return -1;
            break;
        default:
            *gw = A + 6; *gh = ((mode >> 9) & 3u) + 6; H = 0; D = 0;
            break;
        }
    }
    if (R < 2u) return -1;
    *wq = (R - 2u) + 6u * H;
    *dual = D;
    return 0;
}

int astc_oracle_unpack_block(const uint8_t block[16], astc_oracle_symbolic *sym)
{
    uint32_t gw, gh, wq, dual, nweights, wbits, avail;
    uint8_t rev[16], enc[64];
    int i, nvals, epq, q;
    ensure_tables();
    memset(sym, 0, sizeof *sym);
    sym->mode = get_bits(block, 0, 11);
    sym->partitions = get_bits(block, 11, 2) + 1u;
    sym->cem = get_bits(block, 13, 4);
    if (parse_blockmode(sym->mode, &gw, &gh, &wq, &dual) != 0) return -1;
    if (dual || sym->partitions != 1u) return -2;
    if (sym->cem != CEM_LDR_RGB_DIRECT && sym->cem != CEM_LDR_RGBA_DIRECT) return -3;
    nweights = gw * gh;
    if (nweights > 64u) return -4;
    wbits = astc_oracle_ise_bitcount(nweights, (int)wq);
    if (wbits < 24u || wbits > 96u) return -4;
    sym->weight_quant = wq; sym->grid_w = gw; sym->grid_h = gh;

    /* weights are stored bit-reversed from the top of the block */
    for (i = 0; i < 16; ++i) {
        uint8_t b = block[15 - i];
        b = (uint8_t)(((b & 0xF0u) >> 4) | ((b & 0x0Fu) << 4));
        b = (uint8_t)(((b & 0xCCu) >> 2) | ((b & 0x33u) << 2));
        b = (uint8_t)(((b & 0xAAu) >> 1) | ((b & 0x55u) << 1));
        rev[i] = b;
    }
    bise_decode(rev, (int)nweights, (int)wq, enc);
    for (i = 0; i < (int)nweights; ++i) {
        sym->weights[i] = g_unscramble[wq][enc[i] & 31];
        sym->weights_unq[i] = g_weight_unq[wq][enc[i] & 31];
    }

    /* endpoint quant = the largest level that fits the remaining bits */
    nvals = sym->cem == CEM_LDR_RGBA_DIRECT ? 8 : 6;
    avail = 128u - 17u - wbits;
    epq = -1;
    for (q = QUANT_256; q >= 0; --q)
        if (astc_oracle_ise_bitcount((uint32_t)nvals, q) <= avail) { epq = q; break; }
    if (epq != QUANT_256) return -5;               /* only 8-bit endpoints handled */
    for (i = 0; i < nvals; ++i) sym->ep[i] = (uint8_t)get_bits(block, 17u + 8u * (uint32_t)i, 8);
    if (nvals == 6) { sym->ep[6] = 255; sym->ep[7] = 255; }
    return 0;
}

static void decode_block(const uint8_t block[16], int dim, uint8_t texels[36][4], int *bad)
{
    astc_oracle_symbolic s;
    int e0[4], e1[4], c, x, y;
    if (astc_oracle_unpack_block(block, &s) != 0) {
        for (x = 0; x < dim * dim; ++x) {
            texels[x][0] = 255; texels[x][1] = 0; texels[x][2] = 255; texels[x][3] = 255;
        }
        ++*bad;
        return;
    }
    /* CEM 8 / 12 (spec C.2.14), with blue contraction when s1 < s0 */
    if ((int)s.ep[1] + s.ep[3] + s.ep[5] >= (int)s.ep[0] + s.ep[2] + s.ep[4]) {
        for (c = 0; c < 4; ++c) { e0[c] = s.ep[2 * c]; e1[c] = s.ep[2 * c + 1]; }
    } else {
        e0[0] = (s.ep[1] + s.ep[5]) >> 1; e0[1] = (s.ep[3] + s.ep[5]) >> 1; e0[2] = s.ep[5]; e0[3] = s.ep[7];
        e1[0] = (s.ep[0] + s.ep[4]) >> 1; e1[1] = (s.ep[2] + s.ep[4]) >> 1; e1[2] = s.ep[4]; e1[3] = s.ep[6];
    }
    for (y = 0; y < dim; ++y) {
        for (x = 0; x < dim; ++x) {
            /* weight infill (spec C.2.18) */
            int Ds = (1024 + dim / 2) / (dim - 1);
            int cs = Ds * x, ct = Ds * y;
            int gs = (cs * ((int)s.grid_w - 1) + 32) >> 6, gt = (ct * ((int)s.grid_h - 1) + 32) >> 6;
            int js = gs >> 4, fs = gs & 15, jt = gt >> 4, ft = gt & 15;
            int v0 = js + jt * (int)s.grid_w, n = (int)(s.grid_w * s.grid_h);
            int w11 = (fs * ft + 8) >> 4, w10 = ft - w11, w01 = fs - w11, w00 = 16 - fs - ft + w11;
            int p00 = s.weights_unq[v0];
            int p01 = v0 + 1 < n ? s.weights_unq[v0 + 1] : 0;
            int p10 = v0 + (int)s.grid_w < n ? s.weights_unq[v0 + s.grid_w] : 0;
            int p11 = v0 + (int)s.grid_w + 1 < n ? s.weights_unq[v0 + s.grid_w + 1] : 0;
            int w = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + 8) >> 4;
            for (c = 0; c < 4; ++c) {
                int c0 = (e0[c] << 8) | e0[c], c1 = (e1[c] << 8) | e1[c];
                int v = (c0 * (64 - w) + c1 * w + 32) >> 6;
                texels[y * dim + x][c] = (uint8_t)(v >> 8);
            }
        }
    }
}

int astc_oracle_decode_image(const uint8_t *blocks, int width, int height, int block_dim,
                             uint8_t *rgba, size_t pitch)
{
    const int dim = block_dim == 6 ? 6 : 4;
    const int bw = (width + dim - 1) / dim, bh = (height + dim - 1) / dim;
    int bad = 0, by;
    ensure_tables();
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+:bad)
#endif
    for (by = 0; by < bh; ++by) {
        int bx, k;
        for (bx = 0; bx < bw; ++bx) {
            uint8_t t[36][4];
            decode_block(blocks + 16u * ((size_t)by * (size_t)bw + (size_t)bx), dim, t, &bad);
            for (k = 0; k < dim * dim; ++k) {
                int x = bx * dim + k % dim, y = by * dim + k / dim;
                if (x < width && y < height) memcpy(rgba + (size_t)y * pitch + (size_t)x * 4u, t[k], 4);
            }
        }
    }
    return bad;
}
