/*
 * astc_oracle.h -- CPU restatement of the niepp/astc_encoder per-block hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it
 * may be imported / linked / executed only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs, always as the checker
 * or the reported CPU baseline, never as the thing shipped.  The product
 * (astc_encoder_b200/) has no CPU fallback and never links this file.
 *
 * What is restated (file:line relative to the reference checkout):
 *   ASTC_Encode.hlsl:553-582   MainCS texel fetch           -> astc_oracle_encode_image
 *   ASTC_Encode.hlsl:139-168   principal_component_analysis -> pca_endpoints
 *   ASTC_Encode.hlsl:93-106    eigen_vector                 -> power_iteration
 *   ASTC_Encode.hlsl:108-137   find_min_max                 -> pca_endpoints
 *   ASTC_Encode.hlsl:170-227   max_accumulation_pixel_direction (the alternative axis heuristic the
 *                              reference carries with its call commented out at :514) -> accumulation_axis
 *   ASTC_Encode.hlsl:233-245   encode_color                 -> astc_oracle_encode_block
 *   ASTC_Encode.hlsl:316-393   weights                      -> project_weights
 *   ASTC_Encode.hlsl:400-473   assemble_block/_blockmode    -> assemble
 *   ASTC_IntegerSequenceEncoding.hlsl:1-277                 -> astc_oracle_bise_encode
 *   ASTC_Table.hlsl:1-67       scramble_table               -> astc_oracle_scramble
 *
 * Parity pinning: the reference (HLSL cs_5_0 through D3D11) cannot run on
 * Linux, so the oracle is pinned against the reference's one committed golden
 * vector, textures/leaf.png -> textures/leaf.astc (copied as data into
 * tests/golden/).  See DESIGN.md "Oracle" for the measured match rate.
 * RGB mode, 6x6, normal-map and sRGB are NOT pinned by any reference fixture
 * ("parity unpinned" for those variants versus real D3D11 output).
 *
 * Canonical float arithmetic (frozen; the CUDA kernel reproduces it bit for
 * bit): IEEE binary32, round-to-nearest-even, explicit fmaf() exactly where
 * written, no other contraction (compile with -ffp-contract=off), rintf for HLSL
 * round(); the reciprocal (ASTC_Encode.hlsl:366) and the reciprocal square root
 * of normalize() (:103,332) are NVIDIA's MUFU.RCP / MUFU.RSQ approximations,
 * emulated exactly from tables captured on the device (astc_oracle_set_mufu_tables,
 * oracle/tables/, tools/gen_mufu_tables.py) -- the arithmetic the golden shows.
 */
#ifndef ASTC_ORACLE_H
#define ASTC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Delta tables of the MUFU emulation (int8: 2^23 entries for rcp, 2^24 for rsq); must be set
 * before any encode call -- without them the encoders produce NaN poison, never a fallback. */
void astc_oracle_set_mufu_tables(const int8_t *rcp_delta, const int8_t *rsq_delta);
float astc_oracle_mufu_rcp(float x);
float astc_oracle_mufu_rsq(float x);

/* Mirrors encode_option (astc_encode.h:14-28) after CLI resolution. */
typedef struct astc_oracle_opt {
    int block_dim;      /* 4 or 6                                         */
    int has_alpha;      /* HAS_ALPHA   (astc_encode.h:61)                 */
    int is_normal_map;  /* IS_NORMALMAP (astc_encode.h:59)                */
    int srgb;           /* texture format _UNORM_SRGB (main.cpp:38,214);  */
                        /* ignored when is_normal_map is set              */
    int axis_method;    /* 0: principal_component_analysis (:139-168, what the reference ships);          */
                        /* 1: max_accumulation_pixel_direction (:170-227, commented out at :514)           */
} astc_oracle_opt;

/* Sensitivity switches (tools/pin_sensitivity.py ONLY): each replaces one modelling choice that no
 * reference fixture pins by its plausible alternative, so the number of blocks it decides can be
 * measured.  0 = the canonical arithmetic; process-wide, not thread-safe against running encodes. */
enum {
    ASTC_ORACLE_VAR_TRUE_DIVISION = 1,   /* mean = sum / BS and cov / (BS-1) as IEEE divisions (:147,162), not x RN(1/n) */
    ASTC_ORACLE_VAR_UNFUSED_SAMPLE = 2,  /* 6x6 sample_texel (:307-314): every product and sum rounded, no FMA          */
    ASTC_ORACLE_VAR_SRGB_POWF = 4,       /* sRGB decode evaluated in float (powf) instead of double rounded once        */
    ASTC_ORACLE_VAR_EXACT_RCP_RSQ = 8,   /* correctly rounded 1/x and 1/sqrt(x) instead of the MUFU approximations     */
    ASTC_ORACLE_VAR_UNFUSED_DEV = 16     /* texel*255 rounded before the subtraction of mean / e0 (no FMA)             */
};
void astc_oracle_set_variant(unsigned flags);
/* tools/golden_residual.py ONLY: adds the given number of ulps to EVERY result of the emulated MUFU.RCP / MUFU.RSQ
 * (0, 0 = the units as captured).  Answers "would the golden's bits follow if the reciprocal unit of the GPU that
 * made it differed from the B200's in the last bit for this argument?" for the residual blocks. */
void astc_oracle_set_mufu_bias(int rcp_ulps, int rsq_ulps);

/* Diagnostics of one block encode (unrounded endpoints, raw weights). */
typedef struct astc_oracle_trace {
    float mean[4];
    float cov[16];
    float axis[4];
    float e0[4], e1[4];       /* unrounded, after swap / alpha forcing   */
    uint8_t ep[8];            /* r0 r1 g0 g1 b0 b1 a0 a1                 */
    float projw[16];          /* normalised weights in [0,1]             */
    uint8_t q[16];            /* quantised weights, natural order        */
    uint8_t qs[16];           /* after scramble                          */
} astc_oracle_trace;

/* UNORM8 -> float conversion table: c/255.0f, or the D3D sRGB->linear
 * formula evaluated in double then rounded to float (srgb != 0). */
void astc_oracle_unorm_lut(int srgb, float out[256]);

/* Encode one block from already-fetched UNORM floats raw[k][c] in [0,1],
 * k = y*dim + x (ASTC_Encode.hlsl:565-580 without the final *255).     */
void astc_oracle_encode_block(const float (*raw)[4], const astc_oracle_opt *opt,
                              uint8_t out[16], astc_oracle_trace *trace);

/* Whole image: rgba row-major, row 0 first (the caller applies the stb
 * vertical flip), pitch in bytes.  Writes 16*ceil(w/D)*ceil(h/D) bytes.
 * threads <= 0 means "all OpenMP threads".  Returns threads used.       */
int astc_oracle_encode_image(const uint8_t *rgba, int width, int height,
                             size_t pitch, const astc_oracle_opt *opt,
                             uint8_t *blocks, int threads);

/* Encode only block rows [row0, row1) -- used for bounded CPU samples. */
int astc_oracle_encode_rows(const uint8_t *rgba, int width, int height,
                            size_t pitch, const astc_oracle_opt *opt,
                            int row0, int row1, uint8_t *blocks, int threads);

/* ---- integer sequence encoding (ASTC_IntegerSequenceEncoding.hlsl) ---- */
/* quant: 0..20 = QUANT_2..QUANT_256 (ASTC_Encode.hlsl:47-67).            */
void astc_oracle_quant_layout(int quant, int *bits, int *trits, int *quints);
uint32_t astc_oracle_ise_bitcount(uint32_t items, int quant);
/* Appends count values LSB-first into a 128-bit little-endian stream,
 * padding the last trit/quint group with zeros.  Returns bits written
 * (whole groups, as the reference does).                                 */
uint32_t astc_oracle_bise_encode(const uint8_t *values, int count, int quant,
                                 uint8_t stream[16]);
uint8_t astc_oracle_integer_from_trits(int t0, int t1, int t2, int t3, int t4);
uint8_t astc_oracle_integer_from_quints(int q0, int q1, int q2);
/* scramble_table[method*32 + q] (ASTC_Table.hlsl:3-66), method 0..11.    */
uint8_t astc_oracle_scramble(int method, int q);
uint32_t astc_oracle_blockmode(int weight_quant);   /* ASTC_Encode.hlsl:446-473 */

/* ---- independent subset decoder (not in the reference; ASTC spec) ---- */
/* Decodes single-partition, single-plane LDR blocks with CEM 8 / 12 and
 * any first-row 2D block mode into RGBA8 (decode_unorm8 rules).  Unknown
 * blocks decode to magenta and are counted in the return value.          */
int astc_oracle_decode_image(const uint8_t *blocks, int width, int height,
                             int block_dim, uint8_t *rgba, size_t pitch);
/* Symbolic view of one block for parity diagnostics.  Returns 0 if ok.   */
typedef struct astc_oracle_symbolic {
    uint32_t mode, partitions, cem, weight_quant, grid_w, grid_h;
    uint8_t ep[8];
    uint8_t weights[64];     /* un-scrambled natural order 0..range-1     */
    uint8_t weights_unq[64]; /* unquantised 0..64                         */
} astc_oracle_symbolic;
int astc_oracle_unpack_block(const uint8_t block[16], astc_oracle_symbolic *sym);

#ifdef __cplusplus
}
#endif
#endif
