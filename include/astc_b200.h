/*
 * astc_b200.h -- C ABI of the B200-native ASTC block encoder.
 *
 * This is the drop-in boundary: the entry points a maintainer of
 * niepp/astc_encoder binds instead of the D3D11 device / shader-compile / UAV /
 * Dispatch / staging-buffer plumbing.  Plain pointers and sizes only; no C++
 * or torch types.  Every function returns ASTC_B200_OK (0) or a negative
 * astc_b200_status and never throws.  The library keeps no mutable global
 * state besides immutable lookup tables, so calls are re-entrant per
 * (device, stream); use one host thread or one process per GPU.
 *
 * Reference interfaces replaced (file:line in the reference checkout):
 *   encode_option                       astc_encode.h:14-28
 *   encode_astc() = compile+bind+Dispatch astc_encode.h:87-194
 *   CSConstantBuffer                    astc_encode.h:31-36  (plain arguments here)
 *   read_gpu()                          astc_save.h:34-50
 *   save_astc(), astc_header            astc_save.h:3-14,52-76
 *   load_tex() (stb load + flip + upload) main.cpp:19-56
 *   create_device_swapchain()           main.cpp:58-119      (astc_b200_set_device)
 * The C++ header-only mirrors with the reference's own names live in
 * include/astc_encode.h and include/astc_save.h.
 */
#ifndef ASTC_B200_H
#define ASTC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ASTC_B200_API __declspec(dllexport)
#else
#define ASTC_B200_API __attribute__((visibility("default")))
#endif

#define ASTC_B200_BLOCK_BYTES 16            /* BLOCK_BYTES, astc_encode.h:12 */
#define ASTC_B200_MAGIC 0x5CA1AB13u         /* MAGIC_FILE_CONSTANT, astc_save.h:3 */

typedef enum astc_b200_status {
    ASTC_B200_OK = 0,
    ASTC_B200_ERR_INVALID_ARGUMENT = -1,    /* null pointer, negative size, bad pitch */
    ASTC_B200_ERR_CUDA = -2,                /* a CUDA runtime call failed (see last_cuda_error) */
    ASTC_B200_ERR_NO_DEVICE = -3,           /* no sm_100 class device visible */
    ASTC_B200_ERR_OUT_OF_MEMORY = -4,
    ASTC_B200_ERR_IO = -5,                  /* file open / read / write failed */
    ASTC_B200_ERR_BAD_IMAGE = -6,           /* undecodable image or .astc file */
    ASTC_B200_ERR_UNSUPPORTED = -7
} astc_b200_status;

/* POD mirror of encode_option (astc_encode.h:14-28): same fields, same order,
 * same defaults when zero-initialised with astc_b200_option_default().
 * Block size: 6x6 when is6x6 is set or is4x4 is cleared, else 4x4.  (The
 * reference reads only is4x4, which its CLI can never clear, so -6x6 is dead
 * there; making it work is a documented deviation, see DESIGN.md.)
 * srgb is honoured only when is_normal_map is 0 (main.cpp:214). */
typedef struct astc_b200_option {
    uint8_t is4x4;
    uint8_t is6x6;
    uint8_t is_normal_map;
    uint8_t has_alpha;
    uint8_t srgb;
    /* Extension (not a field of the reference struct; 0 keeps the reference's behaviour): which of the
     * two axis heuristics of ASTC_Encode.hlsl picks the endpoint direction.
     *   0  principal_component_analysis  (:139-168, called at :515 -- what the reference ships)
     *   1  max_accumulation_pixel_direction (:170-227, its call is commented out at :514)           */
    uint8_t axis_method;
    uint8_t reserved[2];
} astc_b200_option;

/* One texture of a batch (mip level, array slice ...). All device pointers. */
typedef struct astc_b200_image {
    const uint8_t *d_rgba;      /* RGBA8, row-major, row 0 first                 */
    uint8_t *d_blocks;          /* 16 * ceil(w/D) * ceil(h/D) bytes              */
    size_t pitch_bytes;         /* >= 4*width                                    */
    int32_t width, height;
} astc_b200_image;

typedef struct astc_b200_batch astc_b200_batch;   /* opaque launch plan */

/* ---- library / device ------------------------------------------------- */
ASTC_B200_API const char *astc_b200_version(void);
ASTC_B200_API const char *astc_b200_strerror(int status);
ASTC_B200_API const char *astc_b200_last_cuda_error(void);  /* thread-local text */
ASTC_B200_API int astc_b200_device_count(int *count);
ASTC_B200_API int astc_b200_set_device(int ordinal);
ASTC_B200_API int astc_b200_device_info(int ordinal, char *name, size_t name_len,
                                        int *sm_count, int *cc_major, int *cc_minor,
                                        size_t *global_mem_bytes);

/* ---- geometry --------------------------------------------------------- */
ASTC_B200_API void astc_b200_option_default(astc_b200_option *opt);
ASTC_B200_API int astc_b200_block_dim(const astc_b200_option *opt);      /* 4 or 6 */
/* xBlockNum / yBlockNum / ByteWidth of astc_encode.h:124-129,143 */
ASTC_B200_API int astc_b200_block_counts(int width, int height, const astc_b200_option *opt,
                                         int *blocks_x, int *blocks_y);
ASTC_B200_API size_t astc_b200_output_size(int width, int height, const astc_b200_option *opt);
/* Multi-GPU band sharding: texel rows [*y0, *y0 + *rows) and the byte offset of
 * that band's blocks for part `part` of `parts` (block rows split evenly). */
ASTC_B200_API int astc_b200_band(int width, int height, const astc_b200_option *opt,
                                 int parts, int part, int *y0, int *rows,
                                 size_t *block_byte_offset, size_t *block_bytes);

/* ---- the hot path ----------------------------------------------------- */
/* Replaces encode_astc()'s Dispatch (astc_encode.h:190): asynchronous on
 * `cuda_stream` (a cudaStream_t, NULL = default stream); all pointers are
 * device pointers on the current device.  Texels outside width/height read
 * as 0 like Texture2D.Load (ASTC_Encode.hlsl:574).                        */
ASTC_B200_API int astc_b200_encode_device(const uint8_t *d_rgba, int width, int height,
                                          size_t pitch_bytes, const astc_b200_option *opt,
                                          uint8_t *d_blocks, void *cuda_stream);

/* load_tex upload + encode_astc + read_gpu in one synchronous call on host
 * buffers: banded H2D copy / kernel / D2H copy pipelined over internal
 * streams.  Pinned host memory (astc_b200_host_alloc) gives full PCIe rate;
 * pageable memory is staged through pinned slots by worker threads.        */
ASTC_B200_API int astc_b200_encode_host(const uint8_t *h_rgba, int width, int height,
                                        size_t pitch_bytes, const astc_b200_option *opt,
                                        uint8_t *h_blocks);

/* ---- persistent host-side context ---------------------------------------------------------------
 * What main.cpp creates once per process (device, :199-209) plus what it creates per encode (texture,
 * UAV, staging buffer: main.cpp:46-52, astc_encode.h:137-164, astc_save.h:19-32), kept alive across
 * calls: three streams, an event, a grow-only device workspace and pinned staging.  A context belongs
 * to the device that was current when it was created and to one host thread at a time.
 * astc_b200_encode_host() uses a lazily created thread-local context, so it pays no per-call setup either. */
typedef struct astc_b200_context astc_b200_context;
ASTC_B200_API int astc_b200_context_create(astc_b200_context **ctx);
ASTC_B200_API void astc_b200_context_destroy(astc_b200_context *ctx);
ASTC_B200_API int astc_b200_context_trim(astc_b200_context *ctx);     /* give the workspace back; it regrows on demand */
/* PAGEABLE host buffers (malloc / new[] / stbi_load -- what main.cpp:24,224 holds) of 1 MiB or more take a staged
 * pipeline: worker threads copy each band into pinned slots while the bands before it are on the link and on the SMs
 * (2-3x the driver's own pageable staging).  `threads` = workers besides the caller: -1 automatic (a quarter of the
 * host's hardware threads, 1..7; the default), 0 none.  Call between encodes. */
ASTC_B200_API int astc_b200_context_set_copy_threads(astc_b200_context *ctx, int threads);
ASTC_B200_API int astc_b200_context_encode_host(astc_b200_context *ctx, const uint8_t *h_rgba, int width,
                                                int height, size_t pitch_bytes,
                                                const astc_b200_option *opt, uint8_t *h_blocks);

/* One texture of a host-memory batch: host pointers (pinned gives full PCIe rate, pageable works). */
typedef struct astc_b200_host_image {
    const uint8_t *h_rgba;      /* RGBA8, row-major, row 0 first                 */
    uint8_t *h_blocks;          /* 16 * ceil(w/D) * ceil(h/D) bytes              */
    size_t pitch_bytes;         /* >= 4*width                                    */
    int32_t width, height;
} astc_b200_host_image;
/* Upload + encode + read back MANY textures (e.g. whole mip chains) in one synchronous call: levels under
 * 256 KiB are gathered through pinned staging, larger ones copied directly, ~32 MiB of source per
 * upload / launch / download group, groups pipelined over the context's streams. */
ASTC_B200_API int astc_b200_context_batch_encode_host(astc_b200_context *ctx,
                                                      const astc_b200_host_image *images, int count,
                                                      const astc_b200_option *opt);

/* The same for mip chains of which only the BASE levels exist on the host: each image is a base; it alone is uploaded,
 * the levels below it (down to 1x1, astc_b200_mip_chain_device's box filter) are produced on the device and encoded by
 * the same launch.  h_blocks receives the blocks of ALL levels, base first then level 1, 2, ...:
 * astc_b200_mip_chain_output_size bytes (`levels` counts the base).  Three quarters of the upload of a chain whose
 * levels were made on the host. */
ASTC_B200_API int astc_b200_context_batch_encode_mip_chains_host(astc_b200_context *ctx,
                                                                 const astc_b200_host_image *bases, int count,
                                                                 const astc_b200_option *opt);
ASTC_B200_API int astc_b200_mip_chain_output_size(int width, int height, const astc_b200_option *opt,
                                                  size_t *bytes, int *levels);

/* Many textures (mip chains) in ONE launch over a prefix-summed block table.
 * create() uploads the table; encode() is asynchronous and reusable.       */
ASTC_B200_API int astc_b200_batch_create(const astc_b200_image *images, int count,
                                         const astc_b200_option *opt, astc_b200_batch **out);
ASTC_B200_API int astc_b200_batch_encode(astc_b200_batch *batch, void *cuda_stream);
ASTC_B200_API int astc_b200_batch_total_blocks(const astc_b200_batch *batch, uint64_t *blocks,
                                               uint64_t *texels);
ASTC_B200_API void astc_b200_batch_destroy(astc_b200_batch *batch);

/* Kernels launched by this library since load (all threads); evidence for
 * bench.py's gpu_launches.                                                 */
ASTC_B200_API uint64_t astc_b200_launch_count(void);

/* ---- integer sequence encoding, exposed for known-answer tests --------- */
/* quant: 0..20 = QUANT_2..QUANT_256 (ASTC_Encode.hlsl:47-67).  nseq sequences
 * of `count` values each (<= 64) are packed LSB-first into 16-byte streams by
 * the same device code the encoder uses (ASTC_IntegerSequenceEncoding.hlsl:
 * 142-276 incl. the quint path).                                           */
ASTC_B200_API int astc_b200_bise_encode_device(const uint8_t *d_values, int count, int quant,
                                               int nseq, uint8_t *d_streams, void *cuda_stream);
/* Host views of the compile-time tables (no GPU needed). */
ASTC_B200_API int astc_b200_quant_layout(int quant, int *bits, int *trits, int *quints);
ASTC_B200_API uint32_t astc_b200_ise_bitcount(uint32_t items, int quant);
ASTC_B200_API int astc_b200_integer_from_trits(int t0, int t1, int t2, int t3, int t4);
ASTC_B200_API int astc_b200_integer_from_quints(int q0, int q1, int q2);
ASTC_B200_API int astc_b200_scramble(int method, int q);
ASTC_B200_API uint32_t astc_b200_blockmode(int weight_quant);
/* UNORM8 -> float table the kernel uses (srgb: D3D sRGB->linear). */
ASTC_B200_API int astc_b200_unorm_lut(int srgb, float out[256]);

/* ---- device decode (not in the reference; needed by the PSNR metric) --- */
/* Decodes the subset this encoder emits (1 partition, 1 plane, CEM 8/12,
 * 8-bit endpoints, any bit/trit weight range, 4x4 or 6x6 blocks) to RGBA8. */
ASTC_B200_API int astc_b200_decode_device(const uint8_t *d_blocks, int width, int height,
                                          int block_dim, uint8_t *d_rgba, size_t pitch_bytes,
                                          void *cuda_stream);

/* ---- mip generation on the device (not in the reference, whose caller -----
 *      main.cpp:19-56 load_tex -- uploads one level) ---------------------------------------- */
/* Next mip level of an RGBA8 image by 2x2 box filter, (sum + 2) >> 2 per channel; output is
 * max(1, width/2) x max(1, height/2) (an odd trailing row / column is dropped). Async on the
 * stream; device pointers, 4-byte aligned (16-byte alignment of bases and pitches and an output
 * width that is a multiple of 4 take the vector path). */
ASTC_B200_API int astc_b200_downsample2x2_device(const uint8_t *d_src, int width, int height,
                                                 size_t src_pitch_bytes, uint8_t *d_dst,
                                                 size_t dst_pitch_bytes, void *cuda_stream);

/* The whole chain below a base image in ONE call: levels 1, 2, ... down to 1x1, each the 2x2 box filter of the level
 * before it (the same bytes as repeated astc_b200_downsample2x2_device calls), written into one arena.
 * astc_b200_mip_chain_layout tells where: level l+1 is widths[l] x heights[l], rows tightly packed, at byte offset
 * offsets[l] (a multiple of 256); total_bytes includes a 256-byte scratch tail.  The arrays must hold 24 entries.
 * When width and height are multiples of 64 (and the base is 16-byte aligned) the chain is ONE kernel launch (each
 * CTA reduces a 64x64 tile through six levels, the last CTA to finish reduces the rest); otherwise one launch per
 * level, issued back to back.  d_levels must be 256-byte aligned.  Async on the stream. */
ASTC_B200_API int astc_b200_mip_chain_layout(int width, int height, int *levels, size_t *offsets,
                                             int *widths, int *heights, size_t *total_bytes);
ASTC_B200_API int astc_b200_mip_chain_device(const uint8_t *d_base, int width, int height,
                                             size_t pitch_bytes, uint8_t *d_levels, size_t levels_bytes,
                                             void *cuda_stream);

/* ---- the hardware approximations the arithmetic is defined on ------------------------------ */
/* y[i] = rcp.approx.ftz.f32(x[i]) (op 0) or rsqrt.approx.ftz.f32(x[i]) (op 1): what the reference's
 * `1.0f / x` (ASTC_Encode.hlsl:366) and normalize() (:103,332) execute on the hardware its golden
 * output came from. For tests and for generating the oracle's tables. */
ASTC_B200_API int astc_b200_mufu_device(int op, const float *d_x, float *d_y, size_t count,
                                        void *cuda_stream);

/* ---- memory / streams (so hosts need not link the CUDA runtime) -------- */
ASTC_B200_API int astc_b200_malloc_device(void **d_ptr, size_t bytes);
ASTC_B200_API int astc_b200_free_device(void *d_ptr);
ASTC_B200_API int astc_b200_host_alloc(void **h_ptr, size_t bytes);     /* pinned */
ASTC_B200_API int astc_b200_host_free(void *h_ptr);
ASTC_B200_API int astc_b200_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *cuda_stream);
ASTC_B200_API int astc_b200_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *cuda_stream);
ASTC_B200_API int astc_b200_memcpy2d_h2d(void *d_dst, size_t d_pitch, const void *h_src, size_t h_pitch,
                                         size_t row_bytes, size_t rows, void *cuda_stream);
ASTC_B200_API int astc_b200_stream_create(void **cuda_stream);
ASTC_B200_API int astc_b200_stream_destroy(void *cuda_stream);
ASTC_B200_API int astc_b200_stream_synchronize(void *cuda_stream);

/* ---- host file formats -------------------------------------------------- */
/* astc_header + save_astc (astc_save.h:5-14,52-76), byte-exact. */
ASTC_B200_API int astc_b200_save_astc(const char *path, int xdim, int ydim, int xsize, int ysize,
                                      const uint8_t *blocks, size_t bufsz);
/* Multi-GPU / multi-process variant of save_astc: writes `nbytes` of blocks at `block_byte_offset` of the
 * payload (astc_b200_band) of the .astc file for an xsize x ysize image, creating and sizing the file if
 * need be; with write_header != 0 also the 16-byte header.  Never truncates existing content, so the
 * ranks sharing one file may call it in any order without a barrier (each writes a disjoint range). */
ASTC_B200_API int astc_b200_save_astc_slice(const char *path, int xdim, int ydim, int xsize, int ysize,
                                            size_t block_byte_offset, const uint8_t *blocks, size_t nbytes,
                                            int write_header);
/* Matching reader (2-D files: blockdim_z == 1 and zsize == 1). *blocks is malloc'ed; release with
 * astc_b200_free_host_buffer. */
ASTC_B200_API int astc_b200_load_astc(const char *path, int *xdim, int *ydim, int *xsize, int *ysize,
                                      uint8_t **blocks, size_t *bufsz);
/* stbi_load(..., STBI_rgb_alpha) with optional vertical flip (main.cpp:24-25), every format the
 * reference's stb_image v2.22 reads and with its results: JPEG (baseline + progressive), PNG, BMP, GIF
 * (first frame), PSD, PIC, PPM/PGM, Radiance HDR (tone-mapped like stbi_load), TGA. */
ASTC_B200_API int astc_b200_load_image(const char *path, int flip_vertically, int *width, int *height,
                                       int *components_in_file, uint8_t **rgba);
ASTC_B200_API const char *astc_b200_image_failure_reason(void);
ASTC_B200_API void astc_b200_free_host_buffer(void *p);

#ifdef __cplusplus
}
#endif
#endif /* ASTC_B200_H */
