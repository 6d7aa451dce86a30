// astc_capi.cu -- the C ABI declared in include/astc_b200.h.
//
// Thin layer: argument checks, geometry (astc_encode.h:124-134), kernel launch,
// and the CUDA replacements of the reference's D3D11 plumbing -- texture upload
// (main.cpp:46-52), UAV buffer (astc_encode.h:137-164) and staging read-back
// (astc_save.h:19-50).  There is no CPU encode path: without a CUDA device every
// compute entry point fails with ASTC_B200_ERR_NO_DEVICE / ASTC_B200_ERR_CUDA.
#include "astc_b200.h"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/types.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "astc_capi_internal.h"
#include "astc_kernels.h"
#include "astc_tables.h"
#include "astc_save.h"

namespace {
thread_local std::string g_last_cuda_error;
std::atomic<uint64_t> g_launches{0};
}  // namespace

namespace astc_capi {
int cuda_fail(cudaError_t e, const char *what)
{
    g_last_cuda_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return ASTC_B200_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return ASTC_B200_ERR_OUT_OF_MEMORY;
    return ASTC_B200_ERR_CUDA;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace astc_capi

using astc_capi::cuda_fail;
using astc_capi::dim_of;
using astc_capi::make_desc;

namespace {

// Shared argument validation of every image-shaped entry point.
int check_image(const void *rgba, int width, int height, size_t pitch, const astc_b200_option *opt, const void *blocks)
{
    if (!opt || width < 0 || height < 0 || opt->axis_method > 1) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (width == 0 || height == 0) return ASTC_B200_OK;
    if (!rgba || !blocks) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (pitch < size_t(width) * 4u || pitch % 4u != 0 || reinterpret_cast<uintptr_t>(rgba) % 4u != 0)
        return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(blocks) % 16u != 0) return ASTC_B200_ERR_INVALID_ARGUMENT;
    return ASTC_B200_OK;
}

}  // namespace

struct astc_b200_batch {
    std::vector<astc::ImageDesc> host;
    astc::ImageDesc *device = nullptr;
    astc_b200_option opt{};
    uint64_t blocks = 0, texels = 0;
    int device_ordinal = 0;
};

extern "C" {

const char *astc_b200_version(void) { return "astc_encoder_b200 0.1.0 (sm_100a)"; }

const char *astc_b200_strerror(int status)
{
    switch (status) {
    case ASTC_B200_OK: return "ok";
    case ASTC_B200_ERR_INVALID_ARGUMENT: return "invalid argument";
    case ASTC_B200_ERR_CUDA: return "CUDA runtime error";
    case ASTC_B200_ERR_NO_DEVICE: return "no CUDA device";
    case ASTC_B200_ERR_OUT_OF_MEMORY: return "out of memory";
    case ASTC_B200_ERR_IO: return "file I/O error";
    case ASTC_B200_ERR_BAD_IMAGE: return "bad image or .astc file";
    case ASTC_B200_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
    }
}

const char *astc_b200_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

int astc_b200_device_count(int *count)
{
    if (!count) return ASTC_B200_ERR_INVALID_ARGUMENT;
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount"); }
    return ASTC_B200_OK;
}

int astc_b200_set_device(int ordinal)
{
    CUDA_TRY(cudaSetDevice(ordinal));
    return ASTC_B200_OK;
}

int astc_b200_device_info(int ordinal, char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor,
                          size_t *global_mem_bytes)
{
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, ordinal));
    if (name && name_len) { std::strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (global_mem_bytes) *global_mem_bytes = prop.totalGlobalMem;
    return ASTC_B200_OK;
}

void astc_b200_option_default(astc_b200_option *opt)
{
    if (!opt) return;
    std::memset(opt, 0, sizeof *opt);
    opt->is4x4 = 1;                                  // encode_option() : is4x4(true) (astc_encode.h:21)
}

int astc_b200_block_dim(const astc_b200_option *opt) { return opt ? dim_of(opt) : 4; }

int astc_b200_block_counts(int width, int height, const astc_b200_option *opt, int *blocks_x, int *blocks_y)
{
    if (!opt || width < 0 || height < 0) return ASTC_B200_ERR_INVALID_ARGUMENT;
    const int d = dim_of(opt);
    if (blocks_x) *blocks_x = (width + d - 1) / d;
    if (blocks_y) *blocks_y = (height + d - 1) / d;
    return ASTC_B200_OK;
}

size_t astc_b200_output_size(int width, int height, const astc_b200_option *opt)
{
    int bx = 0, by = 0;
    if (astc_b200_block_counts(width, height, opt, &bx, &by) != ASTC_B200_OK) return 0;
    return size_t(bx) * size_t(by) * ASTC_B200_BLOCK_BYTES;
}

int astc_b200_band(int width, int height, const astc_b200_option *opt, int parts, int part, int *y0, int *rows,
                   size_t *block_byte_offset, size_t *block_bytes)
{
    if (!opt || width < 0 || height < 0 || parts <= 0 || part < 0 || part >= parts) return ASTC_B200_ERR_INVALID_ARGUMENT;
    const int d = dim_of(opt);
    const int64_t bx = (width + d - 1) / d, by = (height + d - 1) / d;
    const int64_t r0 = by * part / parts, r1 = by * (part + 1) / parts;
    const int64_t ty0 = std::min<int64_t>(r0 * d, height), ty1 = std::min<int64_t>(r1 * d, height);
    if (y0) *y0 = int(ty0);
    if (rows) *rows = int(ty1 - ty0);
    if (block_byte_offset) *block_byte_offset = size_t(r0 * bx) * ASTC_B200_BLOCK_BYTES;
    if (block_bytes) *block_bytes = size_t((r1 - r0) * bx) * ASTC_B200_BLOCK_BYTES;
    return ASTC_B200_OK;
}

int astc_b200_encode_device(const uint8_t *d_rgba, int width, int height, size_t pitch_bytes,
                            const astc_b200_option *opt, uint8_t *d_blocks, void *cuda_stream)
{
    const int rc = check_image(d_rgba, width, height, pitch_bytes, opt, d_blocks);
    if (rc != ASTC_B200_OK) return rc;
    if (width == 0 || height == 0) return ASTC_B200_OK;
    const int d = dim_of(opt);
    astc::EncodeParams p{};
    p.single = make_desc(d_rgba, d_blocks, pitch_bytes, width, height, d, 0);
    p.table = nullptr;
    p.count = 1;
    p.total_blocks = uint64_t(p.single.blocks_x) * uint64_t((height + d - 1) / d);
    CUDA_TRY(astc::launch_encode(d, opt->has_alpha != 0, opt->is_normal_map != 0, opt->srgb != 0, opt->axis_method, p,
                                 static_cast<cudaStream_t>(cuda_stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return ASTC_B200_OK;
}

int astc_b200_batch_create(const astc_b200_image *images, int count, const astc_b200_option *opt, astc_b200_batch **out)
{
    if (!out) return ASTC_B200_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!opt || count < 0 || (count > 0 && !images)) return ASTC_B200_ERR_INVALID_ARGUMENT;
    astc_b200_batch *b = new (std::nothrow) astc_b200_batch();
    if (!b) return ASTC_B200_ERR_OUT_OF_MEMORY;
    b->opt = *opt;
    const int d = dim_of(opt);
    for (int i = 0; i < count; ++i) {
        const astc_b200_image &im = images[i];
        const int rc = check_image(im.d_rgba, im.width, im.height, im.pitch_bytes, opt, im.d_blocks);
        if (rc != ASTC_B200_OK) { delete b; return rc; }
        if (im.width == 0 || im.height == 0) continue;
        astc::ImageDesc desc = make_desc(im.d_rgba, im.d_blocks, im.pitch_bytes, im.width, im.height, d, b->blocks);
        b->blocks += uint64_t(desc.blocks_x) * uint64_t((im.height + d - 1) / d);
        b->texels += uint64_t(im.width) * uint64_t(im.height);
        b->host.push_back(desc);
    }
    cudaError_t e = cudaGetDevice(&b->device_ordinal);
    if (e == cudaSuccess && !b->host.empty()) {
        e = cudaMalloc((void **)&b->device, b->host.size() * sizeof(astc::ImageDesc));
        if (e == cudaSuccess)
            e = cudaMemcpy(b->device, b->host.data(), b->host.size() * sizeof(astc::ImageDesc), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        if (b->device) cudaFree(b->device);
        delete b;
        return cuda_fail(e, "astc_b200_batch_create");
    }
    *out = b;
    return ASTC_B200_OK;
}

int astc_b200_batch_encode(astc_b200_batch *batch, void *cuda_stream)
{
    if (!batch) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (batch->host.empty()) return ASTC_B200_OK;
    int current = -1;
    CUDA_TRY(cudaGetDevice(&current));
    if (current != batch->device_ordinal) return ASTC_B200_ERR_INVALID_ARGUMENT;   // the table and the textures live on the device it was created on
    astc::EncodeParams p{};
    p.single = batch->host[0];
    p.table = batch->device;
    p.count = int(batch->host.size());
    p.total_blocks = batch->blocks;
    CUDA_TRY(astc::launch_encode(dim_of(&batch->opt), batch->opt.has_alpha != 0, batch->opt.is_normal_map != 0,
                                 batch->opt.srgb != 0, batch->opt.axis_method, p, static_cast<cudaStream_t>(cuda_stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return ASTC_B200_OK;
}

int astc_b200_batch_total_blocks(const astc_b200_batch *batch, uint64_t *blocks, uint64_t *texels)
{
    if (!batch) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (blocks) *blocks = batch->blocks;
    if (texels) *texels = batch->texels;
    return ASTC_B200_OK;
}

void astc_b200_batch_destroy(astc_b200_batch *batch)
{
    if (!batch) return;
    if (batch->device) cudaFree(batch->device);
    delete batch;
}

uint64_t astc_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int astc_b200_bise_encode_device(const uint8_t *d_values, int count, int quant, int nseq, uint8_t *d_streams,
                                 void *cuda_stream)
{
    if (count < 0 || count > 64 || quant < 0 || quant >= astc::QUANT_MAX || nseq < 0) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (nseq == 0) return ASTC_B200_OK;
    if (!d_streams || (count > 0 && !d_values)) return ASTC_B200_ERR_INVALID_ARGUMENT;
    CUDA_TRY(astc::launch_bise(d_values, count, quant, nseq, d_streams, static_cast<cudaStream_t>(cuda_stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return ASTC_B200_OK;
}

int astc_b200_quant_layout(int quant, int *bits, int *trits, int *quints)
{
    if (quant < 0 || quant >= astc::QUANT_MAX) return ASTC_B200_ERR_INVALID_ARGUMENT;
    const astc::QuantLayout l = astc::quant_layout(quant);
    if (bits) *bits = l.bits;
    if (trits) *trits = l.trits;
    if (quints) *quints = l.quints;
    return ASTC_B200_OK;
}

uint32_t astc_b200_ise_bitcount(uint32_t items, int quant) { return astc::ise_bitcount(items, quant); }

int astc_b200_integer_from_trits(int t0, int t1, int t2, int t3, int t4)
{
    static constexpr astc::TritPack pack = astc::make_trit_pack();
    const int t[5] = {t0, t1, t2, t3, t4};
    for (int v : t) if (v < 0 || v > 2) return ASTC_B200_ERR_INVALID_ARGUMENT;
    return pack.v[t4 * 81 + t3 * 27 + t2 * 9 + t1 * 3 + t0];
}

int astc_b200_integer_from_quints(int q0, int q1, int q2)
{
    static constexpr astc::QuintPack pack = astc::make_quint_pack();
    const int q[3] = {q0, q1, q2};
    for (int v : q) if (v < 0 || v > 4) return ASTC_B200_ERR_INVALID_ARGUMENT;
    return pack.v[q2 * 25 + q1 * 5 + q0];
}

int astc_b200_scramble(int method, int q)
{
    static constexpr astc::WeightTables tables = astc::make_weight_tables();
    if (method < 0 || method >= astc::kWeightMethods || q < 0 || q >= astc::kScrambleStride)
        return ASTC_B200_ERR_INVALID_ARGUMENT;
    return tables.scramble[method][q];
}

uint32_t astc_b200_blockmode(int weight_quant) { return astc::blockmode_4x4grid(weight_quant); }

int astc_b200_unorm_lut(int srgb, float out[256])
{
    if (!out) return ASTC_B200_ERR_INVALID_ARGUMENT;
    for (int c = 0; c < 256; ++c) out[c] = srgb ? astc::host_srgb_lut()[c] : astc::host_unorm_lut()[c];
    return ASTC_B200_OK;
}

int astc_b200_decode_device(const uint8_t *d_blocks, int width, int height, int block_dim, uint8_t *d_rgba,
                            size_t pitch_bytes, void *cuda_stream)
{
    if (width < 0 || height < 0 || (block_dim != 4 && block_dim != 6)) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (width == 0 || height == 0) return ASTC_B200_OK;
    if (!d_blocks || !d_rgba || pitch_bytes < size_t(width) * 4u || pitch_bytes % 4u != 0 ||
        reinterpret_cast<uintptr_t>(d_rgba) % 4u != 0 || reinterpret_cast<uintptr_t>(d_blocks) % 16u != 0)
        return ASTC_B200_ERR_INVALID_ARGUMENT;
    CUDA_TRY(astc::launch_decode(d_blocks, width, height, block_dim, d_rgba, pitch_bytes,
                                 static_cast<cudaStream_t>(cuda_stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return ASTC_B200_OK;
}

int astc_b200_downsample2x2_device(const uint8_t *d_src, int width, int height, size_t src_pitch_bytes, uint8_t *d_dst,
                                   size_t dst_pitch_bytes, void *cuda_stream)
{
    if (width < 0 || height < 0) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (width == 0 || height == 0) return ASTC_B200_OK;
    const size_t ow = width > 1 ? size_t(width) / 2u : 1u;
    if (!d_src || !d_dst || src_pitch_bytes < size_t(width) * 4u || dst_pitch_bytes < ow * 4u || src_pitch_bytes % 4u != 0 ||
        dst_pitch_bytes % 4u != 0 || reinterpret_cast<uintptr_t>(d_src) % 4u != 0 || reinterpret_cast<uintptr_t>(d_dst) % 4u != 0)
        return ASTC_B200_ERR_INVALID_ARGUMENT;
    CUDA_TRY(astc::launch_downsample2x2(d_src, width, height, src_pitch_bytes, d_dst, dst_pitch_bytes,
                                        static_cast<cudaStream_t>(cuda_stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return ASTC_B200_OK;
}

int astc_b200_mip_chain_layout(int width, int height, int *levels, size_t *offsets, int *widths, int *heights, size_t *total_bytes)
{
    if (width <= 0 || height <= 0 || !levels) return ASTC_B200_ERR_INVALID_ARGUMENT;
    *levels = astc_capi::mip_layout(width, height, offsets, widths, heights, total_bytes);
    return ASTC_B200_OK;
}

int astc_b200_mip_chain_device(const uint8_t *d_base, int width, int height, size_t pitch_bytes, uint8_t *d_levels,
                               size_t levels_bytes, void *cuda_stream)
{
    if (width < 0 || height < 0) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (width == 0 || height == 0) return ASTC_B200_OK;
    size_t offsets[astc::kMaxMipLevels], total = 0;
    int widths[astc::kMaxMipLevels], heights[astc::kMaxMipLevels];
    const int n = astc_capi::mip_layout(width, height, offsets, widths, heights, &total);
    if (n == 0) return ASTC_B200_OK;                                  // a 1x1 base has no further levels
    if (!d_base || !d_levels || pitch_bytes < size_t(width) * 4u || pitch_bytes % 4u != 0 || levels_bytes < total ||
        reinterpret_cast<uintptr_t>(d_base) % 4u != 0 || reinterpret_cast<uintptr_t>(d_levels) % 256u != 0)
        return ASTC_B200_ERR_INVALID_ARGUMENT;
    uint8_t *ptrs[astc::kMaxMipLevels];
    for (int l = 0; l < n; ++l) ptrs[l] = d_levels + offsets[l];
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    unsigned *ticket = reinterpret_cast<unsigned *>(d_levels + total - 256u);
    const bool fused = width % 64 == 0 && height % 64 == 0;
    if (fused) CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));    // the arena is the caller's: never assume it is zeroed
    CUDA_TRY(astc::launch_mip_chain(d_base, width, height, pitch_bytes, ptrs, widths, heights, n, ticket, st));
    g_launches.fetch_add(fused && pitch_bytes % 16u == 0 && reinterpret_cast<uintptr_t>(d_base) % 16u == 0 ? 1 : uint64_t(n),
                         std::memory_order_relaxed);
    return ASTC_B200_OK;
}

int astc_b200_mufu_device(int op, const float *d_x, float *d_y, size_t count, void *cuda_stream)
{
    if (op != 0 && op != 1) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (count == 0) return ASTC_B200_OK;
    if (!d_x || !d_y) return ASTC_B200_ERR_INVALID_ARGUMENT;
    CUDA_TRY(astc::launch_mufu(op, d_x, d_y, count, static_cast<cudaStream_t>(cuda_stream)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return ASTC_B200_OK;
}

int astc_b200_malloc_device(void **d_ptr, size_t bytes)
{
    if (!d_ptr) return ASTC_B200_ERR_INVALID_ARGUMENT;
    *d_ptr = nullptr;
    if (bytes == 0) return ASTC_B200_OK;
    CUDA_TRY(cudaMalloc(d_ptr, bytes));
    return ASTC_B200_OK;
}

int astc_b200_free_device(void *d_ptr)
{
    if (d_ptr) CUDA_TRY(cudaFree(d_ptr));
    return ASTC_B200_OK;
}

int astc_b200_host_alloc(void **h_ptr, size_t bytes)
{
    if (!h_ptr) return ASTC_B200_ERR_INVALID_ARGUMENT;
    *h_ptr = nullptr;
    if (bytes == 0) return ASTC_B200_OK;
    CUDA_TRY(cudaHostAlloc(h_ptr, bytes, cudaHostAllocDefault));
    return ASTC_B200_OK;
}

int astc_b200_host_free(void *h_ptr)
{
    if (h_ptr) CUDA_TRY(cudaFreeHost(h_ptr));
    return ASTC_B200_OK;
}

int astc_b200_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *cuda_stream)
{
    if (bytes == 0) return ASTC_B200_OK;
    if (!d_dst || !h_src) return ASTC_B200_ERR_INVALID_ARGUMENT;
    CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(cuda_stream)));
    return ASTC_B200_OK;
}

int astc_b200_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *cuda_stream)
{
    if (bytes == 0) return ASTC_B200_OK;
    if (!h_dst || !d_src) return ASTC_B200_ERR_INVALID_ARGUMENT;
    CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(cuda_stream)));
    return ASTC_B200_OK;
}

int astc_b200_memcpy2d_h2d(void *d_dst, size_t d_pitch, const void *h_src, size_t h_pitch, size_t row_bytes, size_t rows,
                           void *cuda_stream)
{
    if (row_bytes == 0 || rows == 0) return ASTC_B200_OK;
    if (!d_dst || !h_src) return ASTC_B200_ERR_INVALID_ARGUMENT;
    CUDA_TRY(cudaMemcpy2DAsync(d_dst, d_pitch, h_src, h_pitch, row_bytes, rows, cudaMemcpyHostToDevice,
                               static_cast<cudaStream_t>(cuda_stream)));
    return ASTC_B200_OK;
}

int astc_b200_stream_create(void **cuda_stream)
{
    if (!cuda_stream) return ASTC_B200_ERR_INVALID_ARGUMENT;
    cudaStream_t s = nullptr;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *cuda_stream = s;
    return ASTC_B200_OK;
}

int astc_b200_stream_destroy(void *cuda_stream)
{
    if (cuda_stream) CUDA_TRY(cudaStreamDestroy(static_cast<cudaStream_t>(cuda_stream)));
    return ASTC_B200_OK;
}

int astc_b200_stream_synchronize(void *cuda_stream)
{
    CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    return ASTC_B200_OK;
}

int astc_b200_save_astc(const char *path, int xdim, int ydim, int xsize, int ysize, const uint8_t *blocks, size_t bufsz)
{
    if (!path || (!blocks && bufsz)) return ASTC_B200_ERR_INVALID_ARGUMENT;
    return astc_save::write_file(path, xdim, ydim, xsize, ysize, blocks, bufsz) ? ASTC_B200_OK : ASTC_B200_ERR_IO;
}

int astc_b200_save_astc_slice(const char *path, int xdim, int ydim, int xsize, int ysize, size_t block_byte_offset,
                              const uint8_t *blocks, size_t nbytes, int write_header)
{
    if (!path || (!blocks && nbytes) || xdim <= 0 || ydim <= 0 || xsize < 0 || ysize < 0) return ASTC_B200_ERR_INVALID_ARGUMENT;
    const size_t payload = size_t((xsize + xdim - 1) / xdim) * size_t((ysize + ydim - 1) / ydim) * ASTC_B200_BLOCK_BYTES;
    if (block_byte_offset > payload || nbytes > payload - block_byte_offset) return ASTC_B200_ERR_INVALID_ARGUMENT;
    const int fd = ::open(path, O_WRONLY | O_CREAT, 0644);          // no O_TRUNC: other ranks may have written already
    if (fd < 0) return ASTC_B200_ERR_IO;
    bool ok = ::ftruncate(fd, off_t(sizeof(astc_header) + payload)) == 0;   // idempotent: every rank sets the same size
    auto put = [&](const void *src, size_t n, off_t at) {
        const uint8_t *p = static_cast<const uint8_t *>(src);
        while (ok && n) {
            const ssize_t w = ::pwrite(fd, p, n, at);
            if (w <= 0) { ok = false; break; }
            p += w; n -= size_t(w); at += w;
        }
    };
    if (write_header) {
        const astc_header h = astc_save::make_header(xdim, ydim, xsize, ysize);
        put(&h, sizeof h, 0);
    }
    put(blocks, nbytes, off_t(sizeof(astc_header) + block_byte_offset));
    ok = (::close(fd) == 0) && ok;
    return ok ? ASTC_B200_OK : ASTC_B200_ERR_IO;
}

int astc_b200_load_astc(const char *path, int *xdim, int *ydim, int *xsize, int *ysize, uint8_t **blocks, size_t *bufsz)
{
    if (!path || !blocks || !bufsz) return ASTC_B200_ERR_INVALID_ARGUMENT;
    *blocks = nullptr; *bufsz = 0;
    std::FILE *f = std::fopen(path, "rb");
    if (!f) return ASTC_B200_ERR_IO;
    astc_header hdr{};
    int rc = ASTC_B200_OK;
    if (std::fread(&hdr, 1, sizeof hdr, f) != sizeof hdr) rc = ASTC_B200_ERR_BAD_IMAGE;
    if (rc == ASTC_B200_OK) {
        const uint32_t magic = uint32_t(hdr.magic[0]) | (uint32_t(hdr.magic[1]) << 8) | (uint32_t(hdr.magic[2]) << 16) |
                               (uint32_t(hdr.magic[3]) << 24);
        const int zs = hdr.zsize[0] | (hdr.zsize[1] << 8) | (hdr.zsize[2] << 16);
        // this encoder's files are 2-D (astc_save.h:60-66 writes blockdim_z = zsize = 1); a 3-D file is refused, not flattened
        if (magic != ASTC_B200_MAGIC || hdr.blockdim_x == 0 || hdr.blockdim_y == 0 || hdr.blockdim_z != 1 || zs != 1)
            rc = ASTC_B200_ERR_BAD_IMAGE;
    }
    if (rc == ASTC_B200_OK) {
        const int xs = hdr.xsize[0] | (hdr.xsize[1] << 8) | (hdr.xsize[2] << 16);
        const int ys = hdr.ysize[0] | (hdr.ysize[1] << 8) | (hdr.ysize[2] << 16);
        const size_t nb = size_t((xs + hdr.blockdim_x - 1) / hdr.blockdim_x) * size_t((ys + hdr.blockdim_y - 1) / hdr.blockdim_y);
        // the payload must be there before anything is allocated for it (24-bit sizes allow 2^44 blocks)
        long have = -1;
        if (std::fseek(f, 0, SEEK_END) == 0) have = std::ftell(f);
        if (have < 0 || size_t(have) < sizeof hdr || (size_t(have) - sizeof hdr) / 16u < nb || std::fseek(f, long(sizeof hdr), SEEK_SET) != 0)
            rc = ASTC_B200_ERR_BAD_IMAGE;
        uint8_t *buf = rc == ASTC_B200_OK ? static_cast<uint8_t *>(std::malloc(nb * 16u + 1u)) : nullptr;
        if (rc != ASTC_B200_OK) {}
        else if (!buf) rc = ASTC_B200_ERR_OUT_OF_MEMORY;
        else if (std::fread(buf, 1, nb * 16u, f) != nb * 16u) { std::free(buf); rc = ASTC_B200_ERR_BAD_IMAGE; }
        else {
            *blocks = buf; *bufsz = nb * 16u;
            if (xdim) *xdim = hdr.blockdim_x;
            if (ydim) *ydim = hdr.blockdim_y;
            if (xsize) *xsize = xs;
            if (ysize) *ysize = ys;
        }
    }
    std::fclose(f);
    return rc;
}

void astc_b200_free_host_buffer(void *p) { std::free(p); }

}  // extern "C"
