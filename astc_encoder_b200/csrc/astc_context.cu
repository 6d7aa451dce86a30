// astc_context.cu -- the host-buffer entry points of the C ABI: a persistent per-device context
// (streams, event, grow-only device workspace, pinned staging) and the two calls built on it,
//
//   astc_b200_context_encode_host        load_tex's upload + encode_astc + read_gpu (main.cpp:46-52,
//                                        astc_encode.h:87-194, astc_save.h:34-50) for ONE texture
//   astc_b200_context_batch_encode_host  the same for MANY textures (mip chains): pinned staging for the
//                                        small levels, banded uploads for the large ones, one kernel
//                                        launch per group of ~32 MiB over a prefix-summed block table
//
// The reference creates its device objects once per process (main.cpp:199-209) and its texture / UAV
// per encode; here the per-call objects are gone too: nothing is created or destroyed on the hot path
// once the workspace has grown to the largest job seen.  astc_b200_encode_host() (no context argument)
// runs on a lazily created thread-local context per device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <vector>

#include "astc_capi_internal.h"
#include "host_copy_pool.h"

using astc_capi::cuda_fail;
using astc_capi::dim_of;
using astc_capi::make_desc;

namespace {

constexpr int kStreams = 3;
constexpr size_t kSmallImage = 256u << 10;        // sources below this travel through pinned staging (one memcpy beats one cudaMemcpyAsync call)
#ifndef ASTC_GROUP_MIB
#define ASTC_GROUP_MIB 32
#endif
constexpr size_t kGroupBytes = size_t(ASTC_GROUP_MIB) << 20;   // source bytes per upload / launch / download group of a batch
constexpr size_t kStagedMin = 1u << 20;           // pageable textures from this size on go through the staged pipeline below
constexpr size_t kStagedBand = 4u << 20;          // source bytes per staged band (one pinned slot per stream)

template <typename T>
struct Grow {                                      // grow-only buffer: reallocated only when a job is larger than any before
    T *ptr = nullptr;
    size_t cap = 0;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Host -> device copy of `rows` rows.  A 2-D copy moves row by row -- measured on B200 / PCIe Gen5: 13 GB/s for
// 4 KB rows against 55 GB/s for one flat copy -- so rows that are contiguous on both sides go as ONE 1-D copy.
cudaError_t upload_rows(uint8_t *dst, size_t dst_pitch, const uint8_t *src, size_t src_pitch, size_t row_bytes, size_t rows,
                        cudaStream_t st)
{
    if (rows == 0 || row_bytes == 0) return cudaSuccess;
    if (rows == 1 || (dst_pitch == row_bytes && src_pitch == row_bytes))
        return cudaMemcpyAsync(dst, src, row_bytes * rows, cudaMemcpyHostToDevice, st);
    return cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, rows, cudaMemcpyHostToDevice, st);
}

// Pageable host memory (what the reference's caller has: stbi_load's malloc, `new uint8_t[]`, main.cpp:24,224).
// cudaMemcpyAsync from such a buffer is staged by the driver on the calling thread at ~11 GB/s (measured on the
// B200 box, against 55 GB/s from pinned memory).  The staged pipeline of astc_b200_context_encode_host does that
// staging itself, into pinned slots, with a few worker threads: the host memcpy then runs at the rate of several
// cores and overlaps the DMA and the kernel of the bands before it.
bool is_pageable(const void *p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace

struct astc_b200_context {
    int device = 0;
    cudaStream_t streams[kStreams] = {};
    cudaEvent_t ready = nullptr;
    cudaEvent_t slot_done[kStreams] = {};                   // staged pipeline: the band that used slot s has left it
    Grow<uint8_t> d_in, d_out, h_stage_in, h_stage_out;     // device workspace, pinned staging
    Grow<astc::ImageDesc> d_table;
    astc_host::CopyPool pool;

    ~astc_b200_context()
    {
        for (auto &s : streams) if (s) cudaStreamDestroy(s);
        if (ready) cudaEventDestroy(ready);
        for (auto &e : slot_done) if (e) cudaEventDestroy(e);
        if (d_in.ptr) cudaFree(d_in.ptr);
        if (d_out.ptr) cudaFree(d_out.ptr);
        if (d_table.ptr) cudaFree(d_table.ptr);
        if (h_stage_in.ptr) cudaFreeHost(h_stage_in.ptr);
        if (h_stage_out.ptr) cudaFreeHost(h_stage_out.ptr);
    }
};

namespace {

template <typename T>
cudaError_t reserve_device(Grow<T> &g, size_t count)
{
    if (count <= g.cap) return cudaSuccess;
    if (g.ptr) { cudaFree(g.ptr); g.ptr = nullptr; g.cap = 0; }
    const size_t want = count + count / 8;                     // a little head room: a slightly larger next job does not reallocate
    cudaError_t e = cudaMalloc((void **)&g.ptr, want * sizeof(T));
    if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc((void **)&g.ptr, count * sizeof(T)); if (e == cudaSuccess) g.cap = count; return e; }
    g.cap = want;
    return cudaSuccess;
}

cudaError_t reserve_pinned(Grow<uint8_t> &g, size_t bytes)
{
    if (bytes <= g.cap) return cudaSuccess;
    if (g.ptr) { cudaFreeHost(g.ptr); g.ptr = nullptr; g.cap = 0; }
    const cudaError_t e = cudaHostAlloc((void **)&g.ptr, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) g.cap = bytes;
    return e;
}

int create_context(astc_b200_context **out)
{
    *out = nullptr;
    std::unique_ptr<astc_b200_context> c(new (std::nothrow) astc_b200_context());
    if (!c) return ASTC_B200_ERR_OUT_OF_MEMORY;
    CUDA_TRY(cudaGetDevice(&c->device));
    for (auto &s : c->streams) CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming));
    for (auto &e : c->slot_done) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *out = c.release();
    return ASTC_B200_OK;
}

// One lazily created context per (host thread, device) behind astc_b200_encode_host().
astc_b200_context *default_context(int *status)
{
    struct Holder {
        std::vector<astc_b200_context *> per_device;
        ~Holder() { for (auto *c : per_device) delete c; }
    };
    static thread_local Holder holder;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { *status = cuda_fail(e, "cudaGetDevice"); return nullptr; }
    if (size_t(dev) >= holder.per_device.size()) holder.per_device.resize(size_t(dev) + 1, nullptr);
    if (!holder.per_device[dev]) {
        *status = create_context(&holder.per_device[dev]);
        if (*status != ASTC_B200_OK) return nullptr;
    }
    *status = ASTC_B200_OK;
    return holder.per_device[dev];
}

int check_context(const astc_b200_context *ctx)
{
    if (!ctx) return ASTC_B200_ERR_INVALID_ARGUMENT;
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    return dev == ctx->device ? ASTC_B200_OK : ASTC_B200_ERR_INVALID_ARGUMENT;     // a context belongs to the device it was created on
}

bool too_many_blocks(int w, int h, int dim)
{
    return uint64_t((w + dim - 1) / dim) * uint64_t((h + dim - 1) / dim) > 0x7FFFFFFFull;
}

cudaError_t sync_all(astc_b200_context *ctx, cudaError_t err)
{
    for (auto &s : ctx->streams) {
        const cudaError_t e2 = cudaStreamSynchronize(s);
        if (err == cudaSuccess) err = e2;
    }
    return err;
}

}  // namespace

extern "C" {

int astc_b200_context_create(astc_b200_context **ctx)
{
    if (!ctx) return ASTC_B200_ERR_INVALID_ARGUMENT;
    return create_context(ctx);
}

void astc_b200_context_destroy(astc_b200_context *ctx) { delete ctx; }

int astc_b200_context_set_copy_threads(astc_b200_context *ctx, int threads)
{
    if (!ctx || threads < -1 || threads > 64) return ASTC_B200_ERR_INVALID_ARGUMENT;
    ctx->pool.set_workers(threads);
    return ASTC_B200_OK;
}

int astc_b200_context_trim(astc_b200_context *ctx)
{
    const int rc = check_context(ctx);
    if (rc != ASTC_B200_OK) return rc;
    CUDA_TRY(sync_all(ctx, cudaSuccess));
    auto drop = [](auto &g) { if (g.ptr) cudaFree(g.ptr); g.ptr = nullptr; g.cap = 0; };
    drop(ctx->d_in); drop(ctx->d_out); drop(ctx->d_table);
    auto droph = [](Grow<uint8_t> &g) { if (g.ptr) cudaFreeHost(g.ptr); g.ptr = nullptr; g.cap = 0; };
    droph(ctx->h_stage_in); droph(ctx->h_stage_out);
    return ASTC_B200_OK;
}

int astc_b200_context_encode_host(astc_b200_context *ctx, const uint8_t *h_rgba, int width, int height, size_t pitch_bytes,
                                  const astc_b200_option *opt, uint8_t *h_blocks)
{
    if (!opt || width < 0 || height < 0 || opt->axis_method > 1) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (width == 0 || height == 0) return ASTC_B200_OK;
    if (!h_rgba || !h_blocks || pitch_bytes < size_t(width) * 4u) return ASTC_B200_ERR_INVALID_ARGUMENT;
    int rc = check_context(ctx);
    if (rc != ASTC_B200_OK) return rc;
    const int d = dim_of(opt);
    if (too_many_blocks(width, height, d)) return ASTC_B200_ERR_UNSUPPORTED;

    const int64_t bx = (width + d - 1) / d, by = (height + d - 1) / d;
    // device pitch = the row itself when that keeps rows 16-byte aligned (the kernels' vector path), so a contiguous
    // host image uploads as flat 1-D copies; padded to 128 B otherwise
    const size_t row_bytes = size_t(width) * 4u;
    const size_t d_pitch = row_bytes % 16u == 0 ? row_bytes : align_up(row_bytes, 128);
    CUDA_TRY(reserve_device(ctx->d_in, d_pitch * size_t(height)));
    CUDA_TRY(reserve_device(ctx->d_out, size_t(bx * by) * 16u));
    uint8_t *d_in = ctx->d_in.ptr, *d_out = ctx->d_out.ptr;

    // Pageable source and / or destination of at least 1 MiB: the staged pipeline.  Band b uses stream and pinned
    // slot b % 3: [workers copy its rows into the slot] -> flat H2D -> kernel -> D2H [into the slot] -> event;
    // before slot s is refilled, the band that used it three bands ago is retired (event wait, then its blocks
    // are copied out to the caller's buffer).  The host copies of band b overlap the DMA and kernels of b-1, b-2.
    const bool big = d_pitch * size_t(height) >= kStagedMin;            // (the pointer queries cost ~1 us each: not on small textures)
    const bool stage_in = big && is_pageable(h_rgba), stage_out = big && is_pageable(h_blocks);
    if (stage_in || stage_out) {
        // bands of a quarter of the texture, between 512 KiB and 4 MiB: even a 1 MiB texture overlaps its copies
        const size_t band_bytes = std::min(kStagedBand, std::max<size_t>(512u << 10, d_pitch * size_t(height) / 4u));
        const int64_t rows_pb = std::max<int64_t>(1, int64_t(band_bytes / (d_pitch * size_t(d))));      // block rows per band
        const int nb = int((by + rows_pb - 1) / rows_pb);
        const size_t in_slot = size_t(rows_pb) * size_t(d) * d_pitch, out_slot = size_t(rows_pb * bx) * 16u;
        if (stage_in) CUDA_TRY(reserve_pinned(ctx->h_stage_in, in_slot * kStreams));
        if (stage_out) CUDA_TRY(reserve_pinned(ctx->h_stage_out, out_slot * kStreams));
        auto out_range = [&](int b, size_t &off, size_t &bytes) {
            const int64_t r0 = int64_t(b) * rows_pb, r1 = std::min<int64_t>(by, r0 + rows_pb);
            off = size_t(r0 * bx) * 16u;
            bytes = size_t((r1 - r0) * bx) * 16u;
        };
        auto retire = [&](int b) -> cudaError_t {                       // band b is complete: hand its blocks over
            const int s = b % kStreams;
            const cudaError_t e = cudaEventSynchronize(ctx->slot_done[s]);
            if (e != cudaSuccess || !stage_out) return e;
            size_t off, bytes;
            out_range(b, off, bytes);
            ctx->pool.copy_rows(h_blocks + off, bytes, ctx->h_stage_out.ptr + size_t(s) * out_slot, bytes, bytes, 1);
            return cudaSuccess;
        };
        cudaError_t err = cudaSuccess;
        int retired = 0;
        for (int b = 0; b < nb && err == cudaSuccess; ++b) {
            const int s = b % kStreams;
            cudaStream_t st = ctx->streams[s];
            if (b >= kStreams) { err = retire(b - kStreams); ++retired; if (err != cudaSuccess) break; }
            const int64_t r0 = int64_t(b) * rows_pb, r1 = std::min<int64_t>(by, r0 + rows_pb);
            const int64_t y0 = r0 * d, y1 = std::min<int64_t>(int64_t(height), r1 * d);
            uint8_t *d_band = d_in + size_t(y0) * d_pitch;
            if (stage_in) {
                uint8_t *slot = ctx->h_stage_in.ptr + size_t(s) * in_slot;
                ctx->pool.copy_rows(slot, d_pitch, h_rgba + size_t(y0) * pitch_bytes, pitch_bytes, row_bytes, size_t(y1 - y0),
                                    /*streaming=*/true);            // only the DMA engine reads the slot
                // the padding between row_bytes and d_pitch (rows that are not a multiple of 16 bytes) is never read
                err = cudaMemcpyAsync(d_band, slot, size_t(y1 - y0 - 1) * d_pitch + row_bytes, cudaMemcpyHostToDevice, st);
            } else {
                err = upload_rows(d_band, d_pitch, h_rgba + size_t(y0) * pitch_bytes, pitch_bytes, row_bytes, size_t(y1 - y0), st);
            }
            if (err != cudaSuccess) break;
            size_t out_off, out_bytes;
            out_range(b, out_off, out_bytes);
            astc::EncodeParams p{};
            p.single = make_desc(d_band, d_out + out_off, d_pitch, width, int(y1 - y0), d, 0);
            p.count = 1;
            p.total_blocks = uint64_t(bx) * uint64_t(r1 - r0);
            err = astc::launch_encode(d, opt->has_alpha != 0, opt->is_normal_map != 0, opt->srgb != 0, opt->axis_method, p, st);
            if (err != cudaSuccess) break;
            astc_capi::count_launch();
            err = cudaMemcpyAsync(stage_out ? ctx->h_stage_out.ptr + size_t(s) * out_slot : h_blocks + out_off, d_out + out_off, out_bytes,
                                  cudaMemcpyDeviceToHost, st);
            if (err == cudaSuccess) err = cudaEventRecord(ctx->slot_done[s], st);
        }
        for (int b = retired; b < nb && err == cudaSuccess; ++b) err = retire(b);
        if (err != cudaSuccess) {
            sync_all(ctx, err);
            return cuda_fail(err, "astc_b200_context_encode_host (staged)");
        }
        return ASTC_B200_OK;
    }

    // Bands keep the three engines (H2D, SM, D2H) busy at once.  The H2D copies are the bottleneck and
    // run back to back; what the banding costs on top is the drain after the last copy (that band's
    // kernel and its D2H: ~5.2 ps per byte of band) plus ~7.8 us of launch / copy overhead per band
    // (both measured on B200 / PCIe Gen5: 1 GiB in 1 / 8 / 32 / 128 MiB bands = 28.4 / 20.8 / 20.45 /
    // 20.9 ms).  The sum is smallest at sqrt(5.2e-12 / 7.8e-6 * bytes) bands: 27 for 1 GiB, 7 for 64 MiB,
    // one below 1.5 MiB.  Experiment builds (-DASTC_TUNING_HOOKS) let ASTC_B200_HOST_BAND_MIB override it.
    const double src_bytes = double(d_pitch) * double(height);
    int64_t want_bands = std::max<int64_t>(1, int64_t(std::sqrt(6.7e-7 * src_bytes) + 0.5));
#ifdef ASTC_TUNING_HOOKS
    if (const char *env = getenv("ASTC_B200_HOST_BAND_MIB")) {
        const long v = atol(env);
        if (v >= 1 && v <= 1024) want_bands = std::max<int64_t>(1, int64_t(src_bytes / (double(v) * 1048576.0) + 0.5));
    }
#endif
    const int64_t rows_per_band = std::max<int64_t>(1, (by + want_bands - 1) / want_bands);
    const int nbands = int((by + rows_per_band - 1) / rows_per_band);
    cudaError_t err = cudaSuccess;
    for (int b = 0; b < nbands && err == cudaSuccess; ++b) {
        cudaStream_t st = ctx->streams[b % kStreams];
        const int64_t r0 = int64_t(b) * rows_per_band, r1 = std::min<int64_t>(by, r0 + rows_per_band);
        const int64_t y0 = r0 * d, y1 = std::min<int64_t>(int64_t(height), r1 * d);
        const size_t out_off = size_t(r0 * bx) * 16u, out_bytes = size_t((r1 - r0) * bx) * 16u;
        err = upload_rows(d_in + size_t(y0) * d_pitch, d_pitch, h_rgba + size_t(y0) * pitch_bytes, pitch_bytes, row_bytes,
                          size_t(y1 - y0), st);
        if (err != cudaSuccess) break;
        astc::EncodeParams p{};
        p.single = make_desc(d_in + size_t(y0) * d_pitch, d_out + out_off, d_pitch, width, int(y1 - y0), d, 0);
        p.count = 1;
        p.total_blocks = uint64_t(bx) * uint64_t(r1 - r0);
        err = astc::launch_encode(d, opt->has_alpha != 0, opt->is_normal_map != 0, opt->srgb != 0, opt->axis_method, p, st);
        if (err != cudaSuccess) break;
        astc_capi::count_launch();
        err = cudaMemcpyAsync(h_blocks + out_off, d_out + out_off, out_bytes, cudaMemcpyDeviceToHost, st);
    }
    // only the streams that were used need to drain
    for (int s = 0; s < std::min(nbands, kStreams); ++s) {
        const cudaError_t e2 = cudaStreamSynchronize(ctx->streams[s]);
        if (err == cudaSuccess) err = e2;
    }
    if (err != cudaSuccess) return cuda_fail(err, "astc_b200_context_encode_host");
    return ASTC_B200_OK;
}

int astc_b200_encode_host(const uint8_t *h_rgba, int width, int height, size_t pitch_bytes, const astc_b200_option *opt,
                          uint8_t *h_blocks)
{
    if (!opt || width < 0 || height < 0 || opt->axis_method > 1) return ASTC_B200_ERR_INVALID_ARGUMENT;
    if (width == 0 || height == 0) return ASTC_B200_OK;
    if (!h_rgba || !h_blocks || pitch_bytes < size_t(width) * 4u) return ASTC_B200_ERR_INVALID_ARGUMENT;
    int rc = ASTC_B200_OK;
    astc_b200_context *ctx = default_context(&rc);
    if (!ctx) return rc;
    return astc_b200_context_encode_host(ctx, h_rgba, width, height, pitch_bytes, opt, h_blocks);
}

}  // extern "C" (the shared implementation below is C++)

namespace {

// The body of both batch entry points.  `mips`: every image is the BASE of a mip chain -- only it is uploaded, the
// levels below it are produced on the device (astc::launch_mip_chain: one fused launch per chain when both sides are
// multiples of 64) and encoded by the same batch launch; im.h_blocks receives the blocks of all levels, base first.
int batch_encode_host_impl(astc_b200_context *ctx, const astc_b200_host_image *images, int count, const astc_b200_option *opt, bool mips)
{
    if (!opt || count < 0 || (count > 0 && !images) || opt->axis_method > 1) return ASTC_B200_ERR_INVALID_ARGUMENT;
    int rc = check_context(ctx);
    if (rc != ASTC_B200_OK) return rc;
    const int d = dim_of(opt);

    // ---- plan: arena offsets, block prefix sums per GROUP, which images travel through the pinned slots ----
    // An image goes through a slot when it is small (one host memcpy beats one cudaMemcpyAsync call) or when it lies
    // in PAGEABLE memory (the copy workers beat the driver's own staging 2-3x); large pinned images are copied
    // straight from / to the caller's buffer.  A group's slot mirrors the group's stretch of the device arena, so
    // every run of consecutive slot images is uploaded -- and every run of slot outputs downloaded -- with ONE copy.
    struct Item {
        const uint8_t *h_src;            // nullptr: a mip level produced on the device
        uint8_t *h_dst;
        size_t src_pitch;
        int w, h;
        size_t in_off, in_pitch, in_bytes, out_off, out_bytes;
        bool via_in, via_out, chain_head;
        int levels;                      // chain_head: derived levels that follow
        size_t mip_off;                  // chain_head: device offset of the chain's level arena (astc_capi::mip_layout)
        uint64_t blocks;
    };
    struct Group {
        size_t first, count;             // items
        uint64_t blocks;
        size_t in_off, in_span, out_off, out_bytes;
    };
    std::vector<Item> items;
    std::vector<Group> groups;
    std::vector<astc::ImageDesc> table;
    size_t in_total = 0, out_total = 0, slot_in = 0, slot_out = 0;
    try {
        items.reserve(size_t(count) * (mips ? 13u : 1u));
        for (int i = 0; i < count; ++i) {
            const astc_b200_host_image &im = images[i];
            if (im.width < 0 || im.height < 0) return ASTC_B200_ERR_INVALID_ARGUMENT;
            if (im.width == 0 || im.height == 0) continue;
            if (!im.h_rgba || !im.h_blocks || im.pitch_bytes < size_t(im.width) * 4u) return ASTC_B200_ERR_INVALID_ARGUMENT;
            if (too_many_blocks(im.width, im.height, d)) return ASTC_B200_ERR_UNSUPPORTED;
            Item it{};
            it.h_src = im.h_rgba;
            it.h_dst = im.h_blocks;
            it.src_pitch = im.pitch_bytes;
            it.w = im.width;
            it.h = im.height;
            it.in_pitch = align_up(size_t(im.width) * 4u, 16);
            it.in_bytes = it.in_pitch * size_t(im.height);
            it.in_off = in_total;
            in_total += align_up(it.in_bytes, 256);
            it.blocks = astc_capi::block_count(im.width, im.height, d);
            it.out_bytes = size_t(it.blocks) * 16u;
            it.out_off = out_total;
            out_total += it.out_bytes;
            // (images above a group's size never go through a slot: the slots stay bounded; the pointer query costs
            // about a microsecond, so it is skipped for images that take the slot anyway)
            it.via_in = it.in_bytes < kSmallImage || (it.in_bytes <= kGroupBytes && is_pageable(im.h_rgba));
            const bool pageable_out = it.in_bytes <= kGroupBytes && (mips || it.out_bytes >= kSmallImage / 4) && is_pageable(im.h_blocks);
            it.via_out = it.out_bytes < kSmallImage / 4 || pageable_out;
            it.chain_head = mips;
            if (mips) {
                size_t offs[astc::kMaxMipLevels], total = 0;
                int ws[astc::kMaxMipLevels], hs[astc::kMaxMipLevels];
                it.levels = astc_capi::mip_layout(im.width, im.height, offs, ws, hs, &total);
                it.mip_off = in_total;
                in_total += align_up(total, 256);
                items.push_back(it);
                uint8_t *dst = im.h_blocks + it.out_bytes;
                for (int l = 0; l < it.levels; ++l) {
                    Item lv{};
                    lv.h_dst = dst;
                    lv.w = ws[l];
                    lv.h = hs[l];
                    lv.in_pitch = size_t(ws[l]) * 4u;                  // the mip kernels pack rows tightly
                    lv.in_bytes = lv.in_pitch * size_t(hs[l]);
                    lv.in_off = it.mip_off + offs[l];
                    lv.blocks = astc_capi::block_count(ws[l], hs[l], d);
                    lv.out_bytes = size_t(lv.blocks) * 16u;
                    lv.out_off = out_total;
                    out_total += lv.out_bytes;
                    lv.via_out = lv.out_bytes < kSmallImage / 4 || pageable_out;
                    dst += lv.out_bytes;
                    items.push_back(lv);
                }
            } else {
                items.push_back(it);
            }
        }
        if (items.empty()) return ASTC_B200_OK;
        // groups of consecutive images (whole chains), ~kGroupBytes of uploaded source each; block ids restart at 0 in every group
        table.resize(items.size());
        for (size_t i = 0; i < items.size();) {
            Group g{};
            g.first = i;
            g.in_off = items[i].in_off;
            g.out_off = items[i].out_off;
            size_t bytes = 0;
            while (i < items.size() && (bytes == 0 || items[i].h_src == nullptr || bytes + items[i].in_bytes <= kGroupBytes)) {
                if (items[i].h_src) bytes += items[i].in_bytes;
                ++i;
            }
            g.count = i - g.first;
            bool any_in = false, any_out = false;
            for (size_t k = g.first; k < i; ++k) {
                const Item &it = items[k];
                table[k] = make_desc(nullptr, nullptr, it.in_pitch, it.w, it.h, d, g.blocks);   // pointers filled in below
                g.blocks += it.blocks;
                g.out_bytes += it.out_bytes;
                any_in |= it.via_in;
                any_out |= it.via_out;
            }
            g.in_span = items[i - 1].in_off + items[i - 1].in_bytes - g.in_off;
            if (any_in) slot_in = std::max(slot_in, align_up(g.in_span, 256));
            if (any_out) slot_out = std::max(slot_out, align_up(g.out_bytes, 256));
            groups.push_back(g);
        }
    } catch (const std::bad_alloc &) {
        return ASTC_B200_ERR_OUT_OF_MEMORY;
    }

    CUDA_TRY(reserve_device(ctx->d_in, in_total));
    CUDA_TRY(reserve_device(ctx->d_out, out_total));
    CUDA_TRY(reserve_device(ctx->d_table, table.size()));
    CUDA_TRY(reserve_pinned(ctx->h_stage_in, slot_in * kStreams));
    CUDA_TRY(reserve_pinned(ctx->h_stage_out, slot_out * kStreams));
    for (size_t k = 0; k < items.size(); ++k) {
        table[k].rgba = ctx->d_in.ptr + items[k].in_off;
        table[k].blocks = ctx->d_out.ptr + items[k].out_off;
        table[k].flags = astc_capi::align_flags(table[k].rgba, table[k].pitch);
    }
    cudaError_t err = cudaMemcpyAsync(ctx->d_table.ptr, table.data(), table.size() * sizeof(astc::ImageDesc), cudaMemcpyHostToDevice,
                                      ctx->streams[0]);
    if (err == cudaSuccess) err = cudaEventRecord(ctx->ready, ctx->streams[0]);
    for (int s = 1; s < kStreams && err == cudaSuccess; ++s) err = cudaStreamWaitEvent(ctx->streams[s], ctx->ready, 0);

    // group gi is complete: its slot outputs go to the callers' buffers (the slot is free again afterwards)
    auto retire = [&](size_t gi) -> cudaError_t {
        const Group &g = groups[gi];
        const int s = int(gi % kStreams);
        const cudaError_t e = cudaEventSynchronize(ctx->slot_done[s]);
        if (e != cudaSuccess) return e;
        const uint8_t *slot = ctx->h_stage_out.ptr + size_t(s) * slot_out;
        for (size_t k = g.first; k < g.first + g.count; ++k) {
            const Item &it = items[k];
            if (it.via_out) ctx->pool.copy_rows(it.h_dst, it.out_bytes, slot + (it.out_off - g.out_off), it.out_bytes, it.out_bytes, 1);
        }
        return cudaSuccess;
    };

    // ---- per group: slot fill + uploads, [mip chains,] one launch, downloads; groups rotate over the streams and the slots ----
    size_t retired = 0;
    for (size_t gi = 0; gi < groups.size() && err == cudaSuccess; ++gi) {
        const Group &g = groups[gi];
        const int s = int(gi % kStreams);
        cudaStream_t st = ctx->streams[s];
        if (gi >= size_t(kStreams)) { err = retire(gi - kStreams); ++retired; if (err != cudaSuccess) break; }
        uint8_t *slot = ctx->h_stage_in.ptr + size_t(s) * slot_in;
        size_t k = g.first;
        const size_t end = g.first + g.count;
        while (k < end && err == cudaSuccess) {
            const Item &it = items[k];
            if (!it.h_src) { ++k; continue; }                  // a mip level: produced on the device below
            if (!it.via_in) {
                // Neighbours that lie back to back in the caller's memory AND in the arena (the levels of a chain loaded
                // from one file) travel as ONE copy: every copy costs the DMA engine a start-up of a few microseconds,
                // and four fewer copies per chain take the 512-chain batch from 258.5 to 247.6 ms (same box).
                size_t run_end = k + 1;
                size_t bytes = it.in_bytes;
                const bool tight = it.src_pitch == it.in_pitch && size_t(it.w) * 4u == it.in_pitch;
                while (tight && run_end < end) {
                    const Item &nx = items[run_end];
                    if (!nx.h_src || nx.via_in || nx.src_pitch != nx.in_pitch || size_t(nx.w) * 4u != nx.in_pitch ||
                        nx.h_src != it.h_src + bytes || nx.in_off != it.in_off + bytes)
                        break;
                    bytes += nx.in_bytes;
                    ++run_end;
                }
                if (run_end > k + 1) err = cudaMemcpyAsync(ctx->d_in.ptr + it.in_off, it.h_src, bytes, cudaMemcpyHostToDevice, st);
                else err = upload_rows(ctx->d_in.ptr + it.in_off, it.in_pitch, it.h_src, it.src_pitch, size_t(it.w) * 4u, size_t(it.h), st);
                k = run_end;
                continue;
            }
            // a run of slot images: consecutive in the batch, hence consecutive -- with the same 256-byte padding --
            // in the arena and in the slot: gathered by the host (workers for the large ones), uploaded with one copy
            size_t run_end = k;
            for (; run_end < end && items[run_end].h_src && items[run_end].via_in; ++run_end) {
                const Item &sj = items[run_end];
                ctx->pool.copy_rows(slot + (sj.in_off - g.in_off), sj.in_pitch, sj.h_src, sj.src_pitch, size_t(sj.w) * 4u, size_t(sj.h),
                                    /*streaming=*/sj.in_bytes >= kSmallImage);
                if (sj.chain_head) { ++run_end; break; }       // its level arena follows in the device arena: the run ends here
            }
            const Item &last = items[run_end - 1];
            err = cudaMemcpyAsync(ctx->d_in.ptr + it.in_off, slot + (it.in_off - g.in_off), last.in_off + last.in_bytes - it.in_off,
                                  cudaMemcpyHostToDevice, st);
            k = run_end;
        }
        if (err != cudaSuccess) break;
        if (mips) {
            for (size_t c = g.first; c < end && err == cudaSuccess; ++c) {
                const Item &it = items[c];
                if (!it.chain_head || it.levels == 0) continue;
                uint8_t *ptrs[astc::kMaxMipLevels];
                int ws[astc::kMaxMipLevels], hs[astc::kMaxMipLevels];
                for (int l = 0; l < it.levels; ++l) {
                    const Item &lv = items[c + 1 + size_t(l)];
                    ptrs[l] = ctx->d_in.ptr + lv.in_off;
                    ws[l] = lv.w;
                    hs[l] = lv.h;
                }
                size_t total = 0;
                astc_capi::mip_layout(it.w, it.h, nullptr, nullptr, nullptr, &total);
                unsigned *ticket = reinterpret_cast<unsigned *>(ctx->d_in.ptr + it.mip_off + total - 256u);
                err = cudaMemsetAsync(ticket, 0, sizeof(unsigned), st);
                if (err == cudaSuccess)
                    err = astc::launch_mip_chain(ctx->d_in.ptr + it.in_off, it.w, it.h, it.in_pitch, ptrs, ws, hs, it.levels, ticket, st);
                astc_capi::count_launch();
            }
            if (err != cudaSuccess) break;
        }
        astc::EncodeParams p{};
        p.single = table[g.first];
        p.table = g.count > 1 ? ctx->d_table.ptr + g.first : nullptr;
        p.count = int(g.count);
        p.total_blocks = g.blocks;
        err = astc::launch_encode(d, opt->has_alpha != 0, opt->is_normal_map != 0, opt->srgb != 0, opt->axis_method, p, st);
        if (err != cudaSuccess) break;
        astc_capi::count_launch();
        // outputs: large pinned ones straight to the caller's buffer; each run of slot outputs is ONE copy into the slot
        uint8_t *oslot = ctx->h_stage_out.ptr + size_t(s) * slot_out;
        k = g.first;
        while (k < end && err == cudaSuccess) {
            const Item &it = items[k];
            if (!it.via_out) {
                // (merging neighbouring downloads the way the uploads are merged was measured and is NOT done: the
                // from-bases call went from 167.8 to 172.1 ms -- several medium copies share the link with the uploads
                // better than one large one; profiles/r2ac_batch_copy_merge.txt)
                err = cudaMemcpyAsync(it.h_dst, ctx->d_out.ptr + it.out_off, it.out_bytes, cudaMemcpyDeviceToHost, st);
                ++k;
                continue;
            }
            size_t run_end = k, bytes = 0;
            while (run_end < end && items[run_end].via_out) { bytes += items[run_end].out_bytes; ++run_end; }
            err = cudaMemcpyAsync(oslot + (it.out_off - g.out_off), ctx->d_out.ptr + it.out_off, bytes, cudaMemcpyDeviceToHost, st);
            k = run_end;
        }
        if (err == cudaSuccess) err = cudaEventRecord(ctx->slot_done[s], st);
    }
    for (size_t gi = retired; gi < groups.size() && err == cudaSuccess; ++gi) err = retire(gi);
    err = sync_all(ctx, err);
    if (err != cudaSuccess) return cuda_fail(err, mips ? "astc_b200_context_batch_encode_mip_chains_host" : "astc_b200_context_batch_encode_host");
    return ASTC_B200_OK;
}

}  // namespace

extern "C" {

int astc_b200_context_batch_encode_host(astc_b200_context *ctx, const astc_b200_host_image *images, int count,
                                        const astc_b200_option *opt)
{
    return batch_encode_host_impl(ctx, images, count, opt, /*mips=*/false);
}

int astc_b200_context_batch_encode_mip_chains_host(astc_b200_context *ctx, const astc_b200_host_image *bases, int count,
                                                   const astc_b200_option *opt)
{
    return batch_encode_host_impl(ctx, bases, count, opt, /*mips=*/true);
}

int astc_b200_mip_chain_output_size(int width, int height, const astc_b200_option *opt, size_t *bytes, int *levels)
{
    if (width <= 0 || height <= 0 || !opt) return ASTC_B200_ERR_INVALID_ARGUMENT;
    const int d = dim_of(opt);
    int ws[astc::kMaxMipLevels], hs[astc::kMaxMipLevels];
    const int n = astc_capi::mip_layout(width, height, nullptr, ws, hs, nullptr);
    size_t total = size_t(astc_capi::block_count(width, height, d)) * 16u;
    for (int l = 0; l < n; ++l) total += size_t(astc_capi::block_count(ws[l], hs[l], d)) * 16u;
    if (bytes) *bytes = total;
    if (levels) *levels = n + 1;
    return ASTC_B200_OK;
}

}  // extern "C"
