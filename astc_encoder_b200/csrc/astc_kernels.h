// astc_kernels.h -- internal interface between the C ABI (astc_capi.cu) and
// the sm_100a kernels (astc_kernels.cu).  Not installed.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace astc {

enum : uint32_t {
    kFlagAligned16 = 1u,   // base and pitch are multiples of 16 B (4x4 vector path)
    kFlagAligned8 = 2u,    // base and pitch are multiples of 8 B  (6x6 vector path)
};

// One source texture and its output; what the reference passes through the
// SRV, UAV and constant buffer (astc_encode.h:107-187).
struct ImageDesc {
    const uint8_t *rgba;
    uint8_t *blocks;
    size_t pitch;
    uint64_t first_block;    // prefix sum over the batch
    int32_t width, height;
    uint32_t blocks_x;       // xBlockNum (astc_encode.h:127)
    uint32_t flags;
};

// CTAs [cta_begin, cta_end) encode `passes` runs of blockDim.x consecutive block ids each, starting at first_block.
// A launch is a few segments with shrinking pass counts (set by launch_encode): long CTAs first, short ones last, so
// that the SMs run out of work within one short CTA of each other instead of one long one.
struct Segment {
    uint64_t first_block;
    uint32_t cta_begin, cta_end;
    uint32_t passes;
    uint32_t pad;
};
constexpr int kMaxSegments = 4;

struct EncodeParams {
    ImageDesc single;            // used when table == nullptr
    const ImageDesc *table;      // device array, sorted by first_block
    int32_t count;
    uint64_t total_blocks;
    int32_t passes;              // blocks each thread encodes one after the other (set by launch_encode)
    int32_t nseg;                // 0: every CTA runs `passes` passes
    Segment seg[kMaxSegments];
};

// axis_method: 0 = PCA power iteration (the reference's shipped path), 1 = max_accumulation_pixel_direction
cudaError_t launch_encode(int dim, bool alpha, bool normal, bool srgb, int axis_method, const EncodeParams &p, cudaStream_t stream);
cudaError_t launch_bise(const uint8_t *d_values, int count, int quant, int nseq, uint8_t *d_streams, cudaStream_t stream);
cudaError_t launch_decode(const uint8_t *d_blocks, int width, int height, int dim, uint8_t *d_rgba, size_t pitch,
                          cudaStream_t stream);
cudaError_t launch_downsample2x2(const uint8_t *d_src, int width, int height, size_t src_pitch, uint8_t *d_dst, size_t dst_pitch,
                                 cudaStream_t stream);
constexpr int kMaxMipLevels = 24;        // a 2^24-texel side is the .astc header's limit
// levels 1..count of the mip chain of a width x height image; one fused launch when both sides are multiples of 64
// (d_ticket: a zeroed device word the launch leaves zeroed), else one launch per level
cudaError_t launch_mip_chain(const uint8_t *d_base, int width, int height, size_t base_pitch, uint8_t *const *level_ptrs,
                             const int *level_w, const int *level_h, int count, unsigned *d_ticket, cudaStream_t stream);
cudaError_t launch_mufu(int op, const float *d_x, float *d_y, size_t n, cudaStream_t stream);
const float *host_srgb_lut();
const float *host_unorm_lut();   // the c / 255.0f table the 4x4 kernels look up

}  // namespace astc
