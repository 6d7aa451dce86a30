// astc_kernels.cu -- sm_100a kernels: texel fetch + block encode + store,
// the generic BISE packer (known-answer tests) and the subset decoder.
//
// Replaces MainCS (ASTC_Encode.hlsl:553-582) and the Dispatch geometry of
// astc_encode.h:124-134,190.  Layout in HBM: the source is RGBA8 row-major with
// an arbitrary pitch; the output is one uint4 per block, row-major blocks.
#include "astc_kernels.h"

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdlib>
#include <type_traits>

#include "astc_block.cuh"
#include "astc_schedule.h"

namespace astc {

// ---------------------------------------------------------------------------
// constant tables
// ---------------------------------------------------------------------------
// Encoder tables live in global memory (L2-resident after the first CTA) and are copied to
// shared memory with coalesced 16-byte loads; a per-thread-indexed __constant__ read would
// serialise 32 ways.
__device__ const dev::TableImage g_tables_q6 = dev::make_table_image<QUANT_6>();
__device__ const dev::TableImage g_tables_q12 = dev::make_table_image<QUANT_12>();
__constant__ TritPack c_trit_pack = make_trit_pack();
__constant__ QuintPack c_quint_pack = make_quint_pack();
__constant__ WeightTables c_weight_tables = make_weight_tables();

static const float h_srgb_lut[256] = {
#include "srgb_lut.inc"
};
__device__ const dev::LutImage g_srgb_lut = {{
#include "srgb_lut.inc"
}};

__device__ const dev::LutImage g_unorm_lut = dev::make_unorm_lut();
static const dev::LutImage h_unorm_lut = dev::make_unorm_lut();     // same constexpr generator, host copy for tests

const float *host_srgb_lut() { return h_srgb_lut; }
const float *host_unorm_lut() { return h_unorm_lut.v; }

using dev::f2;
using dev::Texel;

// ---------------------------------------------------------------------------
// UNORM8 -> float
// ---------------------------------------------------------------------------
// c/255.0f correctly rounded for every byte c, as one FMUL + one FFMA (packed: two
// channels per instruction): 1/255 = kRcpHi + kRcpLo to ~2^-48; kRcpLo has 16
// significant bits so c*kRcpLo is exact, and the FMA rounds c*(kRcpHi+kRcpLo)
// once.  Exhaustively equal to IEEE division (tests/test_host_math.py).
constexpr float kRcpHi = 0x1.010102p-8f;                  // RN(1/255)
constexpr float kRcpLo = -0x1.fdfep-33f;

__device__ __forceinline__ f2 unorm2(f2 c)
{
    return dev::fma2(c, dev::bc(kRcpHi), dev::mul2(c, dev::bc(kRcpLo)));
}
__device__ __forceinline__ float unorm1(float c)
{
    return dev::ffma(c, kRcpHi, dev::fmul(c, kRcpLo));
}

// Bytes of a packed RGBA8 word as exact floats: splice each into the mantissa of
// 2^23 (one PRMT) and subtract 2^23 (packed FADD2) -- no I2F.
template <int SEL>
__device__ __forceinline__ float byte_bits(uint32_t w)
{
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | SEL));   // {0x4B,0x00,0x00,byte}
}
__device__ __forceinline__ f2 bytes_lo(uint32_t w)
{
    return dev::add2(dev::mk(byte_bits<0>(w), byte_bits<1>(w)), dev::bc(-8388608.0f));
}
__device__ __forceinline__ f2 bytes_hi(uint32_t w)
{
    return dev::add2(dev::mk(byte_bits<2>(w), byte_bits<3>(w)), dev::bc(-8388608.0f));
}

// LUT entry of byte SEL of a packed texel: two ALU-pipe ops for the address, one LDS.
template <int SEL>
__device__ __forceinline__ float lut_at(const float *lut, uint32_t w)
{
    const uint32_t b = __byte_perm(w, 0u, 0x4440 | SEL);
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(b * 4u + uint32_t(__cvta_generic_to_shared(lut))));
    return v;
}

// One texel: its channel values (UNORM floats, rgb sRGB-decoded for -srgb) + its contribution
// texel*255 to the mean's running sum (ASTC_Encode.hlsl:142-147, 565-580).
//
// The values come from 256-entry tables in shared memory.  Computing c/255 costs the FP32 pipe
// -- the kernel's busiest -- four packed operations per texel; a lookup costs it none (address
// arithmetic runs on the ALU pipe, the load on the LSU) and returns the same bits.
// Sum, linear channels: RN(raw*255) == c for every byte, and the running sum of bytes is an exact
// integer below 2^14, so fma(raw, 255, sum) == sum + c == the reference's rounded product plus
// add.  sRGB channels: the product is rounded (scalar FMUL) before the add.  Alpha is never
// sRGB-decoded.  NORMAL: b = a = 1.0 (:575-578), handled by the caller as constants.
// LUT = false (the 6x6 kernel, whose shared-memory pipe is the busy one): linear channels are
// computed (byte splice + c/255 as FMUL + FFMA) and only the sRGB decode is looked up.
template <bool SRGB, bool NORMAL, bool LUT>
__device__ __forceinline__ Texel convert_texel(uint32_t w, const float *lut_rgb, const float *lut_a, f2 &sum_lo, f2 &sum_hi)
{
    Texel t;
    const f2 k255 = dev::bc(255.0f);
    if (!SRGB && LUT) {
        t.lo = dev::mk(lut_at<0>(lut_rgb, w), lut_at<1>(lut_rgb, w));
        sum_lo = dev::fma2(t.lo, k255, sum_lo);
        if (!NORMAL) {
            t.hi = dev::mk(lut_at<2>(lut_rgb, w), lut_at<3>(lut_rgb, w));
            sum_hi = dev::fma2(t.hi, k255, sum_hi);
        }
    } else if (!SRGB) {
        const f2 clo = bytes_lo(w);
        t.lo = unorm2(clo);
        sum_lo = dev::add2(sum_lo, clo);
        if (!NORMAL) {
            const f2 chi = bytes_hi(w);
            t.hi = unorm2(chi);
            sum_hi = dev::add2(sum_hi, chi);
        }
    } else {
        t.lo = dev::mk(lut_at<0>(lut_rgb, w), lut_at<1>(lut_rgb, w));
        sum_lo = dev::add2(sum_lo, dev::mk(dev::fmul(t.lo.x, 255.0f), dev::fmul(t.lo.y, 255.0f)));
        if (!NORMAL) {
            if (LUT) {
                t.hi = dev::mk(lut_at<2>(lut_rgb, w), lut_at<3>(lut_a, w));
                sum_hi = dev::mk(dev::fadd(sum_hi.x, dev::fmul(t.hi.x, 255.0f)), dev::ffma(t.hi.y, 255.0f, sum_hi.y));
            } else {
                const float ca = dev::fsub(byte_bits<3>(w), 8388608.0f);
                t.hi = dev::mk(lut_at<2>(lut_rgb, w), unorm1(ca));
                sum_hi = dev::add2(sum_hi, dev::mk(dev::fmul(t.hi.x, 255.0f), ca));
            }
        }
    }
    if (NORMAL) t.hi = dev::bc(1.0f);
    return t;
}

// ---------------------------------------------------------------------------
// block walk
// ---------------------------------------------------------------------------
// A thread visits the block ids  first, first + stride, first + 2*stride, ...  The walk keeps
// the block's image and its (bx, by) incrementally: one division when it starts and one each
// time it crosses the end of a block row or of an image; the other steps are adds.  (Dividing
// the linear id for every block, as MainCS does at ASTC_Encode.hlsl:561-562, costs ~25 issue
// slots of a ~1100-slot block.)
template <bool BATCH>
struct Walk {
    uint32_t local;          // block index inside the image
    uint32_t bx, by;
    int idx;                 // image index (batch only)

    __device__ __forceinline__ const ImageDesc &desc(const EncodeParams &p) const { return BATCH ? p.table[idx] : p.single; }
    __device__ __forceinline__ uint32_t image_blocks(const EncodeParams &p) const
    {
        if (!BATCH) return uint32_t(p.total_blocks);
        const uint64_t next = idx + 1 < p.count ? __ldg(&p.table[idx + 1].first_block) : p.total_blocks;
        return uint32_t(next - __ldg(&p.table[idx].first_block));
    }
    __device__ __forceinline__ void divide(uint32_t blocks_x)
    {
        by = local / blocks_x;
        bx = local - by * blocks_x;
    }
    __device__ __forceinline__ bool start(const EncodeParams &p, uint64_t id)
    {
        if (id >= p.total_blocks) return false;
        idx = 0;
        uint64_t first = 0;
        if (BATCH) {
            const ImageDesc *__restrict__ table = p.table;
            int lo = 0, hi = p.count - 1;
            while (lo < hi) {                              // last image with first_block <= id
                const int mid = (lo + hi + 1) >> 1;
                if (__ldg(&table[mid].first_block) <= id) lo = mid; else hi = mid - 1;
            }
            idx = lo;
            first = __ldg(&table[lo].first_block);
        }
        local = uint32_t(id - first);
        divide(BATCH ? __ldg(&p.table[idx].blocks_x) : p.single.blocks_x);
        return true;
    }
    __device__ __forceinline__ bool advance(const EncodeParams &p, uint32_t stride)
    {
        local += stride;
        bx += stride;
        uint32_t n = image_blocks(p);
        if (local >= n) {
            if (!BATCH) return false;
            do {
                local -= n;
                if (++idx >= p.count) return false;
                n = image_blocks(p);
            } while (local >= n);
            divide(__ldg(&p.table[idx].blocks_x));
        } else {
            const uint32_t blocks_x = BATCH ? __ldg(&p.table[idx].blocks_x) : p.single.blocks_x;
#ifdef ASTC_WALK_DIVIDE_ALWAYS
            if (bx >= blocks_x) divide(blocks_x);
#else
            if (bx >= blocks_x) {                              // crossed the end of a block row: usually into the next one
                bx -= blocks_x;
                ++by;
                if (bx >= blocks_x) divide(blocks_x);          // rows shorter than the stride
            }
#endif
        }
        return true;
    }
    __device__ __forceinline__ uint4 *out(const EncodeParams &p) const
    {
        return reinterpret_cast<uint4 *>(BATCH ? p.table[idx].blocks : p.single.blocks) + local;
    }
};

// Which ids this CTA encodes: its pass count and the id of thread 0's first block (Segment, astc_kernels.h).
__device__ __forceinline__ uint64_t cta_schedule(const EncodeParams &p, uint32_t threads, int &passes)
{
    if (p.nseg == 0) {
        passes = p.passes;
        return uint64_t(blockIdx.x) * uint32_t(p.passes * int(threads));
    }
    // constant indices only: a dynamically indexed kernel parameter would be copied to local memory
    uint32_t q = p.seg[0].passes, begin = 0;
    uint64_t first = 0;
#pragma unroll
    for (int i = 1; i < kMaxSegments; ++i) {
        if (i < p.nseg && blockIdx.x >= p.seg[i].cta_begin) {
            q = p.seg[i].passes;
            begin = p.seg[i].cta_begin;
            first = p.seg[i].first_block;
        }
    }
    passes = int(q);
    return first + uint64_t(blockIdx.x - begin) * uint32_t(q * threads);
}

// The per-CTA tables (immutable module globals) travel global -> registers -> shared memory in two steps, so that
// other start-up work can be issued while the loads are in flight: one 16-byte piece per thread and table.
struct TableRegs {
    uint4 pack, lut, lut_a;
};
template <bool ALPHA, bool SRGB, bool UNORM_LUT, bool ALPHA_LUT>
__device__ __forceinline__ TableRegs fetch_tables()
{
    static_assert(sizeof(dev::TableImage) % 16 == 0 && offsetof(dev::SharedTables, lut_rgb) == sizeof(dev::TableImage), "table layout");
    constexpr int kPieces = int(sizeof(dev::TableImage) / 16);
    static_assert(kPieces <= 128, "one piece per thread of a 128-thread CTA");
    TableRegs r{};
    const int i = threadIdx.x;
    if (i < kPieces) r.pack = __ldg(reinterpret_cast<const uint4 *>(ALPHA ? &g_tables_q6 : &g_tables_q12) + i);
    if ((SRGB || UNORM_LUT) && i < 64) r.lut = __ldg(reinterpret_cast<const uint4 *>(SRGB ? &g_srgb_lut : &g_unorm_lut) + i);
    if (ALPHA_LUT && i >= 64 && i < 128) r.lut_a = __ldg(reinterpret_cast<const uint4 *>(&g_unorm_lut) + (i - 64));
    return r;
}
template <bool SRGB, bool UNORM_LUT, bool ALPHA_LUT>
__device__ __forceinline__ void store_tables(dev::SharedTables &st, float *lut_a, const TableRegs &r)
{
    constexpr int kPieces = int(sizeof(dev::TableImage) / 16);
    const int i = threadIdx.x;
    if (i < kPieces) reinterpret_cast<uint4 *>(&st)[i] = r.pack;
    if ((SRGB || UNORM_LUT) && i < 64) reinterpret_cast<uint4 *>(st.lut_rgb)[i] = r.lut;
    if (ALPHA_LUT && i >= 64 && i < 128) reinterpret_cast<uint4 *>(lut_a)[i - 64] = r.lut_a;
}

// Programmatic dependent launch (sm_90+).  The encode kernels are launched with the programmatic-stream-serialization
// attribute: a launch may become resident while the kernel before it in the stream is still draining.  Everything a
// CTA does before pdl_wait() touches only kernel parameters and immutable module globals (the table fetch above);
// pdl_wait() returns once the preceding grid has completed and its memory is visible, so every read of the source
// texture, of the batch's descriptor table and every store comes after it.  pdl_trigger() lets the NEXT launch start
// filling SM slots as soon as all CTAs of this one have been dispatched.  Both are no-ops in an ordinary launch.
// Measured on B200 (round 2q, same-box A/B over back-to-back launches, profiles/r2q_ab_pdl.txt): 2048^2 16.6 -> 14.5 us
// per launch, 4096^2 49.6 -> 47.0, 1/8 band of 16384^2 91.1 -> 89.1, 16384^2 682.2 -> 680.5; isolated launches unchanged.
// -DASTC_PDL=0 builds plain launches.
#ifndef ASTC_PDL
#define ASTC_PDL 1
#endif
__device__ __forceinline__ void pdl_trigger()
{
#if ASTC_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait()
{
#if ASTC_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

// ---------------------------------------------------------------------------
// 4x4: sixteen texels live in registers as UNORM float pairs.
// ---------------------------------------------------------------------------
struct Texels4x4 {
    static constexpr bool kStreamed = false;
    Texel t[16];
    __device__ __forceinline__ Texel raw(int k) const { return t[k]; }
    __device__ __forceinline__ void fence() const {}
};

#ifndef ASTC_THREADS_4X4
#define ASTC_THREADS_4X4 128
#endif
#ifndef ASTC_MINBLOCKS_4X4
#define ASTC_MINBLOCKS_4X4 4
#endif
constexpr int kThreads4x4 = ASTC_THREADS_4X4;
constexpr int kMaxPasses = 8;  // most blocks a thread encodes one after the other

// The four 16-byte texel rows of the thread's next block travel global -> shared memory by
// cp.async (LDGSTS) while the current block is being encoded: the copy holds no registers
// (a register prefetch costs 16 and pushed the kernel into spills) and needs no barrier, since a
// thread reads back only the slots it filled itself -- cp.async.wait_group is the whole handshake.
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}

// Returns whether the block took the vector path (whole block inside the image, 16-byte aligned rows).
template <bool BATCH>
__device__ __forceinline__ bool prefetch_rows4x4(const EncodeParams &p, const Walk<BATCH> &wk, uint32_t slot)
{
    const ImageDesc &d = wk.desc(p);
    const uint32_t x0 = wk.bx * 4u, y0 = wk.by * 4u;
    const bool fast = (d.flags & kFlagAligned16) && x0 + 4u <= uint32_t(d.width) && y0 + 4u <= uint32_t(d.height);
    if (fast) {
        // adjacent threads copy adjacent 16 B: 512 contiguous bytes per warp per texel row
        const size_t pitch = d.pitch;
        const uint8_t *src = d.rgba + size_t(y0) * pitch + size_t(x0) * 4u;
#pragma unroll
        for (int r = 0; r < 4; ++r, src += pitch) cp_async16(slot + uint32_t(r) * (kThreads4x4 * 16u), src);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    return fast;
}

#ifndef ASTC_MINBLOCKS_4X4_NORMAL
#define ASTC_MINBLOCKS_4X4_NORMAL 7
#endif
template <bool ALPHA, bool NORMAL, bool SRGB, bool BATCH, bool ACCUM>
__global__ void __launch_bounds__(kThreads4x4, NORMAL ? ASTC_MINBLOCKS_4X4_NORMAL : ASTC_MINBLOCKS_4X4)
encode4x4_kernel(const EncodeParams p)
{
    __shared__ dev::SharedTables st;
    __shared__ uint4 s_rows[2][4][kThreads4x4];                 // cp.async landing slots, double-buffered
    __shared__ __align__(16) float s_lut_a[SRGB ? 256 : 4];
    const uint32_t slot0 = smem_addr(&s_rows[0][0][threadIdx.x]);
    constexpr uint32_t kSlotStride = 4u * kThreads4x4 * 16u;

    // CTA b owns ids [b*BPT*T, (b+1)*BPT*T); pass i takes the i-th run of T consecutive ids, so a
    // warp reads 512 contiguous bytes per texel row and stores 512 contiguous bytes.
    // The tables are requested (into registers), then the first block's rows (cp.async), and only then are the tables
    // stored to shared memory: the two HBM / L2 round trips of a CTA's start-up overlap instead of following each other.
    pdl_trigger();
    const TableRegs tables = fetch_tables<ALPHA, SRGB, true, SRGB>();      // in flight while the first rows are requested
    pdl_wait();
    Walk<BATCH> wk;
    int passes;
    const uint64_t cta_first = cta_schedule(p, kThreads4x4, passes);
    const bool any = wk.start(p, cta_first + threadIdx.x);
    bool fast = any && prefetch_rows4x4<BATCH>(p, wk, slot0);
    store_tables<SRGB, true, SRGB>(st, s_lut_a, tables);
    __syncthreads();
    if (!any) return;
    const uint32_t s_field = smem_addr(st.field), s_trit = smem_addr(st.trit_scattered);
#pragma unroll 1
    for (int pass = 0;; ++pass) {
        Texels4x4 tx;
        f2 sum_lo = dev::bc(0.f), sum_hi = dev::bc(0.f);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (fast) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint4 row = s_rows[pass & 1][r][threadIdx.x];
                tx.t[4 * r + 0] = convert_texel<SRGB, NORMAL, true>(row.x, st.lut_rgb, s_lut_a, sum_lo, sum_hi);
                tx.t[4 * r + 1] = convert_texel<SRGB, NORMAL, true>(row.y, st.lut_rgb, s_lut_a, sum_lo, sum_hi);
                tx.t[4 * r + 2] = convert_texel<SRGB, NORMAL, true>(row.z, st.lut_rgb, s_lut_a, sum_lo, sum_hi);
                tx.t[4 * r + 3] = convert_texel<SRGB, NORMAL, true>(row.w, st.lut_rgb, s_lut_a, sum_lo, sum_hi);
            }
        } else {
            // edge / unaligned: per-texel loads, out-of-range texels read as 0
            // like Texture2D.Load (ASTC_Encode.hlsl:574); the UNORM / sRGB value of byte 0 is 0.
            const ImageDesc &d = wk.desc(p);
            const uint32_t x0 = wk.bx * 4u, y0 = wk.by * 4u;
            const size_t pitch = d.pitch;
            const uint32_t width = uint32_t(d.width), height = uint32_t(d.height);
            const uint8_t *base = d.rgba + size_t(y0) * pitch + size_t(x0) * 4u;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const bool inside = x0 + (k & 3) < width && y0 + (k >> 2) < height;
                const uint32_t w = inside ? __ldg((const uint32_t *)(base + size_t(k >> 2) * pitch + size_t(k & 3) * 4u)) : 0u;
                tx.t[k] = convert_texel<SRGB, NORMAL, true>(w, st.lut_rgb, s_lut_a, sum_lo, sum_hi);
            }
        }
        uint4 *const out = wk.out(p);
        const bool more = pass + 1 < passes && wk.advance(p, kThreads4x4);
        if (more) fast = prefetch_rows4x4<BATCH>(p, wk, slot0 + uint32_t((pass + 1) & 1) * kSlotStride);
        // one coalesced 16-byte store per thread
        *out = dev::encode_block<4, ALPHA, NORMAL, ACCUM>(tx, sum_lo, sum_hi, s_field, s_trit);
        if (!more) break;
    }
}

// ---------------------------------------------------------------------------
// 6x6: 36 texels (144 floats) do not fit the registers of a thread that is to share its SM with 15
// other warps.  Each thread parks the first kPark6x6 texels of its block as float4 in its own
// shared-memory column (bank = lane: conflict-free LDS.128 / STS.128) and keeps the last ten in
// registers; that is 416 B of shared memory and <= 128 registers per thread, i.e. FOUR 128-thread CTAs
// per SM.  (All 36 parked and 168 registers -- the first version -- gave three; measured on one box,
// 8192^2: t = 0.110 + 0.242 / CTAs-per-SM ms, the kernel was latency-bound.)
//
// ptxas and the passes over the texels.  Handed a fully unrolled pass it hoists all of its LDS.128 to
// the top and, short of registers, spills them (800 B of spills measured).  So the covariance and
// min/max passes run the four texel rows that lie wholly in shared memory as a ROLLED loop and only the
// tail unrolled (for_each_texel_streamed in astc_block.cuh), and the weight pass, whose 64 taps
// cannot be rolled, is cut by a scheduling fence every four grid points (__syncwarp: an empty asm
// never reaches ptxas).  A compiler fence after parking stops NVVM from forwarding the parked values
// to the covariance pass in registers, which is what used to cost the 168 registers.
//
// Normal maps: b = a = 1 for every texel, so only (r, g) is parked (float2) and six CTAs fit.
// ---------------------------------------------------------------------------
#ifndef ASTC_THREADS_6X6
#define ASTC_THREADS_6X6 128
#endif
constexpr int kThreads6x6 = ASTC_THREADS_6X6;
// How many of a block's 36 texels are parked in shared memory (the rest stay in registers), per variant.  Every parked
// texel costs one 16-byte store and three 16-byte loads per block, and the 6x6 kernels run at ~3/4 of the SM's
// shared-memory bandwidth (40 M wavefronts per 8192^2 launch, 0.74 per clock and SM) -- so, registers permitting, fewer is
// faster.  With the texel-major weight pass there is room for more register texels than the 10 of round 1
// (measured, round 2af: profiles/r2af_ab_6x6_park.txt).
#ifndef ASTC_6X6_PARK_LINEAR
#define ASTC_6X6_PARK_LINEAR 24
#endif
#ifndef ASTC_6X6_PARK_SRGB
#define ASTC_6X6_PARK_SRGB 18
#endif
#ifndef ASTC_6X6_PARK_NORMAL
#define ASTC_6X6_PARK_NORMAL 16
#endif
template <bool NORMAL, bool SRGB>
constexpr int park6x6() { return NORMAL ? ASTC_6X6_PARK_NORMAL : SRGB ? ASTC_6X6_PARK_SRGB : ASTC_6X6_PARK_LINEAR; }
#ifndef ASTC_EXTRA_SMEM_6X6
#define ASTC_EXTRA_SMEM_6X6 0            // occupancy experiments only (tools/variants.py)
#endif

template <bool NORMAL, int kPark6x6>
struct Texels6x6 {
    using Slot = typename std::conditional<NORMAL, float2, float4>::type;
    static constexpr bool kStreamed = true;
    static constexpr int kLoopRows = kPark6x6 / 6 < 4 ? kPark6x6 / 6 : 4;   // the texel rows that lie wholly in shared memory (rows 0..3 with 26 parked)
    Slot *col;                                            // &smem[threadIdx.x], stride kThreads6x6
    Texel reg[36 - kPark6x6];
    __device__ __forceinline__ void put(int k, const Texel &t)
    {
        if (k >= kPark6x6) reg[k - kPark6x6] = t;
        else if constexpr (NORMAL) col[k * kThreads6x6] = make_float2(t.lo.x, t.lo.y);
        else {
            // Written as two 8-byte stores on purpose.  ptxas fuses them back into one STS.128, but allocates the two
            // packed results (64-bit register pairs) side by side; handed a float4 it re-assembled every texel in a
            // staging quad with four MOVs -- 104 of a block's 2100 instructions (8192^2 RGB 0.178 -> 0.174 ms, round 2h).
            const uint32_t a = uint32_t(__cvta_generic_to_shared(col + k * kThreads6x6));
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(t.lo.x), "f"(t.lo.y) : "memory");
            asm volatile("st.shared.v2.f32 [%0+8], {%1, %2};" ::"r"(a), "f"(t.hi.x), "f"(t.hi.y) : "memory");
        }
    }
    __device__ __forceinline__ Texel raw_dyn(int k) const
    {
        const Slot v = col[k * kThreads6x6];
        if constexpr (NORMAL) return Texel{dev::mk(v.x, v.y), dev::bc(1.0f)};
        else return Texel{dev::mk(v.x, v.y), dev::mk(v.z, v.w)};
    }
    __device__ __forceinline__ Texel raw(int k) const { return k < kPark6x6 ? raw_dyn(k) : reg[k - kPark6x6]; }
    __device__ __forceinline__ void fence() const { asm volatile("" ::: "memory"); }
    // a fence ptxas honours when it schedules; the warp is converged wherever this is called
    __device__ __forceinline__ void sched_fence() const { __syncwarp(); }
};

template <bool NORMAL, bool SRGB>
constexpr size_t smem6x6()
{
    return size_t(park6x6<NORMAL, SRGB>()) * kThreads6x6 * sizeof(typename Texels6x6<NORMAL, park6x6<NORMAL, SRGB>()>::Slot) +
           sizeof(dev::SharedTables) + ASTC_EXTRA_SMEM_6X6;
}
#ifndef ASTC_6X6_CTAS
#define ASTC_6X6_CTAS 4
#endif
template <bool NORMAL>
constexpr int ctas6x6() { return NORMAL ? 6 : ASTC_6X6_CTAS; }

template <bool ALPHA, bool NORMAL, bool SRGB, bool BATCH, bool ACCUM>
__global__ void __launch_bounds__(kThreads6x6, ctas6x6<NORMAL>())
encode6x6_kernel(const EncodeParams p)
{
    constexpr int kPark6x6 = park6x6<NORMAL, SRGB>();
    using TX = Texels6x6<NORMAL, kPark6x6>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename TX::Slot *s_tex = reinterpret_cast<typename TX::Slot *>(smem_raw);         // [kPark6x6][kThreads6x6]
    dev::SharedTables &st = *reinterpret_cast<dev::SharedTables *>(s_tex + kPark6x6 * kThreads6x6);
    // CTA b owns ids [b*BPT*T, (b+1)*BPT*T); a thread re-uses its own shared-memory column for
    // each of its blocks (only it reads or writes that column: no barrier between passes).
    // The first block's rows are pulled towards L2 before the tables are fetched, so that the two round trips of
    // a CTA's start-up overlap.
    pdl_trigger();
    const TableRegs tables = fetch_tables<ALPHA, SRGB, false, false>();
    pdl_wait();
    Walk<BATCH> wk;
    int passes;
    const uint64_t cta_first = cta_schedule(p, kThreads6x6, passes);
    const bool any = wk.start(p, cta_first + threadIdx.x);
    if (any) {
        const ImageDesc &d0 = wk.desc(p);
        if (wk.by * 6u + 6u <= uint32_t(d0.height) && wk.bx * 6u + 6u <= uint32_t(d0.width)) {
            const uint8_t *b0 = d0.rgba + size_t(wk.by * 6u) * d0.pitch + size_t(wk.bx * 6u) * 4u;
#pragma unroll
            for (int r = 0; r < 6; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + size_t(r) * d0.pitch));
        }
    }
    store_tables<SRGB, false, false>(st, nullptr, tables);
    __syncthreads();
    if (!any) return;
    const uint32_t s_field = smem_addr(st.field), s_trit = smem_addr(st.trit_scattered);
#pragma unroll 1
    for (int pass = 0;; ++pass) {
        f2 sum_lo = dev::bc(0.f), sum_hi = dev::bc(0.f);
        TX tx;
        tx.col = s_tex + threadIdx.x;
        const ImageDesc &d = wk.desc(p);
        const uint32_t x0 = wk.bx * 6u, y0 = wk.by * 6u;
        const size_t pitch = d.pitch;
        const uint32_t width = uint32_t(d.width), height = uint32_t(d.height);
        const uint8_t *base = d.rgba + size_t(y0) * pitch + size_t(x0) * 4u;
        auto park = [&](int k, uint32_t w) { tx.put(k, convert_texel<SRGB, NORMAL, false>(w, st.lut_rgb, nullptr, sum_lo, sum_hi)); };
        if ((d.flags & kFlagAligned8) && x0 + 6u <= width && y0 + 6u <= height) {
            // interior: a block row is 24 B = three 8-byte loads; a warp covers 768
            // contiguous bytes per texel row.  All 18 loads are issued before the first use.
            uint2 rows[18];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const uint2 *src = (const uint2 *)(base + size_t(r) * pitch);
                rows[3 * r + 0] = __ldg(src); rows[3 * r + 1] = __ldg(src + 1); rows[3 * r + 2] = __ldg(src + 2);
            }
            // The thread's next block: pull its six rows into L2 now, so the loads above hit L2
            // instead of HBM one pass later.
            if (pass + 1 < passes) {
                Walk<BATCH> nx = wk;
                if (nx.advance(p, kThreads6x6)) {
                    const ImageDesc &dn = nx.desc(p);
                    const size_t pn = dn.pitch;
                    const uint8_t *bn = dn.rgba + size_t(nx.by * 6u) * pn + size_t(nx.bx * 6u) * 4u;
                    if (nx.by * 6u + 6u <= uint32_t(dn.height) && nx.bx * 6u + 6u <= uint32_t(dn.width)) {
#pragma unroll
                        for (int r = 0; r < 6; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(bn + size_t(r) * pn));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 18; ++i) { park(2 * i, rows[i].x); park(2 * i + 1, rows[i].y); }
        } else {
            // edge / unaligned: per-texel loads, out-of-range texels read as 0 like Texture2D.Load
            // (ASTC_Encode.hlsl:574).  Unrolled: the texels kept in registers need constant indices.
#pragma unroll
            for (int k = 0; k < 36; ++k) {
                const uint32_t kx = k % 6, ky = k / 6;
                const bool inside = x0 + kx < width && y0 + ky < height;
                park(k, inside ? __ldg((const uint32_t *)(base + size_t(ky) * pitch + size_t(kx) * 4u)) : 0u);
            }
        }
        if (NORMAL) sum_hi = dev::bc(36.0f * 255.0f);
        tx.fence();                                        // no store-to-load forwarding: the passes stream from shared memory
        *wk.out(p) = dev::encode_block<6, ALPHA, NORMAL, ACCUM>(tx, sum_lo, sum_hi, s_field, s_trit);
        if (pass + 1 >= passes || !wk.advance(p, kThreads6x6)) break;
    }
}


// ---------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------
// Blocks per thread.  More passes amortise the per-CTA table load and the first, unprefetched block
// (4x4, 16384^2: 0.87 ms at 1 pass, 0.78 at 2, 0.75 at 4, 0.736 at 8), but the grid should keep about
// three waves of resident CTAs or the last partial wave idles the SMs (4096^2 is best at 4 passes).
// The 6x6 kernel has no prefetch to amortise and is best at 1-3 passes (measured on 8192^2).
// SMs of the current device, asked once per (host thread, device).
static int sm_count_of_current_device()
{
    static thread_local int cached_device = -1, cached_count = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_device) {
        int n = 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
        cached_device = dev;
        cached_count = n;
    }
    return cached_count;
}

static int choose_passes(uint64_t total_blocks, int threads, int ctas_per_sm, int max_passes)
{
    const int sm_count = sm_count_of_current_device();
#ifdef ASTC_TUNING_HOOKS
    // experiment builds only (tools/variants.py build hooks=ASTC_TUNING_HOOKS): never compiled into the product library
    const char *force = getenv("ASTC_B200_PASSES");
    if (force && atoi(force) > 0) return atoi(force) > kMaxPasses ? kMaxPasses : atoi(force);
#endif
    const uint64_t resident = uint64_t(sm_count) * uint64_t(ctas_per_sm);
    const uint64_t want = total_blocks / (resident * 3u * uint64_t(threads));
    return want < 1 ? 1 : want > uint64_t(max_passes) ? max_passes : int(want);
}

// Measured on B200 (round 2n, same-box A/B, profiles/r2n_ab_taper.txt): 1/8 band of 16384^2 98.3 -> 96.3 us, 1/4 band
// 184.3 -> 182.3 us, 4096^2 6x6 -alpha -srgb 67.6 -> 65.5 us, large launches unchanged; bit-exact.  -DASTC_TAPER=0 builds the uniform schedule.
#ifndef ASTC_TAPER
#define ASTC_TAPER 1
#endif
static uint64_t plan_schedule(EncodeParams &p, int threads, int ctas_per_sm)
{
    return plan_tapered(p, threads, uint64_t(sm_count_of_current_device()) * uint64_t(ctas_per_sm), ASTC_TAPER != 0);
}

template <typename K>
static cudaError_t launch_pdl(K kern, unsigned ctas, int threads, size_t smem, cudaStream_t stream, const EncodeParams &p)
{
#if ASTC_PDL
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(unsigned(threads));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
#else
    kern<<<ctas, threads, smem, stream>>>(p);
    return cudaGetLastError();
#endif
}

template <bool ALPHA, bool NORMAL, bool SRGB, bool BATCH, bool ACCUM>
static cudaError_t launch_variant(int dim, EncodeParams p, cudaStream_t stream)
{
    if (dim == 4) {
        p.passes = choose_passes(p.total_blocks, kThreads4x4, NORMAL ? ASTC_MINBLOCKS_4X4_NORMAL : ASTC_MINBLOCKS_4X4, kMaxPasses);
        const uint64_t ctas = plan_schedule(p, kThreads4x4, NORMAL ? ASTC_MINBLOCKS_4X4_NORMAL : ASTC_MINBLOCKS_4X4);
        if (ctas > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
        return launch_pdl(encode4x4_kernel<ALPHA, NORMAL, SRGB, BATCH, ACCUM>, unsigned(ctas), kThreads4x4, 0, stream, p);
    } else {
        auto kern = encode6x6_kernel<ALPHA, NORMAL, SRGB, BATCH, ACCUM>;
        constexpr size_t kSmem6x6 = smem6x6<NORMAL, SRGB>();
        constexpr int kCtas6x6 = ctas6x6<NORMAL>();
        p.passes = choose_passes(p.total_blocks, kThreads6x6, kCtas6x6, 2);
        const uint64_t ctas = plan_schedule(p, kThreads6x6, kCtas6x6);
        if (ctas > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
        static thread_local int configured_device = -1;
        int devno = 0;
        cudaGetDevice(&devno);
        if (configured_device != devno) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmem6x6));
            if (e != cudaSuccess) return e;
            configured_device = devno;
        }
        return launch_pdl(kern, unsigned(ctas), kThreads6x6, kSmem6x6, stream, p);
    }
}

template <bool BATCH, bool ACCUM>
static cudaError_t dispatch(int dim, bool alpha, bool normal, bool srgb, const EncodeParams &p, cudaStream_t s)
{
    // IS_NORMALMAP / HAS_ALPHA macros of astc_encode.h:55-62 become template
    // parameters; sRGB is dropped for normal maps (main.cpp:214).
    if (normal) srgb = false;
    const int key = (alpha ? 4 : 0) | (normal ? 2 : 0) | (srgb ? 1 : 0);
    switch (key) {
    case 0: return launch_variant<false, false, false, BATCH, ACCUM>(dim, p, s);
    case 1: return launch_variant<false, false, true, BATCH, ACCUM>(dim, p, s);
    case 2: return launch_variant<false, true, false, BATCH, ACCUM>(dim, p, s);
    case 4: return launch_variant<true, false, false, BATCH, ACCUM>(dim, p, s);
    case 5: return launch_variant<true, false, true, BATCH, ACCUM>(dim, p, s);
    case 6: return launch_variant<true, true, false, BATCH, ACCUM>(dim, p, s);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_encode(int dim, bool alpha, bool normal, bool srgb, int axis_method, const EncodeParams &p, cudaStream_t stream)
{
    if (p.total_blocks == 0) return cudaSuccess;
    const bool batch = p.count > 1 || p.table != nullptr;
    if (axis_method == 1)
        return batch ? dispatch<true, true>(dim, alpha, normal, srgb, p, stream) : dispatch<false, true>(dim, alpha, normal, srgb, p, stream);
    return batch ? dispatch<true, false>(dim, alpha, normal, srgb, p, stream) : dispatch<false, false>(dim, alpha, normal, srgb, p, stream);
}

// ---------------------------------------------------------------------------
// Generic BISE packer: bits, trits and quints for every QUANT_* level
// (ASTC_IntegerSequenceEncoding.hlsl:142-276).  One thread per sequence.
// ---------------------------------------------------------------------------
struct BitStream128 {
    uint64_t lo = 0, hi = 0;
    uint32_t pos = 0;
    // orbits8_ptr (:98-119) without its shift-by-32 hazard
    __device__ __forceinline__ void put(uint32_t value, uint32_t count)
    {
        if (count == 0 || pos >= 128) { pos += count; return; }
        const uint64_t v = uint64_t(value) & ((1ull << count) - 1ull);
        if (pos < 64) {
            lo |= v << pos;
            if (pos + count > 64) hi |= v >> (64 - pos);
        } else {
            hi |= v << (pos - 64);
        }
        pos += count;
    }
};

__global__ void bise_encode_kernel(const uint8_t *__restrict__ values, int count, int quant, int nseq,
                                   uint4 *__restrict__ streams)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseq) return;
    const uint8_t *v = values + size_t(s) * size_t(count);
    const QuantLayout l = quant_layout(quant);
    const uint32_t mask = (1u << l.bits) - 1u;
    BitStream128 bs;
    if (l.trits) {
        for (int i = 0; i < count; i += 5) {
            uint32_t g[5];
            for (int j = 0; j < 5; ++j) g[j] = i + j < count ? v[i + j] : 0u;
            const uint32_t T = c_trit_pack.v[(g[4] >> l.bits) * 81 + (g[3] >> l.bits) * 27 + (g[2] >> l.bits) * 9 +
                                             (g[1] >> l.bits) * 3 + (g[0] >> l.bits)];
            bs.put(g[0] & mask, l.bits); bs.put(T & 3u, 2);
            bs.put(g[1] & mask, l.bits); bs.put((T >> 2) & 3u, 2);
            bs.put(g[2] & mask, l.bits); bs.put((T >> 4) & 1u, 1);
            bs.put(g[3] & mask, l.bits); bs.put((T >> 5) & 3u, 2);
            bs.put(g[4] & mask, l.bits); bs.put((T >> 7) & 1u, 1);
        }
    } else if (l.quints) {
        for (int i = 0; i < count; i += 3) {
            uint32_t g[3];
            for (int j = 0; j < 3; ++j) g[j] = i + j < count ? v[i + j] : 0u;
            const uint32_t Q = c_quint_pack.v[(g[2] >> l.bits) * 25 + (g[1] >> l.bits) * 5 + (g[0] >> l.bits)];
            bs.put(g[0] & mask, l.bits); bs.put(Q & 7u, 3);
            bs.put(g[1] & mask, l.bits); bs.put((Q >> 3) & 3u, 2);
            bs.put(g[2] & mask, l.bits); bs.put((Q >> 5) & 3u, 2);
        }
    } else {
        for (int i = 0; i < count; ++i) bs.put(v[i], l.bits);
    }
    streams[s] = make_uint4(uint32_t(bs.lo), uint32_t(bs.lo >> 32), uint32_t(bs.hi), uint32_t(bs.hi >> 32));
}

cudaError_t launch_bise(const uint8_t *d_values, int count, int quant, int nseq, uint8_t *d_streams, cudaStream_t stream)
{
    if (nseq == 0) return cudaSuccess;
    bise_encode_kernel<<<(nseq + 127) / 128, 128, 0, stream>>>(d_values, count, quant, nseq, (uint4 *)d_streams);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Subset decoder (ASTC spec C.2; the reference has none).  One thread per
// block; handles what this encoder emits: 1 partition, 1 plane, CEM 8 / 12,
// 8-bit endpoints, 4x4 weight grid with QUANT_6 / QUANT_12 weights.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bits128(const uint4 b, uint32_t pos, uint32_t count)
{
    const uint32_t w[5] = {b.x, b.y, b.z, b.w, 0u};
    const uint32_t i = pos >> 5, s = pos & 31u;
    const uint64_t two = (uint64_t(w[i + 1 > 4 ? 4 : i + 1]) << 32) | w[i];
    return uint32_t(two >> s) & ((1u << count) - 1u);
}

__global__ void decode_kernel(const uint4 *__restrict__ blocks, int width, int height, int dim,
                              uint8_t *__restrict__ rgba, size_t pitch, uint32_t total, uint32_t bw)
{
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    const uint4 b = __ldg(blocks + id);
    const uint32_t mode = b.x & 0x7FFu, cem = (b.x >> 13) & 0xFu, parts = (b.x >> 11) & 3u;
    const uint32_t by = id / bw, bx = id - by * bw;
    const bool rgba_mode = mode == blockmode_4x4grid(QUANT_6) && cem == CEM_LDR_RGBA_DIRECT;
    const bool rgb_mode = mode == blockmode_4x4grid(QUANT_12) && cem == CEM_LDR_RGB_DIRECT;
    int wq[16];
    int e0[4], e1[4];
    if ((rgba_mode || rgb_mode) && parts == 0) {
        const int method = rgba_mode ? QUANT_6 : QUANT_12;
        const uint32_t n = rgba_mode ? 1u : 2u, gbits = 5u * n + 8u;
        // weight stream = block bits read downward from bit 127
        const uint4 rev = make_uint4(__brev(b.w), __brev(b.z), __brev(b.y), __brev(b.x));
        for (int g = 0; g < 4; ++g) {
            const uint32_t grp = bits128(rev, g * gbits, gbits);
            const uint32_t T = ((grp >> n) & 3u) | (((grp >> (2 * n + 2)) & 3u) << 2) | (((grp >> (3 * n + 4)) & 1u) << 4) |
                               (((grp >> (4 * n + 5)) & 3u) << 5) | (((grp >> (5 * n + 7)) & 1u) << 7);
            const Trits t = trits_from_integer(int(T));
            const uint32_t m[5] = {grp & ((1u << n) - 1u), (grp >> (n + 2)) & ((1u << n) - 1u), (grp >> (2 * n + 4)) & ((1u << n) - 1u),
                                   (grp >> (3 * n + 5)) & ((1u << n) - 1u), (grp >> (4 * n + 7)) & ((1u << n) - 1u)};
            for (int j = 0; j < 5 && 5 * g + j < 16; ++j)
                wq[5 * g + j] = c_weight_tables.unq[method][((uint32_t(t.t[j]) << n) | m[j]) & 31u];
        }
        int ep[8];
        for (int i = 0; i < 8; ++i) ep[i] = int(bits128(b, 17u + 8u * i, 8));
        if (rgb_mode) { ep[6] = 255; ep[7] = 255; }
        if (ep[1] + ep[3] + ep[5] >= ep[0] + ep[2] + ep[4]) {
            for (int c = 0; c < 4; ++c) { e0[c] = ep[2 * c]; e1[c] = ep[2 * c + 1]; }
        } else {                                           // blue contraction (spec C.2.14)
            e0[0] = (ep[1] + ep[5]) >> 1; e0[1] = (ep[3] + ep[5]) >> 1; e0[2] = ep[5]; e0[3] = ep[7];
            e1[0] = (ep[0] + ep[4]) >> 1; e1[1] = (ep[2] + ep[4]) >> 1; e1[2] = ep[4]; e1[3] = ep[6];
        }
    } else {
        for (int i = 0; i < 16; ++i) wq[i] = 0;
        e0[0] = e1[0] = 255; e0[1] = e1[1] = 0; e0[2] = e1[2] = 255; e0[3] = e1[3] = 255;   // error colour
    }
    const int Ds = (1024 + dim / 2) / (dim - 1);
    for (int y = 0; y < dim; ++y) {
        const int py = int(by) * dim + y;
        if (py >= height) break;
        for (int x = 0; x < dim; ++x) {
            const int px = int(bx) * dim + x;
            if (px >= width) break;
            // infill of the 4x4 grid (spec C.2.18)
            const int gs = (Ds * x * 3 + 32) >> 6, gt = (Ds * y * 3 + 32) >> 6;
            const int js = gs >> 4, fs = gs & 15, jt = gt >> 4, ft = gt & 15;
            const int v0 = js + jt * 4;
            const int w11 = (fs * ft + 8) >> 4, w10 = ft - w11, w01 = fs - w11, w00 = 16 - fs - ft + w11;
            const int p00 = wq[v0], p01 = v0 + 1 < 16 ? wq[v0 + 1] : 0;
            const int p10 = v0 + 4 < 16 ? wq[v0 + 4] : 0, p11 = v0 + 5 < 16 ? wq[v0 + 5] : 0;
            const int w = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + 8) >> 4;
            uint32_t px4 = 0;
            for (int c = 0; c < 4; ++c) {
                const int c0 = (e0[c] << 8) | e0[c], c1 = (e1[c] << 8) | e1[c];
                const int v = (c0 * (64 - w) + c1 * w + 32) >> 6;
                px4 |= uint32_t(v >> 8) << (8 * c);
            }
            *(uint32_t *)(rgba + size_t(py) * pitch + size_t(px) * 4u) = px4;
        }
    }
}

// The same decoder, written for throughput (decode_kernel above stays as the any-alignment path and as the checker):
// everything in registers with compile-time indices -- the sixteen weights packed four to a register, an endpoint
// pair per channel packed in one register -- so that a texel costs
//     6x6 only: one PRMT (the four grid weights around it) + one IDP4A against the constant tap weights + a shift,
//     per channel: one IDP4A  e0 * (64 - w) + e1 * w,  one IMAD + shift  (257 t + 32) >> 14,
// and whole texel rows leave as 16-byte (4x4) or 8-byte (6x6) stores: a warp writes 512 / 768 contiguous bytes per row.
// The generic kernel wrote 4 bytes per thread with a 16 / 24-byte stride and kept its arrays in local memory: 1.07 TB/s
// of traffic on 16384^2; this one is bound by HBM.
struct DecodeTables {
    uint16_t trit[256];              // trits_from_integer(T): t0 | t1 << 2 | ... | t4 << 8
    uint8_t unq6[32], unq12[32];     // weight unquantisation of an encoded index (spec C.2.17)
};
constexpr DecodeTables make_decode_tables()
{
    DecodeTables d{};
    for (int T = 0; T < 256; ++T) {
        const Trits t = trits_from_integer(T);
        d.trit[T] = uint16_t(t.t[0] | (t.t[1] << 2) | (t.t[2] << 4) | (t.t[3] << 6) | (t.t[4] << 8));
    }
    const WeightTables w = make_weight_tables();
    for (int v = 0; v < 32; ++v) { d.unq6[v] = w.unq[QUANT_6][v]; d.unq12[v] = w.unq[QUANT_12][v]; }
    return d;
}
__device__ const DecodeTables g_decode_tables = make_decode_tables();

// bits [POS, POS + COUNT) of the 128-bit value (x = bits 0..31), compile-time position
template <int POS, int COUNT>
__device__ __forceinline__ uint32_t field128(const uint4 v)
{
    constexpr int i = POS >> 5, sh = POS & 31;
    const uint32_t w[5] = {v.x, v.y, v.z, v.w, 0u};
    const uint32_t lo = w[i], hi = w[i + 1];
    const uint32_t r = sh == 0 ? lo : __funnelshift_r(lo, hi, sh);
    return COUNT >= 32 ? r : r & ((1u << COUNT) - 1u);
}

// the sixteen unquantised weights of a block, packed four to a register (weight k in byte k & 3 of wq4[k >> 2])
template <int N>                                                   // plain bits per weight: 1 (QUANT_6) or 2 (QUANT_12)
__device__ __forceinline__ void decode_weights(const uint4 rev, const DecodeTables &st, uint32_t (&wq4)[4])
{
    constexpr int G = 5 * N + 8;
    constexpr uint32_t M = (1u << N) - 1u;
    const uint8_t *unq = N == 1 ? st.unq6 : st.unq12;
    uint32_t wq[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        // a group is at most 18 bits: read it from a 32-bit window
        const uint32_t grp = g == 0 ? field128<0 * G, G>(rev) : g == 1 ? field128<1 * G, G>(rev) : g == 2 ? field128<2 * G, G>(rev) : field128<3 * G, G>(rev);
        const uint32_t T = ((grp >> N) & 3u) | (((grp >> (2 * N + 2)) & 3u) << 2) | (((grp >> (3 * N + 4)) & 1u) << 4) |
                           (((grp >> (4 * N + 5)) & 3u) << 5) | (((grp >> (5 * N + 7)) & 1u) << 7);
        const uint32_t t = st.trit[T];
        const uint32_t m[5] = {grp & M, (grp >> (N + 2)) & M, (grp >> (2 * N + 4)) & M, (grp >> (3 * N + 5)) & M, (grp >> (4 * N + 7)) & M};
#pragma unroll
        for (int j = 0; j < 5; ++j)
            if (5 * g + j < 16) wq[5 * g + j] = unq[(((t >> (2 * j)) & 3u) << N) | m[j]];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) wq4[r] = wq[4 * r] | (wq[4 * r + 1] << 8) | (wq[4 * r + 2] << 16) | (wq[4 * r + 3] << 24);
}

// one texel: the four channels interpolated with weight w (0..64) between the packed endpoint pairs
__device__ __forceinline__ uint32_t decode_texel(const uint32_t (&e01)[4], uint32_t w)
{
    const uint32_t ww = w * 255u + 64u;                            // (64 - w) | w << 8
    uint32_t px = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t t = __dp4a(e01[c], ww, 0u);                 // e0 * (64 - w) + e1 * w
        px |= ((t * 257u + 32u) >> 14) << (8 * c);                 // ((c0 * (64 - w) + c1 * w + 32) >> 6) >> 8 with c = e * 257
    }
    return px;
}

template <int DIM>
__global__ void __launch_bounds__(128)
decode_fast_kernel(const uint4 *__restrict__ blocks, int width, int height, uint8_t *__restrict__ rgba, size_t pitch, uint32_t total,
                   uint32_t bw)
{
    __shared__ __align__(16) DecodeTables st;
    static_assert(sizeof(DecodeTables) % 16 == 0 && sizeof(DecodeTables) / 16 <= 128, "table copy");
    if (threadIdx.x < sizeof(DecodeTables) / 16)
        reinterpret_cast<uint4 *>(&st)[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(&g_decode_tables) + threadIdx.x);
    __syncthreads();
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    const uint4 b = __ldcs(blocks + id);
    const uint32_t mode = b.x & 0x7FFu, cem = (b.x >> 13) & 0xFu, parts = (b.x >> 11) & 3u;
    const uint32_t by = id / bw, bx = id - by * bw;
    const bool rgba_mode = mode == blockmode_4x4grid(QUANT_6) && cem == CEM_LDR_RGBA_DIRECT && parts == 0;
    const bool rgb_mode = mode == blockmode_4x4grid(QUANT_12) && cem == CEM_LDR_RGB_DIRECT && parts == 0;
    uint32_t wq4[4] = {0u, 0u, 0u, 0u};
    uint32_t e01[4] = {0xFFFFu, 0u, 0xFFFFu, 0xFFFFu};             // error colour (magenta) for anything this subset does not cover
    if (rgba_mode || rgb_mode) {
        const uint4 rev = make_uint4(__brev(b.w), __brev(b.z), __brev(b.y), __brev(b.x));   // the weight stream runs down from bit 127
        if (rgba_mode) decode_weights<1>(rev, st, wq4); else decode_weights<2>(rev, st, wq4);
        uint32_t ep[8];
        ep[0] = field128<17, 8>(b); ep[1] = field128<25, 8>(b); ep[2] = field128<33, 8>(b); ep[3] = field128<41, 8>(b);
        ep[4] = field128<49, 8>(b); ep[5] = field128<57, 8>(b); ep[6] = field128<65, 8>(b); ep[7] = field128<73, 8>(b);
        if (rgb_mode) { ep[6] = 255u; ep[7] = 255u; }
        if (ep[1] + ep[3] + ep[5] >= ep[0] + ep[2] + ep[4]) {
#pragma unroll
            for (int c = 0; c < 4; ++c) e01[c] = ep[2 * c] | (ep[2 * c + 1] << 8);
        } else {                                                   // blue contraction (spec C.2.14)
            e01[0] = ((ep[1] + ep[5]) >> 1) | (((ep[0] + ep[4]) >> 1) << 8);
            e01[1] = ((ep[3] + ep[5]) >> 1) | (((ep[2] + ep[4]) >> 1) << 8);
            e01[2] = ep[5] | (ep[4] << 8);
            e01[3] = ep[7] | (ep[6] << 8);
        }
    }
    const uint32_t x0 = bx * DIM, y0 = by * DIM;
    const bool whole = x0 + DIM <= uint32_t(width) && y0 + DIM <= uint32_t(height);
    uint8_t *dst = rgba + size_t(y0) * pitch + size_t(x0) * 4u;
    constexpr int Ds = (1024 + DIM / 2) / (DIM - 1);
#pragma unroll
    for (int y = 0; y < DIM; ++y) {
        uint32_t row[DIM];
#pragma unroll
        for (int x = 0; x < DIM; ++x) {
            uint32_t w;
            if (DIM == 4) {
                w = (wq4[y] >> (8 * x)) & 0xFFu;                    // the 4x4 grid IS the texel grid
            } else {
                // infill of the 4x4 grid (spec C.2.18); every constant below folds at compile time
                constexpr int dummy = 0; (void)dummy;
                const int gs = (Ds * x * 3 + 32) >> 6, gt = (Ds * y * 3 + 32) >> 6;
                const int js = gs >> 4, fs = gs & 15, jt = gt >> 4, ft = gt & 15;
                const int w11 = (fs * ft + 8) >> 4, w10 = ft - w11, w01 = fs - w11, w00 = 16 - fs - ft + w11;
                const uint32_t taps = uint32_t(w00) | (uint32_t(w01) << 8) | (uint32_t(w10) << 16) | (uint32_t(w11) << 24);
                // grid weights (js, jt), (js+1, jt), (js, jt+1), (js+1, jt+1): two packed rows, one byte permute; a tap beyond
                // the grid has weight 0, so whatever byte the wrapped selector picks does not matter
                const uint32_t sel = uint32_t(js) | (uint32_t((js + 1) & 3) << 4) | (uint32_t(4 + js) << 8) | (uint32_t(4 + ((js + 1) & 3)) << 12);
                const uint32_t p = __byte_perm(wq4[jt], jt + 1 < 4 ? wq4[jt + 1 < 4 ? jt + 1 : 3] : 0u, sel);
                w = __dp4a(p, taps, 8u) >> 4;
            }
            row[x] = decode_texel(e01, w);
        }
        uint8_t *r = dst + size_t(y) * pitch;
        if (whole) {
            if (DIM == 4) {
                *reinterpret_cast<uint4 *>(r) = make_uint4(row[0], row[1], row[2], row[3]);
            } else {
#pragma unroll
                for (int x = 0; x < DIM; x += 2) *reinterpret_cast<uint2 *>(r + 4 * x) = make_uint2(row[x], row[x + 1]);
            }
        } else if (y0 + uint32_t(y) < uint32_t(height)) {
#pragma unroll
            for (int x = 0; x < DIM; ++x)
                if (x0 + uint32_t(x) < uint32_t(width)) reinterpret_cast<uint32_t *>(r)[x] = row[x];
        }
    }
}

cudaError_t launch_decode(const uint8_t *d_blocks, int width, int height, int dim, uint8_t *d_rgba, size_t pitch,
                          cudaStream_t stream)
{
    const uint32_t bw = uint32_t((width + dim - 1) / dim), bh = uint32_t((height + dim - 1) / dim);
    const uint64_t total = uint64_t(bw) * bh;
    if (total == 0) return cudaSuccess;
    if (total > 0xFFFFFFFFull) return cudaErrorInvalidConfiguration;
    const unsigned ctas = unsigned((total + 127) / 128);
    const uintptr_t base = reinterpret_cast<uintptr_t>(d_rgba);
    if (dim == 4 && base % 16u == 0 && pitch % 16u == 0)
        decode_fast_kernel<4><<<ctas, 128, 0, stream>>>((const uint4 *)d_blocks, width, height, d_rgba, pitch, uint32_t(total), bw);
    else if (dim == 6 && base % 8u == 0 && pitch % 8u == 0)
        decode_fast_kernel<6><<<ctas, 128, 0, stream>>>((const uint4 *)d_blocks, width, height, d_rgba, pitch, uint32_t(total), bw);
    else
        decode_kernel<<<ctas, 128, 0, stream>>>((const uint4 *)d_blocks, width, height, dim, d_rgba, pitch, uint32_t(total), bw);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// 2x2 box-filter downsample of an RGBA8 image: the next level of a mip chain, produced on the
// device for the batch configuration (the reference has no mip generation, its
// caller would upload each level).  out(x, y) = (sum of the 2x2 source texels + 2) >> 2 per
// channel (round half up); an odd trailing row / column is dropped, a dimension of 1 stays 1.
// HBM-bound: 4 bytes read + 1 written per source texel.  Fast path: a thread makes four output
// texels from two rows of eight -- four 16-byte loads and one 16-byte store, a warp reads 2 x 1 KB
// and writes 512 B, all contiguous.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t box4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    constexpr uint32_t M = 0x00FF00FFu, R = 0x00020002u;      // two channels per word in 16-bit lanes
    const uint32_t rb = (a & M) + (b & M) + (c & M) + (d & M) + R;
    const uint32_t ga = ((a >> 8) & M) + ((b >> 8) & M) + ((c >> 8) & M) + ((d >> 8) & M) + R;
    return ((rb >> 2) & M) | (((ga >> 2) & M) << 8);
}

__global__ void __launch_bounds__(256)
downsample2x2_vec_kernel(const uint8_t *__restrict__ src, size_t src_pitch, uint8_t *__restrict__ dst, size_t dst_pitch,
                         uint32_t quads_x, uint64_t total_quads)
{
    for (uint64_t q = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < total_quads; q += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t y = uint32_t(q / quads_x), xq = uint32_t(q - uint64_t(y) * quads_x);
        const uint4 *r0 = reinterpret_cast<const uint4 *>(src + size_t(2u * y) * src_pitch) + 2u * xq;
        const uint4 *r1 = reinterpret_cast<const uint4 *>(src + size_t(2u * y + 1u) * src_pitch) + 2u * xq;
        const uint4 a0 = __ldcs(r0), a1 = __ldcs(r0 + 1), b0 = __ldcs(r1), b1 = __ldcs(r1 + 1);   // streamed: read once
        uint4 o;
        o.x = box4(a0.x, a0.y, b0.x, b0.y);
        o.y = box4(a0.z, a0.w, b0.z, b0.w);
        o.z = box4(a1.x, a1.y, b1.x, b1.y);
        o.w = box4(a1.z, a1.w, b1.z, b1.w);
        reinterpret_cast<uint4 *>(dst + size_t(y) * dst_pitch)[xq] = o;
    }
}

// any size / alignment: one thread per output texel
__global__ void __launch_bounds__(256)
downsample2x2_scalar_kernel(const uint8_t *__restrict__ src, size_t src_pitch, int width, int height, uint8_t *__restrict__ dst,
                            size_t dst_pitch, uint32_t out_w, uint64_t total)
{
    const uint32_t dx = width > 1 ? 1u : 0u, dy = height > 1 ? 1u : 0u;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t y = uint32_t(i / out_w), x = uint32_t(i - uint64_t(y) * out_w);
        const uint8_t *p0 = src + size_t(y << dy) * src_pitch + size_t(x << dx) * 4u;
        const uint8_t *p1 = p0 + size_t(dy) * src_pitch;
        const uint32_t a = *reinterpret_cast<const uint32_t *>(p0), b = *reinterpret_cast<const uint32_t *>(p0 + 4u * dx);
        const uint32_t c = *reinterpret_cast<const uint32_t *>(p1), d = *reinterpret_cast<const uint32_t *>(p1 + 4u * dx);
        *reinterpret_cast<uint32_t *>(dst + size_t(y) * dst_pitch + size_t(x) * 4u) = box4(a, b, c, d);
    }
}

cudaError_t launch_downsample2x2(const uint8_t *d_src, int width, int height, size_t src_pitch, uint8_t *d_dst, size_t dst_pitch,
                                 cudaStream_t stream)
{
    const uint32_t ow = uint32_t(width > 1 ? width / 2 : 1), oh = uint32_t(height > 1 ? height / 2 : 1);
    const uint64_t max_ctas = uint64_t(sm_count_of_current_device()) * 8u * 4u;                   // 8 resident CTAs of 256 threads per SM, four waves
    const bool vec = width > 1 && height > 1 && ow % 4u == 0 && src_pitch % 16u == 0 && dst_pitch % 16u == 0 &&
                     reinterpret_cast<uintptr_t>(d_src) % 16u == 0 && reinterpret_cast<uintptr_t>(d_dst) % 16u == 0;
    if (vec) {
        const uint64_t total = uint64_t(ow / 4u) * oh;
        const uint64_t ctas = (total + 255) / 256;
        downsample2x2_vec_kernel<<<unsigned(ctas < max_ctas ? ctas : max_ctas), 256, 0, stream>>>(d_src, src_pitch, d_dst, dst_pitch, ow / 4u, total);
    } else {
        const uint64_t total = uint64_t(ow) * oh;
        const uint64_t ctas = (total + 255) / 256;
        downsample2x2_scalar_kernel<<<unsigned(ctas < max_ctas ? ctas : max_ctas), 256, 0, stream>>>(d_src, src_pitch, width, height, d_dst,
                                                                                                   dst_pitch, ow, total);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// A whole mip chain in ONE launch (the "single pass downsampler" scheme): the level-by-level path above costs one
// launch per level, and below 256^2 a level is nothing but launch latency (2048^2 -> 1x1: eleven launches, ~30 us
// of GPU time issued from C, ~115 us from Python, against 3.6 us of HBM traffic).  Here a 256-thread CTA takes a
// 64x64 tile of the base and produces its 32x32 ... 1x1 reductions (levels 1-6): every thread reduces a 4x4 patch to
// 2x2 (level 1) and 1x1 (level 2) in registers, levels 3-6 go through shared memory.  The CTA that finishes last
// (a ticket in global memory) reduces the level-6 image -- one texel per tile -- down to 1x1, level by level.
// Every level is the 2x2 box filter of the ROUNDED level before it, exactly as in the level-by-level path, so the
// two produce the same bytes (tests/test_gpu_edges.py).  Requires width and height to be multiples of 64.
// ---------------------------------------------------------------------------
struct MipLevels {
    uint8_t *ptr[kMaxMipLevels];          // level l+1 of the chain: (max(1, w >> (l+1))) x (max(1, h >> (l+1))), rows tightly packed
    int32_t count;
};

__global__ void __launch_bounds__(256)
mip_chain_fused_kernel(const uint8_t *__restrict__ base, size_t base_pitch, int width, int height, MipLevels lv, unsigned *ticket)
{
    __shared__ uint32_t s_a[16 * 16], s_b[8 * 8];
    __shared__ bool s_last;
    const uint32_t tiles_x = uint32_t(width) / 64u;
    const uint32_t tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const uint32_t t = threadIdx.x, tx = t & 15u, ty = t >> 4;
    // ---- levels 1 and 2: a 4x4 patch of the base per thread (a warp reads 2 x 256 contiguous bytes per row) ----
    const uint8_t *src = base + size_t(tile_y * 64u + ty * 4u) * base_pitch + size_t(tile_x * 64u + tx * 4u) * 4u;
    const uint4 r0 = __ldcs(reinterpret_cast<const uint4 *>(src));
    const uint4 r1 = __ldcs(reinterpret_cast<const uint4 *>(src + base_pitch));
    const uint4 r2 = __ldcs(reinterpret_cast<const uint4 *>(src + 2 * base_pitch));
    const uint4 r3 = __ldcs(reinterpret_cast<const uint4 *>(src + 3 * base_pitch));
    const uint32_t a00 = box4(r0.x, r0.y, r1.x, r1.y), a01 = box4(r0.z, r0.w, r1.z, r1.w);
    const uint32_t a10 = box4(r2.x, r2.y, r3.x, r3.y), a11 = box4(r2.z, r2.w, r3.z, r3.w);
    {
        const uint32_t w1 = uint32_t(width) >> 1;
        uint8_t *d = lv.ptr[0] + (size_t(tile_y * 32u + ty * 2u) * w1 + (tile_x * 32u + tx * 2u)) * 4u;
        *reinterpret_cast<uint2 *>(d) = make_uint2(a00, a01);
        *reinterpret_cast<uint2 *>(d + size_t(w1) * 4u) = make_uint2(a10, a11);
    }
    const uint32_t l2 = box4(a00, a01, a10, a11);
    if (lv.count > 1) {
        const uint32_t w2 = uint32_t(width) >> 2;
        reinterpret_cast<uint32_t *>(lv.ptr[1])[size_t(tile_y * 16u + ty) * w2 + (tile_x * 16u + tx)] = l2;
    }
    s_a[ty * 16u + tx] = l2;
    __syncthreads();
    // ---- levels 3..6 of the tile through shared memory: 8x8, 4x4, 2x2, 1x1 ----
    uint32_t *cur = s_a, *nxt = s_b;
#pragma unroll
    for (int l = 3, n = 8; l <= 6; ++l, n >>= 1) {                 // n = side of level l inside the tile
        if (t < uint32_t(n * n)) {
            const uint32_t x = t % uint32_t(n), y = t / uint32_t(n), m = uint32_t(2 * n);
            const uint32_t v = box4(cur[(2u * y) * m + 2u * x], cur[(2u * y) * m + 2u * x + 1u], cur[(2u * y + 1u) * m + 2u * x],
                                    cur[(2u * y + 1u) * m + 2u * x + 1u]);
            nxt[y * uint32_t(n) + x] = v;
            if (l <= lv.count) {
                const uint32_t wl = uint32_t(width) >> l;
                reinterpret_cast<uint32_t *>(lv.ptr[l - 1])[size_t(tile_y * uint32_t(n) + y) * wl + (tile_x * uint32_t(n) + x)] = v;
            }
        }
        __syncthreads();
        uint32_t *tmp = cur; cur = nxt; nxt = tmp;
    }
    if (lv.count <= 6) return;
    // ---- the last CTA to get here reduces the level-6 image (one texel per tile) to 1x1 ----
    __threadfence();                                               // this CTA's level-6 texel is visible before its ticket
    if (t == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    uint32_t w = uint32_t(width) >> 6, h = uint32_t(height) >> 6;
    for (int l = 7; l <= lv.count; ++l) {
        const uint32_t *in = reinterpret_cast<const uint32_t *>(lv.ptr[l - 2]);
        uint32_t *out = reinterpret_cast<uint32_t *>(lv.ptr[l - 1]);
        const uint32_t dx = w > 1u ? 1u : 0u, dy = h > 1u ? 1u : 0u;
        const uint32_t ow = w > 1u ? w >> 1 : 1u, oh = h > 1u ? h >> 1 : 1u;
        for (uint32_t i = t; i < ow * oh; i += 256u) {
            const uint32_t y = i / ow, x = i - y * ow;
            const uint32_t *p0 = in + size_t(y << dy) * w + (x << dx), *p1 = p0 + size_t(dy) * w;
            out[i] = box4(__ldcg(p0), __ldcg(p0 + dx), __ldcg(p1), __ldcg(p1 + dx));     // L2, never a stale L1 line
        }
        __syncthreads();                                           // the level just written is read by other threads next
        w = ow;
        h = oh;
    }
    if (t == 0) *ticket = 0u;                                      // ready for the next launch on this stream
}

// Levels 1.. of the chain of a width x height image into `levels` (layout: mip_chain_layout in astc_capi.cu).
cudaError_t launch_mip_chain(const uint8_t *d_base, int width, int height, size_t base_pitch, uint8_t *const *level_ptrs,
                             const int *level_w, const int *level_h, int count, unsigned *d_ticket, cudaStream_t stream)
{
    if (count <= 0) return cudaSuccess;
    const bool fused = width % 64 == 0 && height % 64 == 0 && base_pitch % 16u == 0 && reinterpret_cast<uintptr_t>(d_base) % 16u == 0 &&
                       count <= kMaxMipLevels && d_ticket != nullptr;
    if (fused) {
        MipLevels lv{};
        lv.count = count;
        for (int l = 0; l < count; ++l) lv.ptr[l] = level_ptrs[l];
        const unsigned ctas = unsigned(width / 64) * unsigned(height / 64);
        mip_chain_fused_kernel<<<ctas, 256, 0, stream>>>(d_base, base_pitch, width, height, lv, d_ticket);
        return cudaGetLastError();
    }
    // any size: one launch per level, issued back to back
    const uint8_t *src = d_base;
    size_t pitch = base_pitch;
    int w = width, h = height;
    for (int l = 0; l < count; ++l) {
        const cudaError_t e = launch_downsample2x2(src, w, h, pitch, level_ptrs[l], size_t(level_w[l]) * 4u, stream);
        if (e != cudaSuccess) return e;
        src = level_ptrs[l];
        w = level_w[l];
        h = level_h[l];
        pitch = size_t(w) * 4u;
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------------------
// The two hardware approximations the encoder's arithmetic is defined on (MUFU.RCP, MUFU.RSQ),
// applied to an array: lets tests and tools/gen_mufu_tables.py compare the oracle's table-driven
// emulation with the device, value by value.
// ---------------------------------------------------------------------------
__global__ void mufu_kernel(int op, const float *__restrict__ x, float *__restrict__ y, size_t n)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        float r;
        if (op == 0) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i]));
        else asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i]));
        y[i] = r;
    }
}

cudaError_t launch_mufu(int op, const float *d_x, float *d_y, size_t n, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    const size_t ctas = (n + 255) / 256;
    mufu_kernel<<<unsigned(ctas < 148u * 32u ? ctas : 148u * 32u), 256, 0, stream>>>(op, d_x, d_y, n);
    return cudaGetLastError();
}

}  // namespace astc
