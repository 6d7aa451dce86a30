// astc_schedule.h -- which block ids each CTA of an encode launch takes (host side; plain C++, unit-tested on the
// CPU by tests/cpp/schedule_test.cpp).  Replaces the Dispatch geometry of astc_encode.h:124-134.
#pragma once
#include <cstdint>

#include "astc_kernels.h"

namespace astc {

// Tapered schedule: CTAs of `p.passes` passes, then one wave of `resident` CTAs each of passes/2, passes/4, ... 1 --
// the hardware hands CTAs out in index order, so the short ones run last and the SMs run out of work within one
// 1-pass CTA of each other instead of one `passes`-pass CTA.  Fills p.seg / p.nseg and returns the CTA count;
// p.nseg stays 0 (every CTA runs p.passes passes) when tapering is off or the job is too small for it.
inline uint64_t plan_tapered(EncodeParams &p, int threads, uint64_t resident, bool taper)
{
    const uint64_t per_cta = uint64_t(threads) * uint64_t(p.passes);
    const uint64_t uniform = (p.total_blocks + per_cta - 1) / per_cta;
    p.nseg = 0;
    if (!taper || p.passes < 2) return uniform;
    const uint64_t runs = (p.total_blocks + uint64_t(threads) - 1) / uint64_t(threads);      // runs of `threads` consecutive ids
    uint64_t tail_runs = 0;
    for (int q = p.passes >> 1; q >= 1; q >>= 1) tail_runs += resident * uint64_t(q);
    if (runs < tail_runs + resident * uint64_t(p.passes)) return uniform;                    // not even one full wave of long CTAs
    const uint64_t main_ctas = (runs - tail_runs) / uint64_t(p.passes);
    int n = 0;
    p.seg[n++] = Segment{0, 0, uint32_t(main_ctas), uint32_t(p.passes), 0};
    uint64_t cta = main_ctas, run = main_ctas * uint64_t(p.passes);
    for (int q = p.passes >> 1; q >= 1 && n < kMaxSegments; q >>= 1) {
        const bool last = q == 1 || n == kMaxSegments - 1;
        const uint64_t left = runs - run;
        const uint64_t c = last ? (left + uint64_t(q) - 1) / uint64_t(q) : resident;       // the last segment takes what is left
        p.seg[n++] = Segment{run * uint64_t(threads), uint32_t(cta), uint32_t(cta + c), uint32_t(q), 0};
        cta += c;
        run += c * uint64_t(q);
        if (last) break;
    }
    p.nseg = n;
    return cta;
}

// The device side of the same mapping (cta_schedule in astc_kernels.cu), restated for the host: first block id of
// CTA `cta` and its pass count.
inline uint64_t cta_first_block(const EncodeParams &p, int threads, uint64_t cta, int &passes)
{
    if (p.nseg == 0) {
        passes = p.passes;
        return cta * uint64_t(p.passes) * uint64_t(threads);
    }
    uint32_t q = p.seg[0].passes, begin = 0;
    uint64_t first = 0;
    for (int i = 1; i < kMaxSegments; ++i)
        if (i < p.nseg && cta >= p.seg[i].cta_begin) { q = p.seg[i].passes; begin = p.seg[i].cta_begin; first = p.seg[i].first_block; }
    passes = int(q);
    return first + (cta - begin) * uint64_t(q) * uint64_t(threads);
}

}  // namespace astc
