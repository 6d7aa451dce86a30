// CPU unit test of astc::plan_tapered (csrc/astc_schedule.h): whatever the job size, pass count and number of resident
// CTAs, the CTAs of a launch must cover the block ids [0, total) exactly once, in ascending order, with non-increasing
// pass counts, and a partial run only at the very end.
#include <cstdio>
#include <cstdlib>

#include "astc_schedule.h"

using namespace astc;

static int check(uint64_t total, int threads, int passes, uint64_t resident, bool taper)
{
    EncodeParams p{};
    p.total_blocks = total;
    p.passes = passes;
    const uint64_t ctas = plan_tapered(p, threads, resident, taper);
    uint64_t next = 0;
    int prev = 1 << 30;
    for (uint64_t c = 0; c < ctas; ++c) {
        int q = 0;
        const uint64_t first = cta_first_block(p, threads, c, q);
        if (first != next || q < 1 || q > prev || first >= total) {
            std::printf("FAIL total=%llu threads=%d passes=%d resident=%llu taper=%d: cta %llu first %llu (expected %llu) passes %d (prev %d)\n",
                        (unsigned long long)total, threads, passes, (unsigned long long)resident, int(taper), (unsigned long long)c,
                        (unsigned long long)first, (unsigned long long)next, q, prev);
            return 1;
        }
        prev = q;
        next = first + uint64_t(q) * uint64_t(threads);
    }
    if (next < total || (ctas && next - total >= uint64_t(prev) * uint64_t(threads))) {
        std::printf("FAIL coverage total=%llu threads=%d passes=%d resident=%llu: ids up to %llu\n", (unsigned long long)total, threads, passes,
                    (unsigned long long)resident, (unsigned long long)next);
        return 1;
    }
    if (taper && p.nseg) {
        if (p.seg[p.nseg - 1].passes != 1 && p.nseg < kMaxSegments) { std::printf("FAIL: taper does not end at 1 pass\n"); return 1; }
        for (int i = 1; i < p.nseg; ++i)
            if (p.seg[i].cta_begin != p.seg[i - 1].cta_end) { std::printf("FAIL: segments not contiguous\n"); return 1; }
    }
    return 0;
}

int main()
{
    int bad = 0, tapered = 0;
    const uint64_t totals[] = {1, 127, 128, 129, 4096, 65536, 1048576, 2097152, 1865956, 16777216, 178957824, 75776 * 8 + 3, 592ull * 15 * 128};
    for (uint64_t total : totals)
        for (int threads : {128, 256})
            for (int passes = 1; passes <= 8; ++passes)
                for (uint64_t resident : {1ull, 148ull * 4, 148ull * 6, 148ull * 7, 132ull * 4})
                    for (bool taper : {false, true}) {
                        bad += check(total, threads, passes, resident, taper);
                        EncodeParams p{};
                        p.total_blocks = total; p.passes = passes;
                        plan_tapered(p, threads, resident, taper);
                        tapered += p.nseg > 0;
                    }
    // the shapes the benchmark runs: 4096^2 (4 passes) and a 1/8 band of 16384^2 (8 passes) do get a taper
    EncodeParams p{};
    p.total_blocks = 1048576; p.passes = 4;
    const uint64_t c4k = plan_tapered(p, 128, 592, true);
    if (p.nseg != 3 || c4k != 1604 + 592 + 592) { std::printf("FAIL: 4096^2 schedule nseg %d ctas %llu\n", p.nseg, (unsigned long long)c4k); ++bad; }
    p = EncodeParams{};
    p.total_blocks = 2097152; p.passes = 8;
    plan_tapered(p, 128, 592, true);
    if (p.nseg != 4 || p.seg[3].passes != 1) { std::printf("FAIL: band schedule\n"); ++bad; }
    std::printf(bad ? "schedule_test: %d failures\n" : "schedule_test: ok (%d tapered plans)\n", bad ? bad : tapered);
    return bad ? 1 : 0;
}
