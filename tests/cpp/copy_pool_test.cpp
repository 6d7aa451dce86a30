// CPU unit test of astc_host::CopyPool (csrc/host_copy_pool.h): every shape the staged pipeline hands it --
// contiguous blobs, pitched rows on either side, rows shorter and longer than a worker's grain, tiny jobs that stay
// on the caller -- must copy exactly the bytes memcpy row by row would, and nothing outside them.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "host_copy_pool.h"

static int check(astc_host::CopyPool &pool, size_t row_bytes, size_t rows, size_t src_pad, size_t dst_pad, unsigned seed, bool streaming = false)
{
    const size_t sp = row_bytes + src_pad, dp = row_bytes + dst_pad;
    std::vector<uint8_t> src(sp * rows + 64), dst(dp * rows + 64, 0xA5), want(dp * rows + 64, 0xA5);
    std::mt19937 rng(seed);
    for (auto &b : src) b = uint8_t(rng());
    for (size_t y = 0; y < rows; ++y) std::memcpy(want.data() + (seed % 3) + y * dp, src.data() + (seed % 5) + y * sp, row_bytes);
    pool.copy_rows(dst.data() + (seed % 3), dp, src.data() + (seed % 5), sp, row_bytes, rows, streaming);
    if (dst != want) {
        std::printf("FAIL row_bytes=%zu rows=%zu src_pad=%zu dst_pad=%zu\n", row_bytes, rows, src_pad, dst_pad);
        return 1;
    }
    return 0;
}

int main()
{
    astc_host::CopyPool pool;
    int bad = 0;
    unsigned seed = 1;
    const size_t G = astc_host::kCopyGrain;
    const size_t shapes[][4] = {
        {1, 1, 0, 0}, {17, 3, 5, 0}, {4096, 64, 0, 0}, {4096, 1024, 0, 0}, {4096, 1024, 128, 0}, {4096, 1023, 0, 112},
        {65536, 64, 0, 0}, {65532, 67, 4, 12}, {G, 8, 0, 0}, {G + 1, 7, 3, 0}, {G - 1, 9, 0, 1}, {3 * G + 5, 3, 0, 0},
        {3 * G + 5, 3, 7, 9}, {1000 * 4, 2100, 96, 0}, {8 * G, 1, 0, 0}, {2 * G - 1, 1, 0, 0}, {2 * G, 1, 0, 0},
    };
    for (int rep = 0; rep < 4; ++rep) {                     // the pool is reused: generations must not leak into each other
        if (rep == 2) pool.set_workers(0);                  // caller only
        if (rep == 3) pool.set_workers(5);
        for (const auto &s : shapes) bad += check(pool, s[0], s[1], s[2], s[3], seed++, (seed & 1) != 0);
    }
    // many small jobs back to back: late-waking workers must never touch a later job's pieces
    for (int i = 0; i < 500; ++i) bad += check(pool, 2 * G + 64 * size_t(i % 7), 1 + size_t(i % 3), size_t(i % 2) * 16, 0, seed++, (i & 1) != 0);
    pool.copy_rows(nullptr, 0, nullptr, 0, 0, 5);           // empty jobs are no-ops
    pool.copy_rows(nullptr, 0, nullptr, 0, 5, 0);
    std::printf(bad ? "copy_pool_test: %d failures\n" : "copy_pool_test: ok\n", bad);
    return bad ? 1 : 0;
}
