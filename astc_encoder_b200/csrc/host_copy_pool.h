// host_copy_pool.h -- a few persistent host threads that copy rows between pageable and pinned memory for the
// staged pipeline of astc_b200_context_encode_host (astc_context.cu).  Plain C++17, no CUDA: unit-tested on the CPU
// by tests/test_host.py (tests/cpp/copy_pool_test.cpp, also run under ThreadSanitizer during development).
//
// Design.  A job is a 2-D copy cut into pieces of kCopyGrain bytes; pieces are claimed from one atomic word that
// carries the job's generation in its upper half, so a worker that wakes up late can never take a piece of a LATER
// job with an earlier job's descriptor.  The calling thread claims pieces like any worker and returns as soon as
// all pieces of ITS job are done -- it never waits for a sleeping worker to check in (waking a halted vCPU takes
// 100+ us on a VM guest: measured 733 us for a 4 MiB copy when the caller waited for the workers, against ~400 us
// for copying alone).  Workers spin for a short while after a job before they block, so the bands of one texture and
// back-to-back calls find them awake.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace astc_host {

constexpr size_t kCopyGrain = 256u << 10;         // bytes a copy worker takes at a time
constexpr int kSpinMicros = 150;                  // how long an idle worker keeps polling before it blocks

// memcpy whose stores bypass the cache (the destination is a pinned staging slot that only the DMA engine reads):
// no read-for-ownership of the destination lines, a third less memory traffic than a cached copy.
inline void stream_copy(uint8_t *dst, const uint8_t *src, size_t n)
{
#if defined(__SSE2__)
    if (n >= 256) {
        const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
        std::memcpy(dst, src, head);
        dst += head; src += head; n -= head;
        const size_t body = n & ~size_t(63);
        for (size_t i = 0; i < body; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
        }
        std::memcpy(dst + body, src + body, n - body);
        return;
    }
#endif
    std::memcpy(dst, src, n);
}

class CopyPool {
public:
    CopyPool() = default;
    CopyPool(const CopyPool &) = delete;
    CopyPool &operator=(const CopyPool &) = delete;
    ~CopyPool() { stop_workers(); }

    // Worker threads besides the caller: -1 = automatic (a quarter of the host's hardware threads, 1..7), 0 = none.
    // Takes effect at the next copy; call between copies only.
    void set_workers(int n)
    {
        stop_workers();
        wanted_ = n;
        tried_ = false;
    }
    int workers() const { return int(threads_.size()); }

    // dst / src: `rows` rows of `row_bytes` bytes, `*_pitch` apart.  `streaming`: non-temporal stores.
    void copy_rows(uint8_t *dst, size_t dst_pitch, const uint8_t *src, size_t src_pitch, size_t row_bytes, size_t rows,
                   bool streaming = false)
    {
        if (rows == 0 || row_bytes == 0) return;
        if (dst_pitch == row_bytes && src_pitch == row_bytes) {          // contiguous on both sides: one long row
            row_bytes *= rows;
            dst_pitch = src_pitch = row_bytes;
            rows = 1;
        }
        Job j{dst, src, dst_pitch, src_pitch, row_bytes, rows, (row_bytes * rows + kCopyGrain - 1) / kCopyGrain, streaming};
        if (j.pieces < 2 || !start()) {
            for (size_t i = 0; i < j.pieces; ++i) copy_piece(j, i);
            finish(j);
            return;
        }
        uint64_t gen;
        {
            std::lock_guard<std::mutex> l(m_);
            gen = ++generation_;
            job_ = j;
            done_.store(0, std::memory_order_relaxed);
            claim_.store(gen << 32, std::memory_order_release);
        }
        wake_.notify_all();
        drain(j, gen);
        while (done_.load(std::memory_order_acquire) < j.pieces) cpu_relax();      // pieces still in a worker's hands
        finish(j);
    }

private:
    struct Job {
        uint8_t *dst;
        const uint8_t *src;
        size_t dst_pitch, src_pitch, row_bytes, rows, pieces;
        bool streaming;
    };

    static void cpu_relax()
    {
#if defined(__SSE2__)
        _mm_pause();
#else
        std::this_thread::yield();
#endif
    }
    static void finish(const Job &j)
    {
#if defined(__SSE2__)
        if (j.streaming) _mm_sfence();                                   // the DMA that follows must see the streamed stores
#else
        (void)j;
#endif
    }
    // piece i = bytes [i * kCopyGrain, (i + 1) * kCopyGrain) of the rows laid end to end
    static void copy_piece(const Job &j, size_t i)
    {
        const size_t total = j.row_bytes * j.rows;
        size_t pos = i * kCopyGrain;
        const size_t end = std::min(total, pos + kCopyGrain);
        while (pos < end) {
            const size_t y = pos / j.row_bytes, x = pos - y * j.row_bytes;
            const size_t n = std::min(end - pos, j.row_bytes - x);
            if (j.streaming) stream_copy(j.dst + y * j.dst_pitch + x, j.src + y * j.src_pitch + x, n);
            else std::memcpy(j.dst + y * j.dst_pitch + x, j.src + y * j.src_pitch + x, n);
            pos += n;
        }
    }
    // claim pieces of generation `gen` until none is left or the pool has moved on to another job
    void drain(const Job &j, uint64_t gen)
    {
        for (;;) {
            uint64_t v = claim_.load(std::memory_order_acquire);
            for (;;) {
                if ((v >> 32) != (gen & 0xFFFFFFFFu) || (v & 0xFFFFFFFFu) >= j.pieces) return;
                if (claim_.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel)) break;
            }
            copy_piece(j, size_t(v & 0xFFFFFFFFu));
            if (j.streaming) finish(j);
            done_.fetch_add(1, std::memory_order_release);
        }
    }
    bool start()
    {
        if (!threads_.empty()) return true;
        if (tried_) return false;
        tried_ = true;
        int n = wanted_;
        if (n < 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            n = int(std::min(7u, std::max(1u, hw / 4u)));
            if (hw < 2) n = 0;
        }
        try {
            for (int i = 0; i < n; ++i) threads_.emplace_back([this] { worker(); });
        } catch (...) {
            // fewer (or no) threads: the caller copies the rest itself
        }
        return !threads_.empty();
    }
    void stop_workers()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
        }
        wake_.notify_all();
        for (auto &t : threads_) t.join();
        threads_.clear();
        stop_ = false;
    }
    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            Job j;
            uint64_t gen;
            {
                std::unique_lock<std::mutex> l(m_);
                if (!stop_ && generation_ == seen) {
                    // poll for a while first: the next band / the next call usually follows within microseconds
                    l.unlock();
                    const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(kSpinMicros);
                    while ((claim_.load(std::memory_order_acquire) >> 32) == (seen & 0xFFFFFFFFu) &&
                           std::chrono::steady_clock::now() < until)
                        cpu_relax();
                    l.lock();
                    wake_.wait(l, [&] { return stop_ || generation_ != seen; });
                }
                if (stop_) return;
                seen = gen = generation_;
                j = job_;
            }
            drain(j, gen);
        }
    }

    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable wake_;
    Job job_{};
    std::atomic<uint64_t> claim_{0};              // (generation & 0xFFFFFFFF) << 32 | next piece
    std::atomic<size_t> done_{0};                 // pieces of the current job that are finished
    uint64_t generation_ = 0;
    int wanted_ = -1;
    bool stop_ = false, tried_ = false;
};

}  // namespace astc_host
