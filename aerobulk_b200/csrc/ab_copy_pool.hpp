// ab_copy_pool.hpp -- persistent host threads that copy a list of memory pieces (used by ab_api.cu for PAGEABLE caller
// arrays: caller arrays <-> the library's pinned slab).  Plain C++, no CUDA: tests/ builds it with -fsanitize=thread.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace abpool
{
// memcpy with streaming (non-temporal) stores: the destination is not read into the cache first (no read-for-ownership),
// which saves a third of the memory traffic of a large copy whose destination is not reused by this core
inline void copy_streaming(void *dst, const void *src, size_t bytes)
{
#if defined(__x86_64__)
    char *d = static_cast<char *>(dst);
    const char *s = static_cast<const char *>(src);
    const size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;
    if (bytes < 256 + head) {
        memcpy(d, s, bytes);
        return;
    }
    memcpy(d, s, head);
    d += head; s += head; bytes -= head;
    const size_t lines = bytes / 64;
    for (size_t i = 0; i < lines; ++i) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(d), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 48), e);
        s += 64; d += 64;
    }
    _mm_sfence();
    memcpy(d, s, bytes - lines * 64);
#else
    memcpy(dst, src, bytes);
#endif
}

struct CopyPiece {
    void *dst;
    const void *src;
    size_t bytes;
};
class CopyPool
{
  public:
    explicit CopyPool(int nthreads, bool streaming = false) : streaming_(streaming)
    {
        for (int t = 1; t < nthreads; ++t) std::thread([this] { work(); }).detach();   // the caller is thread 0
    }
    // copies every piece; returns when all of them are done
    void run(const CopyPiece *pieces, int np)
    {
        if (np <= 0) return;
        const unsigned long long gen = (ticket_.load(std::memory_order_relaxed) >> 32) + 1;
        // slot gen&1 was last used by generation gen-2: a thread still holding one of its (exhausted) tickets entered
        // drain() before generation gen-1 was published, so it is gone once the in-drain count has been seen at zero
        while (in_drain_.load(std::memory_order_seq_cst) != 0) cpu_relax();
        Job &j = job_[gen & 1];
        j.pieces = pieces;
        j.np = np;
        j.done.store(0, std::memory_order_relaxed);
        ticket_.store(gen << 32, std::memory_order_seq_cst);
        if (sleepers_.load(std::memory_order_acquire) > 0) {
            { std::lock_guard<std::mutex> lk(mu_); }
            cv_.notify_all();
        }
        drain();
        while (j.done.load(std::memory_order_acquire) < np) cpu_relax();
    }

  private:
    struct Job {
        const CopyPiece *pieces = nullptr;
        int np = 0;
        std::atomic<int> done{0};
    };
    static void cpu_relax()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    // take tickets of the current generation until none is left; returns that generation.  A ticket carries its
    // generation, so a late thread can never apply an old index to a new job.
    unsigned long long drain()
    {
        in_drain_.fetch_add(1, std::memory_order_seq_cst);
        unsigned long long gen;
        for (;;) {
            const unsigned long long v = ticket_.fetch_add(1, std::memory_order_seq_cst);
            gen = v >> 32;
            const unsigned idx = (unsigned)(v & 0xffffffffu);
            Job &j = job_[gen & 1];
            if (gen == 0 || idx >= (unsigned)j.np) break;
            const CopyPiece &c = j.pieces[idx];
            if (streaming_) copy_streaming(c.dst, c.src, c.bytes);
            else memcpy(c.dst, c.src, c.bytes);
            j.done.fetch_add(1, std::memory_order_release);
        }
        in_drain_.fetch_sub(1, std::memory_order_seq_cst);
        return gen;
    }
    void work()
    {
        for (;;) {
            const unsigned long long seen = drain();
            // spin briefly (back-to-back chunks of one call), then sleep until the next job
            bool woke = false;
            for (int spin = 0; spin < 40000; ++spin) {
                if ((ticket_.load(std::memory_order_acquire) >> 32) != seen) { woke = true; break; }
                cpu_relax();
            }
            if (woke) continue;
            std::unique_lock<std::mutex> lk(mu_);
            sleepers_.fetch_add(1, std::memory_order_acq_rel);
            cv_.wait(lk, [&] { return (ticket_.load(std::memory_order_acquire) >> 32) != seen; });
            sleepers_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    std::atomic<unsigned long long> ticket_{0};   // generation << 32 | next piece index
    Job job_[2];
    const bool streaming_;
    std::atomic<int> in_drain_{0};   // threads between taking a ticket and having finished with its job slot
    std::atomic<int> sleepers_{0};
    std::mutex mu_;
    std::condition_variable cv_;
};
}  // namespace abpool
