// host_unpack.cpp — expands 4-bit packed observation records on the host cores (SSE2: baseline x86-64), in parallel
// with the DMA of the part of the batch that crosses PCIe unpacked.  See host_api.cu.
#include "host_unpack.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace dmfb {

struct UnpackPool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_start;
    std::atomic<uint64_t> generation{0};
    std::atomic<int> pending{0};
    std::atomic<bool> stop{false};
    int n_threads = 1;
    // current job
    const uint8_t* in = nullptr;
    size_t stride = 0;
    int8_t* out = nullptr;
    int cells = 0;
    size_t n = 0;
};

namespace {

inline void unpack_one(const uint8_t* in, int8_t* out, int cells)
{
    int k = 0;
#if defined(__SSE2__)
    const __m128i low = _mm_set1_epi8(0x0F);
    for (; k + 32 <= cells; k += 32) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(in + (k >> 1)));
        const __m128i lo = _mm_and_si128(v, low), hi = _mm_and_si128(_mm_srli_epi16(v, 4), low);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(out + k), _mm_unpacklo_epi8(lo, hi));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(out + k + 16), _mm_unpackhi_epi8(lo, hi));
    }
#endif
    for (; k < cells; ++k) out[k] = (int8_t)((in[k >> 1] >> ((k & 1) * 4)) & 15);
    const int nb = (cells + 1) / 2;
    out[cells] = (int8_t)in[nb];
    out[cells + 1] = (int8_t)in[nb + 1];
}

// One thread's share of the records.  The destination is written exactly once and never read, so it is streamed
// with non-temporal stores (no read-for-ownership: half the DRAM traffic of ordinary stores): records are expanded
// into a small cache-resident buffer that keeps the destination's 16-byte phase, whole 16-byte blocks leave with
// MOVNTDQ, and the < 16 trailing bytes are carried over to the next group.
void run_slice(const UnpackPool* p, int t)
{
    const size_t r0 = p->n * (size_t)t / (size_t)p->n_threads, r1 = p->n * (size_t)(t + 1) / (size_t)p->n_threads;
    const size_t D = (size_t)p->cells + 2;
#if defined(__SSE2__)
    constexpr size_t kGroupBytes = 4096;
    const size_t group = kGroupBytes / D > 0 ? kGroupBytes / D : 1;       // records per flush
    static thread_local std::vector<int8_t> storage;
    if (storage.size() < 64 + 16 + group * D) storage.resize(64 + 16 + group * D);
    int8_t* const buf16 = reinterpret_cast<int8_t*>((reinterpret_cast<uintptr_t>(storage.data()) + 63) & ~(uintptr_t)63);
    int8_t* dst = p->out + r0 * D;
    int8_t* base = buf16 + (reinterpret_cast<uintptr_t>(dst) & 15);      // same 16-byte phase as dst
    size_t fill = 0;
    for (size_t r = r0; r < r1;) {
        const size_t g1 = r + group < r1 ? r + group : r1;
        for (; r < g1; ++r, fill += D) unpack_one(p->in + r * p->stride, base + fill, p->cells);
        size_t n = fill;
        const size_t mis = reinterpret_cast<uintptr_t>(dst) & 15;
        if (mis) {                                                          // only before the first flush
            const size_t h = n < 16 - mis ? n : 16 - mis;
            memcpy(dst, base, h);
            dst += h; base += h; n -= h;
        }
        const size_t blocks = n >> 4;
        for (size_t k = 0; k < blocks; ++k)
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + k, _mm_load_si128(reinterpret_cast<const __m128i*>(base) + k));
        dst += blocks << 4;
        const size_t rem = n & 15;
        if (r < r1) {
            memmove(buf16, base + (blocks << 4), rem);                      // dst is 16-byte aligned from here on
            base = buf16;
            fill = rem;
        } else {
            memcpy(dst, base + (blocks << 4), rem);
        }
    }
    _mm_sfence();
#else
    for (size_t r = r0; r < r1; ++r) unpack_one(p->in + r * p->stride, p->out + r * D, p->cells);
#endif
}

inline void cpu_relax()
{
#if defined(__SSE2__)
    _mm_pause();
#endif
}

// A job arrives every few hundred microseconds while an env is being stepped, and lasts about as long: a condition
// variable wake-up (tens of microseconds per thread) would eat the gain, so workers spin for a while after a job
// before they go to sleep.
constexpr auto kSpinFor = std::chrono::milliseconds(3);

void worker_main(UnpackPool* p, int t)
{
    uint64_t seen = 0;
    for (;;) {
        const auto t0 = std::chrono::steady_clock::now();
        unsigned polls = 0;
        while (p->generation.load(std::memory_order_acquire) == seen && !p->stop.load(std::memory_order_relaxed)) {
            cpu_relax();
            if ((++polls & 1023u) == 0 && std::chrono::steady_clock::now() - t0 > kSpinFor) {
                std::unique_lock<std::mutex> lk(p->m);
                p->cv_start.wait(lk, [&] { return p->stop.load() || p->generation.load() != seen; });
            }
        }
        if (p->stop.load()) return;
        seen = p->generation.load(std::memory_order_acquire);
        run_slice(p, t);
        p->pending.fetch_sub(1, std::memory_order_release);
    }
}

}  // namespace

UnpackPool* unpack_pool_create(int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    UnpackPool* p = new UnpackPool();
    p->n_threads = n_threads;
    for (int t = 1; t < n_threads; ++t) p->workers.emplace_back(worker_main, p, t);
    return p;
}

void unpack_pool_destroy(UnpackPool* p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->stop.store(true);
    }
    p->cv_start.notify_all();
    for (std::thread& w : p->workers) w.join();
    delete p;
}

void unpack_records(UnpackPool* p, const uint8_t* packed, size_t packed_stride, int8_t* out, int cells, size_t n_records)
{
    if (n_records == 0) return;
    {
        std::lock_guard<std::mutex> lk(p->m);    // orders the job fields before the generation bump for sleepers
        p->in = packed; p->stride = packed_stride; p->out = out; p->cells = cells; p->n = n_records;
        p->pending.store(p->n_threads - 1, std::memory_order_relaxed);
        p->generation.fetch_add(1, std::memory_order_release);
    }
    p->cv_start.notify_all();
    run_slice(p, 0);
    while (p->pending.load(std::memory_order_acquire) != 0) cpu_relax();
}

}  // namespace dmfb
