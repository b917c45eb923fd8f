// Test harness (not product code): drives the window batcher of libsnacb.so the way a serving process would --
// several PRODUCER THREADS push the token streams of many concurrent requests a few tokens at a time
// (vllm_inference/modal_audio_stream.py:352-372 does this one token at a time from the asyncio loop) while the calling
// thread flushes every ready window through the pipelined flush -- and reports windows/s.  Called through ctypes by
// tests/gpu_batcher_bench.py; built by tts_inference_b200/build.py into tests/native/libbatcherbench.so.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/snacb.h"

namespace {
inline uint32_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return static_cast<uint32_t>(x ^ (x >> 31));
}
}  // namespace

extern "C" int batcher_bench(snacb_batcher b, int n_streams, int tokens_per_stream, int n_threads, int push_tokens,
                             int max_windows, int16_t* pcm0, int16_t* pcm1, size_t pcm_cap, double* out) {
    if (!b || n_streams <= 0 || tokens_per_stream <= 0 || n_threads <= 0 || push_tokens <= 0 || !pcm0 || !pcm1 || !out) return -1;
    std::atomic<int> producers_left{n_threads};
    std::atomic<long long> pushes{0};
    std::atomic<int> failed{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int k = 0; k < n_threads; ++k)
        th.emplace_back([&, k]() {
            std::vector<int32_t> piece(static_cast<size_t>(push_tokens));
            long long mine = 0;
            for (int pos = 0; pos < tokens_per_stream && !failed.load(std::memory_order_relaxed); pos += push_tokens) {
                const int n = pos + push_tokens <= tokens_per_stream ? push_tokens : tokens_per_stream - pos;
                for (int s = k; s < n_streams; s += n_threads) {
                    for (int i = 0; i < n; ++i) {
                        const int p = pos + i;
                        piece[static_cast<size_t>(i)] = 128266 + 4096 * (p % 7) + static_cast<int32_t>(mix((static_cast<uint64_t>(s) << 32) | p) & 4095u);
                    }
                    if (snacb_batcher_push(b, static_cast<uint64_t>(s), piece.data(), n) != 0) { failed.store(1); break; }
                    ++mine;
                }
                // back-pressure: a producer far ahead of the decoder only grows the queues
                while (snacb_batcher_pending(b) > 8 * max_windows && !failed.load(std::memory_order_relaxed)) std::this_thread::yield();
            }
            pushes.fetch_add(mine);
            producers_left.fetch_sub(1);
        });
    std::vector<uint64_t> ids(static_cast<size_t>(max_windows));
    std::vector<int64_t> off(static_cast<size_t>(max_windows));
    std::vector<int32_t> len(static_cast<size_t>(max_windows));
    long long windows = 0, flushes = 0;
    int outstanding = 0, slot = 0, rc = 0;
    uint64_t seed = 1;
    for (;;) {
        const bool done = producers_left.load() == 0;
        // one full batch per flush while producers run (a serving tick waits for its batch); drain at the end
        if (!done && snacb_batcher_pending(b) < max_windows) { std::this_thread::yield(); continue; }
        const int n = snacb_batcher_flush_submit(b, seed++, max_windows, ids.data(), off.data(), len.data(), slot ? pcm1 : pcm0, pcm_cap);
        if (n < 0) { rc = n; break; }
        if (n > 0) { windows += n; ++flushes; ++outstanding; slot ^= 1; }
        if (outstanding == 2 || (n == 0 && outstanding > 0)) {
            const int w = snacb_batcher_flush_wait(b);
            if (w) { rc = w; break; }
            --outstanding;
        }
        if (done && n == 0 && outstanding == 0 && snacb_batcher_pending(b) == 0) break;
    }
    if (rc) failed.store(1);
    for (auto& t : th) t.join();
    while (outstanding-- > 0) snacb_batcher_flush_wait(b);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out[0] = secs; out[1] = static_cast<double>(windows); out[2] = static_cast<double>(pushes.load()); out[3] = static_cast<double>(flushes);
    return rc ? rc : (failed.load() ? -2 : 0);
}
