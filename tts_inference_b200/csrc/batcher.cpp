// Multi-stream window batcher (C ABI in include/snacb.h).
//
// Replaces the per-stream buffer policy of stream_audio (vllm_inference/modal_audio_stream.py:352-396),
// which decodes ONE stream's chunk per call under a global asyncio.Lock (:83, :476).  Here any number of
// producers push token ids; flush() packs every ready window of every stream into one batched decode.
#include <algorithm>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/snacb.h"

namespace {
constexpr int kFrame = 7;
constexpr int kWindow = 28;     // MIN_FRAMES_SUBSEQ, modal_audio_stream.py:92

struct StreamState {
    std::vector<int32_t> buf;   // policy 0: unconsumed codes; policy 1: last <= 28 codes
    uint64_t count = 0;         // tokens seen (policy 1)
    bool ended = false;
};
struct Item {
    uint64_t id;
    int frames;
    int32_t tok[kWindow];
};
}  // namespace

struct snacb_batcher_s {
    snacb_handle h;
    int policy, flags, max_windows;
    std::mutex mu;
    std::unordered_map<uint64_t, StreamState> streams;
    std::deque<Item> ready;
    int32_t* pin_tok = nullptr;  // pinned staging [max_windows][28]
    int16_t* pin_pcm = nullptr;  // pinned staging [max_windows][8192]
};

static void queue_from_stream(snacb_batcher b, uint64_t id, StreamState& s) {
    if (b->policy == 0) {
        size_t pos = 0;
        while (s.buf.size() - pos >= static_cast<size_t>(kWindow)) {
            Item it{id, 4, {}};
            std::memcpy(it.tok, s.buf.data() + pos, kWindow * sizeof(int32_t));
            b->ready.push_back(it);
            pos += kWindow;
        }
        if (s.ended) {
            const int fr = static_cast<int>((s.buf.size() - pos) / kFrame);
            if (fr > 0) {
                Item it{id, fr, {}};
                std::memcpy(it.tok, s.buf.data() + pos, static_cast<size_t>(fr) * kFrame * sizeof(int32_t));
                b->ready.push_back(it);
            }
            pos = s.buf.size();
        }
        if (pos) s.buf.erase(s.buf.begin(), s.buf.begin() + pos);
    }
}

extern "C" {

int snacb_batcher_create(snacb_batcher* out, snacb_handle h, int policy, int flags, int max_windows) {
    if (!out || !h || (policy != 0 && policy != 1) || max_windows <= 0) return SNACB_ERR_ARG;
    snacb_batcher b = new (std::nothrow) snacb_batcher_s();
    if (!b) return SNACB_ERR_NOMEM;
    b->h = h; b->policy = policy; b->max_windows = max_windows;
    b->flags = flags & (SNACB_RAW_IDS | SNACB_FP32 | SNACB_STREAM_FP32 | SNACB_BF16);
    if (policy == 1) b->flags |= SNACB_EXTRACT_SLICE;
    if (cudaMallocHost(reinterpret_cast<void**>(&b->pin_tok), static_cast<size_t>(max_windows) * kWindow * 4) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void**>(&b->pin_pcm), static_cast<size_t>(max_windows) * 8192 * 2) != cudaSuccess) {
        snacb_batcher_destroy(b);
        return SNACB_ERR_NOMEM;
    }
    *out = b;
    return SNACB_OK;
}

void snacb_batcher_destroy(snacb_batcher b) {
    if (!b) return;
    if (b->pin_tok) cudaFreeHost(b->pin_tok);
    if (b->pin_pcm) cudaFreeHost(b->pin_pcm);
    delete b;
}

int snacb_batcher_push(snacb_batcher b, uint64_t id, const int32_t* tokens, int n) {
    if (!b || n < 0 || (n > 0 && !tokens)) return SNACB_ERR_ARG;
    std::lock_guard<std::mutex> lk(b->mu);
    StreamState& s = b->streams[id];
    if (s.ended) return SNACB_ERR_STATE;
    if (b->policy == 0) {
        s.buf.insert(s.buf.end(), tokens, tokens + n);
        queue_from_stream(b, id, s);
    } else {
        // sliding rule: once >= 28 tokens have been seen, every 7th token decodes the last 28
        for (int i = 0; i < n; ++i) {
            s.buf.push_back(tokens[i]);
            if (s.buf.size() > static_cast<size_t>(kWindow)) s.buf.erase(s.buf.begin());
            ++s.count;
            if (s.count % kFrame == 0 && s.count >= static_cast<uint64_t>(kWindow)) {
                Item it{id, 4, {}};
                std::memcpy(it.tok, s.buf.data(), kWindow * sizeof(int32_t));
                b->ready.push_back(it);
            }
        }
    }
    return SNACB_OK;
}

int snacb_batcher_end(snacb_batcher b, uint64_t id) {
    if (!b) return SNACB_ERR_ARG;
    std::lock_guard<std::mutex> lk(b->mu);
    auto it = b->streams.find(id);
    if (it == b->streams.end()) return SNACB_OK;
    it->second.ended = true;
    queue_from_stream(b, id, it->second);
    b->streams.erase(it);
    return SNACB_OK;
}

int snacb_batcher_pending(snacb_batcher b) {
    if (!b) return SNACB_ERR_ARG;
    std::lock_guard<std::mutex> lk(b->mu);
    return static_cast<int>(b->ready.size());
}

int snacb_batcher_flush(snacb_batcher b, uint64_t seed, int max_chunks, uint64_t* ids, int64_t* offsets,
                        int32_t* lengths, int16_t* pcm_host, size_t pcm_capacity) {
    if (!b || max_chunks < 0 || !ids || !offsets || !lengths || !pcm_host) return SNACB_ERR_ARG;
    std::vector<Item> items;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        size_t used = 0;
        while (!b->ready.empty() && static_cast<int>(items.size()) < max_chunks &&
               static_cast<int>(items.size()) < b->max_windows) {
            const Item& it = b->ready.front();
            const size_t n = static_cast<size_t>(snacb_samples_out(it.frames, b->flags));
            if (used + n > pcm_capacity) break;
            used += n;
            items.push_back(it);
            b->ready.pop_front();
        }
    }
    if (items.empty()) return 0;
    // output layout follows queue order; decode launches are grouped by frame count
    int64_t off = 0;
    for (size_t i = 0; i < items.size(); ++i) {
        ids[i] = items[i].id;
        offsets[i] = off;
        lengths[i] = snacb_samples_out(items[i].frames, b->flags);
        off += lengths[i];
    }
    for (int fr = 4; fr >= 1; --fr) {
        std::vector<size_t> sel;
        for (size_t i = 0; i < items.size(); ++i) if (items[i].frames == fr) sel.push_back(i);
        if (sel.empty()) continue;
        const int n = snacb_samples_out(fr, b->flags);
        for (size_t j = 0; j < sel.size(); ++j)
            std::memcpy(b->pin_tok + j * kWindow, items[sel[j]].tok, kWindow * sizeof(int32_t));
        int rc = snacb_decode_host(b->h, b->pin_tok, static_cast<int>(sel.size()), kWindow, fr, b->flags, seed, b->pin_pcm);
        if (rc) return rc;
        for (size_t j = 0; j < sel.size(); ++j)
            std::memcpy(pcm_host + offsets[sel[j]], b->pin_pcm + j * static_cast<size_t>(n), static_cast<size_t>(n) * 2);
    }
    return static_cast<int>(items.size());
}

}  // extern "C"
