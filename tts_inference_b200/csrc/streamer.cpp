// Native multi-stream front of the stateful streaming session (C ABI in include/snacb.h, snacb_streamer_*).
//
// The reference's stream_audio (vllm_inference/modal_audio_stream.py:340-409) buffers one stream's codes in Python and
// decodes a fresh window per chunk under a global lock; snacb_batcher_* batches those windows.  This is the same role for
// the stateful session: any number of producer threads push token ids of any number of streams; tick() -- one caller --
// hands every stream's new whole frames to ONE snacb_session_step_multi per distinct frame count (streams at different
// positions share the launch sequence), flushes finished streams, and returns each stream's newly final samples in the
// caller's host buffer.  Streams own a session slot from their first push until they have been flushed.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/snacb.h"

namespace {
constexpr int kFrame = 7;
constexpr int kShards = 64;

struct Stream {
    std::vector<int32_t> pending;   // ids not yet handed to the session
    int slot = -1;
    long long frames = 0;           // frames handed to the session
    int32_t key = 0;                // NoiseBlock noise key: order of first appearance
    bool ended = false;
};
struct Shard {
    std::mutex mu;
    std::unordered_map<uint64_t, Stream> streams;
};
struct Due {
    uint64_t id;
    int slot, nframes;
    long long frames;
    int32_t key;
    bool final;
    std::vector<int32_t> tok;
};
}  // namespace

struct snacb_streamer_s {
    snacb_handle h = nullptr;
    snacb_session sess = nullptr;
    int max_streams = 0, window = 0, min_frames = 1, cap = 0;
    Shard shards[kShards];
    std::mutex slot_mu;
    std::vector<int> free_slots;
    int32_t next_key = 0;
    std::mutex tick_mu;
    cudaStream_t st = nullptr;
    int32_t *pin_tok = nullptr, *dev_tok = nullptr, *dev_keys = nullptr, *pin_keys = nullptr;
    int16_t* dev_pcm = nullptr;
    size_t dev_pcm_elems = 0;
};

static inline Shard& shard_of(snacb_streamer s, uint64_t id) { return s->shards[(id * 0x9E3779B97F4A7C15ull) >> 58]; }

extern "C" {

int snacb_streamer_create(snacb_streamer* out, snacb_handle h, int max_streams, int window_frames, int flags, int min_frames) {
    if (!out || !h || max_streams < 1 || min_frames < 1) return SNACB_ERR_ARG;
    *out = nullptr;
    snacb_streamer s = new (std::nothrow) snacb_streamer_s();
    if (!s) return SNACB_ERR_NOMEM;
    s->h = h; s->max_streams = max_streams; s->min_frames = min_frames;
    int rc = snacb_session_create(h, max_streams, window_frames, flags & (SNACB_RAW_IDS | SNACB_BF16), &s->sess);
    if (rc) { delete s; return rc; }
    s->window = snacb_session_max_frames(s->sess);
    s->cap = s->window - 16;                                   // frames one step may add to a non-empty window
    if (s->min_frames > s->cap) s->min_frames = s->cap;
    for (int i = max_streams - 1; i >= 0; --i) s->free_slots.push_back(i);
    const size_t tok_ints = static_cast<size_t>(max_streams) * kFrame * s->cap;
    bool ok = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void**>(&s->pin_tok), tok_ints * sizeof(int32_t)) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void**>(&s->dev_tok), tok_ints * sizeof(int32_t)) == cudaSuccess &&
              cudaMallocHost(reinterpret_cast<void**>(&s->pin_keys), max_streams * sizeof(int32_t)) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void**>(&s->dev_keys), max_streams * sizeof(int32_t)) == cudaSuccess;
    if (!ok) { cudaGetLastError(); snacb_streamer_destroy(s); return SNACB_ERR_NOMEM; }
    *out = s;
    return SNACB_OK;
}

void snacb_streamer_destroy(snacb_streamer s) {
    if (!s) return;
    if (s->st) { cudaStreamSynchronize(s->st); cudaStreamDestroy(s->st); }
    if (s->sess) snacb_session_destroy(s->sess);
    if (s->pin_tok) cudaFreeHost(s->pin_tok);
    if (s->pin_keys) cudaFreeHost(s->pin_keys);
    if (s->dev_tok) cudaFree(s->dev_tok);
    if (s->dev_keys) cudaFree(s->dev_keys);
    if (s->dev_pcm) cudaFree(s->dev_pcm);
    delete s;
}

int snacb_streamer_push(snacb_streamer s, uint64_t stream_id, const int32_t* tokens_host, int n) {
    if (!s || n < 0 || (n > 0 && !tokens_host)) return SNACB_ERR_ARG;
    Shard& sh = shard_of(s, stream_id);
    std::lock_guard<std::mutex> g(sh.mu);
    auto it = sh.streams.find(stream_id);
    if (it == sh.streams.end()) {
        std::lock_guard<std::mutex> gs(s->slot_mu);
        if (s->free_slots.empty()) return SNACB_ERR_STATE;      // every slot holds a live stream
        Stream stn;
        stn.slot = s->free_slots.back(); s->free_slots.pop_back();
        stn.key = s->next_key; s->next_key = (s->next_key + 1) & 0x7FFFFFFF;
        it = sh.streams.emplace(stream_id, std::move(stn)).first;
    }
    if (it->second.ended) return SNACB_ERR_STATE;
    it->second.pending.insert(it->second.pending.end(), tokens_host, tokens_host + n);
    return SNACB_OK;
}

int snacb_streamer_end(snacb_streamer s, uint64_t stream_id) {
    if (!s) return SNACB_ERR_ARG;
    Shard& sh = shard_of(s, stream_id);
    std::lock_guard<std::mutex> g(sh.mu);
    auto it = sh.streams.find(stream_id);
    if (it == sh.streams.end()) return SNACB_ERR_STATE;
    it->second.ended = true;
    return SNACB_OK;
}

int snacb_streamer_active(snacb_streamer s) {
    if (!s) return SNACB_ERR_ARG;
    std::lock_guard<std::mutex> gs(s->slot_mu);
    return s->max_streams - static_cast<int>(s->free_slots.size());
}

int snacb_streamer_tick(snacb_streamer s, uint64_t seed, int max_chunks, uint64_t* ids, int64_t* offsets, int32_t* lengths,
                        int16_t* pcm_host, size_t pcm_capacity) {
    if (!s || max_chunks < 0 || (max_chunks > 0 && (!ids || !offsets || !lengths || !pcm_host))) return SNACB_ERR_ARG;
    std::lock_guard<std::mutex> gt(s->tick_mu);
    // ---- what is due: whole frames not yet handed over, at most `cap` per step; a finished stream with nothing left over
    std::vector<Due> due;
    for (auto& sh : s->shards) {
        std::lock_guard<std::mutex> g(sh.mu);
        for (auto& kv : sh.streams) {
            Stream& t = kv.second;
            const int avail = static_cast<int>(t.pending.size() / kFrame);
            const int nf = std::min(avail, s->cap);
            const bool final = t.ended && nf == avail;
            if (!final && nf < s->min_frames) continue;
            if (static_cast<int>(due.size()) >= max_chunks) continue;              // stays queued for the next tick
            Due d{kv.first, t.slot, nf, t.frames, t.key, final, {}};
            d.tok.assign(t.pending.begin(), t.pending.begin() + static_cast<size_t>(nf) * kFrame);
            t.pending.erase(t.pending.begin(), t.pending.begin() + static_cast<size_t>(nf) * kFrame);
            if (final) t.pending.clear();                                           // a ragged tail (< 7 ids) is dropped, as the reference does
            t.frames += nf;
            due.push_back(std::move(d));
        }
    }
    if (due.empty()) return 0;
    // ---- groups: streams past their third frame share a step per frame count; younger ones by exact position; finals alone
    std::stable_sort(due.begin(), due.end(), [](const Due& a, const Due& b) {
        const long long ca = a.final ? (1LL << 40) + a.slot : (a.frames >= 3 ? -1 : a.frames);
        const long long cb = b.final ? (1LL << 40) + b.slot : (b.frames >= 3 ? -1 : b.frames);
        if (ca != cb) return ca < cb;
        if (a.nframes != b.nframes) return a.nframes < b.nframes;
        return a.slot < b.slot;
    });
    int n_chunks = 0;
    int64_t off = 0;
    size_t i = 0;
    std::vector<int32_t> slots;
    int rc = SNACB_OK;
    while (i < due.size() && rc == SNACB_OK) {
        size_t j = i + 1;
        if (!due[i].final) {
            const long long cls = due[i].frames >= 3 ? -1 : due[i].frames;
            while (j < due.size() && !due[j].final && due[j].nframes == due[i].nframes &&
                   (due[j].frames >= 3 ? -1 : due[j].frames) == cls) ++j;
        }
        const int n = static_cast<int>(j - i), nf = due[i].nframes;
        const int n_emit = snacb_session_next_emit(s->sess, due[i].slot, nf, due[i].final ? 1 : 0);
        if (n_emit < 0) { rc = n_emit; break; }
        if (static_cast<size_t>(off) + static_cast<size_t>(n) * n_emit > pcm_capacity) { rc = SNACB_ERR_ARG; break; }
        if (static_cast<size_t>(n) * n_emit > s->dev_pcm_elems) {
            if (s->dev_pcm) cudaFree(s->dev_pcm);
            s->dev_pcm = nullptr; s->dev_pcm_elems = 0;
            if (cudaMalloc(reinterpret_cast<void**>(&s->dev_pcm), static_cast<size_t>(n) * n_emit * sizeof(int16_t)) != cudaSuccess) { rc = SNACB_ERR_NOMEM; break; }
            s->dev_pcm_elems = static_cast<size_t>(n) * n_emit;
        }
        slots.resize(n);
        for (int k = 0; k < n; ++k) {
            slots[k] = due[i + k].slot;
            s->pin_keys[k] = due[i + k].key;
            if (nf > 0) std::memcpy(s->pin_tok + static_cast<size_t>(k) * kFrame * nf, due[i + k].tok.data(), static_cast<size_t>(kFrame) * nf * sizeof(int32_t));
        }
        // the pinned staging is reused by the next group: its copies must have been consumed
        if (cudaStreamSynchronize(s->st) != cudaSuccess) { rc = SNACB_ERR_CUDA; break; }
        cudaMemcpyAsync(s->dev_keys, s->pin_keys, n * sizeof(int32_t), cudaMemcpyHostToDevice, s->st);
        if (nf > 0) cudaMemcpyAsync(s->dev_tok, s->pin_tok, static_cast<size_t>(n) * kFrame * nf * sizeof(int32_t), cudaMemcpyHostToDevice, s->st);
        int got = 0;
        if (due[i].final)
            rc = snacb_session_step(s->sess, slots[0], 1, nf > 0 ? s->dev_tok : nullptr, kFrame * nf, nf, 1, seed, s->dev_keys, s->dev_pcm,
                                    n_emit, &got, s->st);
        else
            rc = snacb_session_step_multi(s->sess, n, slots.data(), s->dev_tok, kFrame * nf, nf, seed, s->dev_keys, s->dev_pcm, n_emit,
                                          &got, s->st);
        if (rc) break;
        if (got > 0)
            cudaMemcpyAsync(pcm_host + off, s->dev_pcm, static_cast<size_t>(n) * got * sizeof(int16_t), cudaMemcpyDeviceToHost, s->st);
        for (int k = 0; k < n; ++k) {
            if (got > 0) { ids[n_chunks] = due[i + k].id; offsets[n_chunks] = off + static_cast<int64_t>(k) * got; lengths[n_chunks] = got; ++n_chunks; }
        }
        off += static_cast<int64_t>(n) * got;
        // dev_pcm is reused by the next group
        if (cudaStreamSynchronize(s->st) != cudaSuccess) { rc = SNACB_ERR_CUDA; break; }
        i = j;
    }
    if (rc != SNACB_OK) return rc;
    // ---- finished streams leave: slot reset and returned, id forgotten
    for (const Due& d : due) {
        if (!d.final) continue;
        snacb_session_reset(s->sess, d.slot, 1);
        {
            Shard& sh = shard_of(s, d.id);
            std::lock_guard<std::mutex> g(sh.mu);
            sh.streams.erase(d.id);
        }
        std::lock_guard<std::mutex> gs(s->slot_mu);
        s->free_slots.push_back(d.slot);
    }
    return n_chunks;
}

}  // extern "C"
