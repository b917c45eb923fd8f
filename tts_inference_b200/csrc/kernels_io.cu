// The two steps either side of the decode path, device-resident:
//
//   * k_ingest   LLM token ids -> ready SNAC windows, for many streams at once.  Restates, per stream,
//                generate_audio_tokens' gate (vllm_inference/modal_audio_stream.py:313-333: drop everything up to
//                and including the first TOKEN_SOS, stop at TOKEN_EOS; same rule in tensorrt_tts/inference.py:231-241)
//                followed by stream_audio's buffer policy (:352-396: pop the first 28 codes whenever 28 are
//                buffered; at the end of the stream emit the remaining whole frames).  Integer work, bit-exact.
//   * k_base64   int16 PCM chunks -> base64 text, what the /ws/audio endpoint sends per chunk
//                (modal_audio_stream.py:483-487: base64.b64encode(audio_chunk)).
//   * k_wav      int16 PCM -> RIFF/WAVE byte stream, what /generate and /generate-batch return
//                (modal_audio_stream.py:561-566, 650-657: wave.open(...); setnchannels(1); setsampwidth(2);
//                setframerate(24000); writeframes(...)).
//
// All three are HBM-bound byte / integer passes; none touches the tensor cores.
#include <cstdint>
#include <new>

#include <cuda_runtime.h>

#include "../../include/snacb.h"

namespace {

constexpr int kFrame = 7;
constexpr int kWindow = 28;             // CHUNK_TOKENS, modal_audio_stream.py:353
constexpr int kTailTok = 21;            // at most 3 whole frames are left at the end of a stream
constexpr int kTokenSos = 128257;       // TOKEN_SOS, modal_audio_stream.py:101
constexpr int kTokenEos = 128258;       // TOKEN_EOS, modal_audio_stream.py:102
constexpr int kIngestThreads = 1024;

enum { ST_WAIT_SOS = 0, ST_SPEECH = 1, ST_ENDED = 2 };

struct IngestOut {
    int32_t* win_tok;      // [cap][28] token ids of the full windows, (stream, time) order
    int32_t* win_stream;   // [cap]
    int32_t* tail_tok;     // [S][21] token ids of end-of-stream remainders, zero padded
    int32_t* tail_stream;  // [S]
    int32_t* tail_frames;  // [S] 1..3
    int32_t* counts;       // [2] full windows, tails
};

// One stream, n tokens of one step.  WRITE = false: count what the step will emit.  WRITE = true: emit at
// (wpos, tpos) and store the new state.  Both passes start from the same stored state, so they agree.
// The unfinished window of a stream lives in buf[s][28] between steps.
template <bool WRITE>
__device__ __forceinline__ void ingest_stream(const int32_t* __restrict__ tok, int n, bool finish, int32_t* state,
                                              int32_t* count, int32_t* buf /*[28]*/, int s, int& n_win, int& n_tail,
                                              const IngestOut& o, int wpos, int tpos) {
    const int st0 = *state;
    int st = st0, cnt = *count;
    n_win = 0; n_tail = 0;
    if (st0 == ST_ENDED) return;                               // :330 break / engine.abort: later tokens are never seen
    for (int i = 0; i < n; ++i) {
        const int32_t t = tok[i];
        if (st == ST_WAIT_SOS) {                               // :320-325 skip until TOKEN_SOS, and skip it too
            if (t == kTokenSos) st = ST_SPEECH;
            continue;
        }
        if (t == kTokenEos) { finish = true; break; }          // :328-330
        // :366-367 every other id is appended (ids stay raw here: snacb_decode subtracts 128266, SNACB_RAW_IDS)
        if (WRITE) buf[cnt] = t;
        ++cnt;
        if (cnt == kWindow) {                                  // :370-372 pop the first 28
            if (WRITE) {
                int32_t* row = o.win_tok + static_cast<size_t>(wpos + n_win) * kWindow;
                for (int k = 0; k < kWindow; ++k) row[k] = buf[k];
                o.win_stream[wpos + n_win] = s;
            }
            ++n_win;
            cnt = 0;
        }
    }
    if (finish) {                                              // :391-396 remainder trimmed to whole frames
        const int fr = cnt / kFrame;                           // cnt is 0 if TOKEN_SOS never came
        if (fr > 0) {
            if (WRITE) {
                int32_t* row = o.tail_tok + static_cast<size_t>(tpos) * kTailTok;
                for (int k = 0; k < kTailTok; ++k) row[k] = (k < fr * kFrame) ? buf[k] : 0;
                o.tail_stream[tpos] = s;
                o.tail_frames[tpos] = fr;
            }
            n_tail = 1;
        }
        st = ST_ENDED;
        cnt = 0;
    }
    if (WRITE) { *state = st; *count = cnt; }
}

__device__ __forceinline__ unsigned long long warp_inclusive(unsigned long long v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
    }
    return v;
}

// Single CTA: the output order is (stream, time), i.e. a prefix sum over streams, and a step is tiny
// (S streams x a handful of tokens), so one CTA walking the streams 1024 at a time is the latency-optimal shape.
__global__ void __launch_bounds__(kIngestThreads)
k_ingest(const int32_t* __restrict__ tok, int S, int n_tok, const int32_t* __restrict__ n_valid,
         const uint8_t* __restrict__ finish, int32_t* state, int32_t* count, int32_t* buf, IngestOut o) {
    __shared__ unsigned long long warp_off[32];
    __shared__ unsigned long long round_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long base = 0ull;                            // (full windows | tails << 32) emitted by earlier rounds
    for (int s0 = 0; s0 < S; s0 += kIngestThreads) {
        const int s = s0 + tid;
        int nw = 0, nt = 0, n = 0;
        bool fin = false;
        if (s < S) {
            n = n_valid ? min(max(n_valid[s], 0), n_tok) : n_tok;
            fin = finish != nullptr && finish[s] != 0;
            ingest_stream<false>(tok + static_cast<size_t>(s) * n_tok, n, fin, state + s, count + s,
                                 buf + static_cast<size_t>(s) * kWindow, s, nw, nt, o, 0, 0);
        }
        const unsigned long long v = static_cast<unsigned long long>(nw) | (static_cast<unsigned long long>(nt) << 32);
        const unsigned long long inc = warp_inclusive(v, lane);
        if (lane == 31) warp_off[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = warp_off[lane], winc = warp_inclusive(w, lane);
            warp_off[lane] = winc - w;
            if (lane == 31) round_total = winc;
        }
        __syncthreads();
        const unsigned long long excl = base + warp_off[warp] + (inc - v);
        if (s < S)
            ingest_stream<true>(tok + static_cast<size_t>(s) * n_tok, n, fin, state + s, count + s,
                                buf + static_cast<size_t>(s) * kWindow, s, nw, nt, o,
                                static_cast<int>(excl & 0xffffffffull), static_cast<int>(excl >> 32));
        base += round_total;
        __syncthreads();
    }
    if (tid == 0) {
        o.counts[0] = static_cast<int32_t>(base & 0xffffffffull);
        o.counts[1] = static_cast<int32_t>(base >> 32);
    }
}

// ---------------------------------------------------------------------------------------------------------
// base64 (RFC 4648, '=' padded) of n_chunks independent byte strings of `bytes` bytes each: one thread per
// 3-byte group -> one aligned 4-byte store.  Reference: base64.b64encode(audio_chunk), modal_audio_stream.py:484.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t b64_char(uint32_t c) {
    // 0..25 'A'.., 26..51 'a'.., 52..61 '0'.., 62 '+', 63 '/'
    return c < 26u ? c + 65u : (c < 52u ? c + 71u : (c < 62u ? c - 4u : (c == 62u ? 43u : 47u)));
}

__global__ void __launch_bounds__(256)
k_base64(const uint8_t* __restrict__ in, long long n_chunks, long long bytes, uint8_t* __restrict__ out) {
    const long long groups = (bytes + 2) / 3;                  // per chunk
    const long long total = n_chunks * groups;
#pragma unroll 4
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < total;
         g += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long c = g / groups, k = g - c * groups;
        const uint8_t* src = in + c * bytes + 3 * k;
        const long long left = bytes - 3 * k;                  // 1, 2 or >= 3
        const uint32_t b0 = src[0], b1 = left > 1 ? src[1] : 0u, b2 = left > 2 ? src[2] : 0u;
        const uint32_t w = (b0 << 16) | (b1 << 8) | b2;
        uint32_t o = b64_char(w >> 18) | (b64_char((w >> 12) & 63u) << 8);
        o |= (left > 1 ? b64_char((w >> 6) & 63u) : 61u) << 16;            // '='
        o |= (left > 2 ? b64_char(w & 63u) : 61u) << 24;
        reinterpret_cast<uint32_t*>(out)[g] = o;               // chunk c starts at 4 * groups * c: always 4-aligned
    }
}

// ---------------------------------------------------------------------------------------------------------
// RIFF/WAVE framing of n independent mono 16-bit PCM strings: 44-byte header + samples, exactly the bytes the
// stdlib `wave` writer produces for setnchannels(1) / setsampwidth(2) / setframerate(rate) / writeframes(pcm)
// (modal_audio_stream.py:561-566, 650-657).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_wav(const int16_t* __restrict__ pcm, long long n, long long samples, uint32_t rate, uint8_t* __restrict__ out) {
    const long long stride = 44 + 2 * samples;                 // bytes per output record, even
    const long long words = stride / 2;                        // 16-bit words per record (22 header words + samples)
    const long long total = n * words;
    const uint32_t data_len = static_cast<uint32_t>(2 * samples);
#pragma unroll 4
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / words, k = i - r * words;
        uint16_t v;
        if (k >= 22) {
            v = static_cast<uint16_t>(pcm[r * samples + (k - 22)]);
        } else {
            const uint32_t riff_len = 36u + data_len, byte_rate = rate * 2u;
            switch (k) {
                case 0: v = 0x4952; break;  /* "RI" */      case 1: v = 0x4646; break;  /* "FF" */
                case 2: v = riff_len & 0xffffu; break;      case 3: v = riff_len >> 16; break;
                case 4: v = 0x4157; break;  /* "WA" */      case 5: v = 0x4556; break;  /* "VE" */
                case 6: v = 0x6d66; break;  /* "fm" */      case 7: v = 0x2074; break;  /* "t " */
                case 8: v = 16; break;                      case 9: v = 0; break;        /* fmt chunk size */
                case 10: v = 1; break;      /* PCM */       case 11: v = 1; break;       /* mono */
                case 12: v = rate & 0xffffu; break;         case 13: v = rate >> 16; break;
                case 14: v = byte_rate & 0xffffu; break;    case 15: v = byte_rate >> 16; break;
                case 16: v = 2; break;      /* block align */ case 17: v = 16; break;    /* bits */
                case 18: v = 0x6164; break; /* "da" */      case 19: v = 0x6174; break;  /* "ta" */
                case 20: v = data_len & 0xffffu; break;     default: v = data_len >> 16; break;
            }
        }
        reinterpret_cast<uint16_t*>(out)[i] = v;
    }
}

}  // namespace

// =========================================================================================================
// C ABI (include/snacb.h)
// =========================================================================================================
struct snacb_ingest_s {
    int device = 0, max_streams = 0;
    int32_t *state = nullptr, *count = nullptr, *buf = nullptr;
};

extern "C" {

int snacb_ingest_create(snacb_ingest* out, int device, int max_streams) {
    if (!out || max_streams <= 0) return SNACB_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SNACB_ERR_NO_GPU;
    if (cudaSetDevice(device) != cudaSuccess) return SNACB_ERR_CUDA;
    snacb_ingest g = new (std::nothrow) snacb_ingest_s();
    if (!g) return SNACB_ERR_NOMEM;
    g->device = device; g->max_streams = max_streams;
    const size_t n = static_cast<size_t>(max_streams);
    if (cudaMalloc(reinterpret_cast<void**>(&g->state), n * 4) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&g->count), n * 4) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&g->buf), n * kWindow * 4) != cudaSuccess ||
        cudaMemset(g->state, 0, n * 4) != cudaSuccess || cudaMemset(g->count, 0, n * 4) != cudaSuccess) {
        snacb_ingest_destroy(g);
        return SNACB_ERR_NOMEM;
    }
    *out = g;
    return SNACB_OK;
}

void snacb_ingest_destroy(snacb_ingest g) {
    if (!g) return;
    cudaSetDevice(g->device);
    cudaFree(g->state); cudaFree(g->count); cudaFree(g->buf);
    delete g;
}

int snacb_ingest_reset(snacb_ingest g, int first, int n, void* stream) {
    if (!g || first < 0 || n < 0 || first + n > g->max_streams) return SNACB_ERR_ARG;
    if (n == 0) return SNACB_OK;
    if (cudaSetDevice(g->device) != cudaSuccess) return SNACB_ERR_CUDA;     // the caller's current device may be another GPU
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(g->state + first, 0, static_cast<size_t>(n) * 4, st) != cudaSuccess ||
        cudaMemsetAsync(g->count + first, 0, static_cast<size_t>(n) * 4, st) != cudaSuccess)
        return SNACB_ERR_CUDA;
    return SNACB_OK;
}

int snacb_ingest_window_capacity(int S, int n_tok) {
    if (S < 0 || n_tok < 0) return SNACB_ERR_ARG;
    const long long cap = static_cast<long long>(S) * ((n_tok + kWindow - 1) / kWindow + 1);
    return cap > 0x7fffffffLL ? SNACB_ERR_ARG : static_cast<int>(cap);
}

int snacb_ingest_step(snacb_ingest g, const int32_t* tok, int S, int n_tok, const int32_t* n_valid,
                      const uint8_t* finish, int32_t* win_tok, int32_t* win_stream, int win_cap, int32_t* tail_tok,
                      int32_t* tail_stream, int32_t* tail_frames, int32_t* counts, void* stream) {
    if (!g || S < 0 || S > g->max_streams || n_tok < 0 || (n_tok > 0 && !tok) || !win_tok || !win_stream ||
        !tail_tok || !tail_stream || !tail_frames || !counts)
        return SNACB_ERR_ARG;
    const int need = snacb_ingest_window_capacity(S, n_tok);
    if (need < 0 || win_cap < need) return SNACB_ERR_ARG;
    const IngestOut o{win_tok, win_stream, tail_tok, tail_stream, tail_frames, counts};
    if (cudaSetDevice(g->device) != cudaSuccess) return SNACB_ERR_CUDA;
    k_ingest<<<1, kIngestThreads, 0, static_cast<cudaStream_t>(stream)>>>(tok, S, n_tok, n_valid, finish, g->state,
                                                                          g->count, g->buf, o);
    return cudaGetLastError() == cudaSuccess ? SNACB_OK : SNACB_ERR_CUDA;
}

int snacb_ingest_state(snacb_ingest g, int32_t* state_host, int32_t* count_host, int n) {
    if (!g || n < 0 || n > g->max_streams) return SNACB_ERR_ARG;
    if (cudaSetDevice(g->device) != cudaSuccess) return SNACB_ERR_CUDA;
    if (state_host && cudaMemcpy(state_host, g->state, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return SNACB_ERR_CUDA;
    if (count_host && cudaMemcpy(count_host, g->count, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return SNACB_ERR_CUDA;
    return SNACB_OK;
}

// The egress calls take no handle: launch on the GPU that owns the PCM buffer, whatever the caller's current device is.
static bool use_device_of(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return false;
    return cudaSetDevice(at.device) == cudaSuccess;
}

long long snacb_base64_len(long long bytes) { return bytes < 0 ? SNACB_ERR_ARG : 4 * ((bytes + 2) / 3); }

int snacb_pcm_to_base64(const int16_t* pcm, long long n_chunks, long long samples, uint8_t* out, void* stream) {
    if (n_chunks < 0 || samples < 0 || ((n_chunks > 0 && samples > 0) && (!pcm || !out))) return SNACB_ERR_ARG;
    const long long groups = n_chunks * ((2 * samples + 2) / 3);
    if (groups == 0) return SNACB_OK;
    if (!use_device_of(pcm)) return SNACB_ERR_ARG;                 // pcm must be a device pointer
    const long long want = (groups + 255) / 256;
    const int blocks = static_cast<int>(want < 148LL * 32 ? want : 148LL * 32);    // grid-stride beyond 32 CTAs per SM
    k_base64<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint8_t*>(pcm), n_chunks,
                                                                    2 * samples, out);
    return cudaGetLastError() == cudaSuccess ? SNACB_OK : SNACB_ERR_CUDA;
}

int snacb_pcm_to_wav(const int16_t* pcm, long long n, long long samples, int sample_rate, uint8_t* out, void* stream) {
    if (n < 0 || samples < 0 || sample_rate <= 0 || (n > 0 && (!out || (samples > 0 && !pcm)))) return SNACB_ERR_ARG;
    if (2 * samples + 36 > 0xffffffffLL) return SNACB_ERR_ARG;
    const long long total = n * (22 + samples);
    if (total == 0) return SNACB_OK;
    if (!use_device_of(out)) return SNACB_ERR_ARG;                 // out must be a device pointer
    const long long want = (total + 255) / 256;
    const int blocks = static_cast<int>(want < 148LL * 32 ? want : 148LL * 32);
    k_wav<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(pcm, n, samples, static_cast<uint32_t>(sample_rate), out);
    return cudaGetLastError() == cudaSuccess ? SNACB_OK : SNACB_ERR_CUDA;
}

}  // extern "C"
