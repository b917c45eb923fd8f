// Shared device helpers and the launch-side problem descriptors of the SNAC decode path.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace snacb {

constexpr int kTokenAudioBase = 128266;   // reference: vllm_inference/modal_audio_stream.py:103
constexpr int kFrame = 7;
constexpr int kLatent = 768;
constexpr int kDecDim = 1024;
constexpr int kCodebook = 4096;
constexpr int kCodeDim = 8;

// Epilogue variants of the row-GEMMs (see DESIGN.md "kernels").
enum Epi : int {
    EPI_BIAS = 0,        // out = acc + bias[o]                          (ConvTranspose1d)
    EPI_BIAS_SNAKE = 1,  // out = snake(acc + bias[o]; alpha[o])         (stem 1x1 -> block-0 input)
    EPI_NOISE = 2,       // out = y[row][o] + noise[row] * acc           (NoiseBlock)
    EPI_RES = 3,         // out = x[row][o] + acc + bias[o]              (ResidualUnit)
    EPI_RES_SNAKE = 4,   // out = snake(x + acc + bias; alpha[o])        (last ResidualUnit of a block)
};

// Per-stream addressing of a streaming-session launch (all null for the stateless decode): stream s of the launch lives in
// buffer slot slot[s]; its row range starts off[s] FRAMES after the launch's reference range (streams at different
// positions that advance by the same number of frames share one launch: past a stream's first three frames every
// stage's frontier is affine in the frame count); its window's row 0 is stream frame org[s] (the NoiseBlock noise is
// keyed by the absolute time step).  `rpf` (rows per frame of the tensor at hand) travels beside the map.
struct StreamMap { const int* slot; const int* off; const int* org; };
__device__ __forceinline__ int sm_slot(const StreamMap& m, int s) { return m.slot ? m.slot[s] : s; }
__device__ __forceinline__ int sm_off(const StreamMap& m, int s, int rpf) { return m.off ? m.off[s] * rpf : 0; }
__device__ __forceinline__ int sm_org(const StreamMap& m, int s, int rpf) { return m.org ? m.org[s] * rpf : 0; }

// A "row GEMM with taps":  out[(s, m*up + p), o] = epi( sum_tap sum_k A[(s, m + shift(p,tap)), k] * W[(p,o), tap*K + k] )
// with A rows outside [0, Tin) of their stream reading as zero.  Plain 1x1 convs are ntaps=1, up=1.
struct GemmArgs {
    int S;          // streams (windows / utterances)
    int Tin;        // A rows per stream
    int K;          // channels in  (per tap)
    int N;          // total output columns = up * Cout
    int Cout;       // channels out (per phase)
    int ntaps;      // 1 (1x1 conv) or 2 (ConvTranspose1d, k = 2*stride)
    int up;         // output rows per input row (ConvTranspose stride), 1 for 1x1
    int t_lo, t_n;  // A rows [t_lo, t_lo + t_n) of every stream are processed (t_n = 0: all of [0, Tin)); dead-sample
                    // trimming for the sliced output -- rows outside are neither read nor written
    int Tbox;       // rows of one stream per 128-row tile  (Tbox * Wbox == 128)
    int Wbox;       // streams per tile
    const float* bias;      // [Cout] or null
    const float* alpha;     // [Cout] snake alpha (EPI_*_SNAKE)
    const float* inv_alpha; // [Cout] 1 / (alpha + 1e-9)
    const float* noise;     // [S][Tin] injected noise (EPI_NOISE) or null -> counter RNG
    unsigned long long seed;
    int noise_stage;        // decoder block index, keys the counter RNG
    int stream_offset;      // global index of stream 0 of this chunk (counter RNG addressing)
    const int* stream_keys; // optional [S]: counter RNG key of each stream instead of stream_offset + s
    int t0;                 // absolute index of output row 0 in its stream (counter RNG): a streaming session keeps a sliding
                            // window of each stream in its buffers, the noise stays keyed by the absolute time step
    StreamMap map; int rpf; // streaming session: per-stream slot / row offset / origin; rpf = A (input) rows per frame
    int a_wrap;             // > 0: the A tensor holds only a_wrap 64-column chunks; K chunk kc reads chunk kc % a_wrap (bf16x3:
                            // A'' = [hi | lo | hi] and [a | a] are never materialised beyond [hi | lo] and [a])
    int tma_epi;            // k_gemm_tc, NoiseBlock epilogue: 1 = y arrives and x leaves by TMA through per-warp staging tiles
                            // (needs Tbox = 128, Wbox = 1, up = 1, BN = 256, 16-bit out; whole 32-row groups only)
    int mma_bf16;           // 1: the MMA reads its operands as bf16 whatever the storage type of resid / out is (the bf16x3
                            // path: fp16 storage, A and W pre-split into bf16 hi / lo column blocks, see snacb.cu)
    const void* resid;      // residual / y tensor, [S*Tin*up][Cout] (16-bit operand type on the tensor-core path)
    void* out;              // [S*Tin*up][Cout]
};

struct ResUnitArgs {
    int S, T, C, dil;
    int t_lo, t_n;           // output rows [t_lo, t_lo + t_n) of every stream (t_n = 0: all)
    const void* x;           // [S*T][C] residual stream (XT)
    void* out;               // [S*T][C]
    const float* alpha1; const float* inv_alpha1;   // [C]
    const float* dw_w;       // [7][C]
    const float* dw_b;       // [C]
    const float* alpha2; const float* inv_alpha2;   // [C]
    const float* pw_b;       // [C]
    const float* alpha_next; const float* inv_alpha_next;  // [C] (EPI_RES_SNAKE)
    unsigned long long* prof;   // debug: clock64 sums of CTA 0 (SNACB_RES_PROF=1), else null
    StreamMap map; int rpf;     // streaming session (see StreamMap); rpf = rows per frame of x / out
};

// Fused NoiseBlock + 3 ResidualUnits of one DecoderBlock (kernels_chain.cu).
constexpr int kChainWarps = 16;        // at most; a launch configuration may use 8
constexpr int kChainSpans = 4;         // spans per warp and layer, at most
constexpr int kChainWsP = 7;           // prologue warps of the warp-specialised chain kernel (kernels_chain_ws.cu)
constexpr int kChainHalo = 40;
struct ChainSpan { short r_first, n_oct, kc, pad; };   // a warp's run of 8-step octets along one dilation class; pad: see ChainArgs
struct ChainLayer {
    const float *alpha1, *inv1;   // [C]
    const float* dw_w;            // [7][C]
    const float* dw_b;            // [C]
    const float *alpha2, *inv2;   // [C]
};
struct ChainArgs {
    int S, T, C;
    int t_lo, t_n;                // output rows [t_lo, t_lo + t_n) of every stream (t_n = 0: all)
    void* out;                    // [S*T][C] 16-bit, Snake of the next layer applied
    ChainLayer res[3];
    const float* bias_cum;        // [3][C]: b_0, b_0+b_1, b_0+b_1+b_2 (1x1 biases, added when TMEM is read)
    const float *alpha_next, *inv_next;   // [C]
    const float* noise;           // [S][T] injected noise or null -> counter RNG
    unsigned long long seed;
    int noise_stage, stream_offset;
    const int* stream_keys;       // optional [S]: counter RNG key of each stream instead of stream_offset + s
    int t0;                       // absolute index of row 0 in its stream (counter RNG; see GemmArgs::t0)
    StreamMap map; int rpf;       // streaming session (see StreamMap); rpf = rows per frame of y / out
    // Tiles and strips (kernels_chain.cu).  A stream's row range is cut into STRIPS of strip_tiles tiles that one CTA
    // walks in order.  The first tile of a strip is a HALO-TOP tile (40 rows of recomputed context above its owned rows,
    // schedule `spans`); every further tile is a CARRY-TOP tile: it owns its rows from row 0 on, the three dilation-class
    // rows above each class coming from the previous tile through `carry` (schedule `spans_carry`, whose top spans carry
    // 1 + the number of steps above row 0 in ChainSpan::pad).  strip_tiles = 1: every tile is a halo-top tile.
    // The last tile of a stream's last strip owns only last_rows rows when the range is not a whole number of tiles
    // (0: it is a full tile): it runs `spans_last` (of its own type) and skips everything past its right halo.
    ChainSpan spans[3][kChainWarps][kChainSpans];
    ChainSpan spans_carry[3][kChainWarps][kChainSpans];
    ChainSpan spans_last[3][kChainWarps][kChainSpans];
    int last_rows;
    int strip_tiles, sps, last_strip_tiles;   // tiles per strip, strips per stream, tiles of a stream's last strip
    void* carry;                  // [gridDim.x][2 (tile parity)][39][C] 16-bit scratch (strip_tiles > 1)
    unsigned jitter;              // debug (SNACB_CHAIN_JITTER=seed): every warp spins a pseudo-random 0..4095 cycles after each
                                  // barrier of the in-place prologue -- the race detector for the schedule (see k_chain)
    int* tile_counter;            // zeroed before the launch: tiles beyond the first gridDim.x are claimed dynamically
    unsigned long long* prof;     // debug: 20 per-phase clock64 sums of CTA 0 / thread 0 (SNACB_CHAIN_PROF=1), else null
};

// ---------------------------------------------------------------- math
// snake(x) = x + (alpha + 1e-9)^-1 * sin(alpha x)^2   (upstream snac layers.py; oracle/snac_ref.py)
template <bool kFast>
__device__ __forceinline__ float snake_f(float x, float alpha, float inv_alpha) {
    float s = kFast ? __sinf(alpha * x) : sinf(alpha * x);
    return fmaf(inv_alpha, s * s, x);
}

// Packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 on 64-bit register pairs) -- halves the issue slots of the
// channel-pair math in the ResidualUnit prologue.
__device__ __forceinline__ unsigned long long f2_as_u64(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 u64_as_f2(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)), "l"(f2_as_u64(c)));
    return u64_as_f2(r);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_as_u64(a)), "l"(f2_as_u64(b)));
    return u64_as_f2(r);
}
// snake on a channel pair (fast sin): x + inv_alpha * sin(alpha x)^2
__device__ __forceinline__ float2 snake_pair(float2 x, float2 alpha, float2 inv_alpha) {
    const float2 t = fmul2(alpha, x);
    const float2 s = make_float2(__sinf(t.x), __sinf(t.y));
    return ffma2(inv_alpha, fmul2(s, s), x);
}

// int16 quantise of the reference helper: (x*32767).clamp(-32768,32767).to(int16) -- truncation
// toward zero (vllm_inference/modal_audio_stream.py:201; numpy astype in tensorrt_tts/inference.py:110).
__device__ __forceinline__ int16_t pcm16(float x) {
    float v = fminf(fmaxf(x * 32767.0f, -32768.0f), 32767.0f);
    return static_cast<int16_t>(__float2int_rz(v));
}

// Counter-based N(0,1): splitmix64(key + counter) -> Box-Muller.  Same construction as
// oracle/synth_ckpt.py::rng_normal (key = splitmix64(seed*0x100000001B3 + stream)).
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// Counter of the NoiseBlock value of (stream, time step): independent of the decoded length, so that decoding a longer
// prefix of a stream redraws the SAME noise for the samples it shares with the shorter one (streaming policies re-decode
// prefixes: tts_inference_b200/policy.py) and of the batch split.  synth.make_noises_rng mirrors it.
__host__ __device__ __forceinline__ unsigned long long noise_counter(long long stream, int t) {
    return (static_cast<unsigned long long>(stream) << 32) | static_cast<unsigned int>(t);
}
__device__ __forceinline__ float counter_normal(unsigned long long key, unsigned long long ctr) {
    unsigned long long b = splitmix64(key + ctr);
    float u1 = (static_cast<float>(static_cast<unsigned>(b >> 32)) + 0.5f) * (1.0f / 4294967296.0f);
    float u2 = (static_cast<float>(static_cast<unsigned>(b & 0xFFFFFFFFull)) + 0.5f) * (1.0f / 4294967296.0f);
    u1 = fminf(u1, 0.99999994f);
    return sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a thread that owns a 64-byte row segment moves it in two
// requests instead of four -- the row-per-thread epilogues are bound by L1 wavefronts (32 distinct lines per request)
struct alignas(32) U32x8 { uint32_t v[8]; };
__device__ __forceinline__ U32x8 ld_global_v8(const void* p) {
    U32x8 r;
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_global_v8(void* p, const U32x8& r) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
                 : "memory");
}

// ---------------------------------------------------------------- typed load / store of 8 channels
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
}
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
    uint4 raw;
    __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
}
// two fp32 -> packed pair of the 16-bit operand type
__device__ __forceinline__ uint32_t pack2(float a, float b, const __nv_bfloat16*) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack2(float a, float b, const __half*) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ void from_f32(__half& d, float v) { d = __float2half_rn(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f32(float& d, float v) { d = v; }
__device__ __forceinline__ void from_f32(__nv_bfloat16& d, float v) { d = __float2bfloat16_rn(v); }

__device__ __forceinline__ void load32(const __half* p, float (&v)[32]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const U32x8 r = ld_global_v8(p + 16 * h);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r.v[i]));
            v[16 * h + 2 * i] = f.x; v[16 * h + 2 * i + 1] = f.y;
        }
    }
}
__device__ __forceinline__ void load32(const __nv_bfloat16* p, float (&v)[32]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const U32x8 r = ld_global_v8(p + 16 * h);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.v[i]));
            v[16 * h + 2 * i] = f.x; v[16 * h + 2 * i + 1] = f.y;
        }
    }
}
template <typename T>
__device__ __forceinline__ void load32(const T* p, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float t[8];
        load8(p + 8 * i, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * i + j] = t[j];
    }
}
__device__ __forceinline__ void store32(__half* p, const float (&v)[32]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        U32x8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = pack2(v[16 * h + 2 * i], v[16 * h + 2 * i + 1], static_cast<const __half*>(nullptr));
        st_global_v8(p + 16 * h, r);
    }
}
__device__ __forceinline__ void store32(__nv_bfloat16* p, const float (&v)[32]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        U32x8 r;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            r.v[i] = pack2(v[16 * h + 2 * i], v[16 * h + 2 * i + 1], static_cast<const __nv_bfloat16*>(nullptr));
        st_global_v8(p + 16 * h, r);
    }
}
template <typename T>
__device__ __forceinline__ void store32(T* p, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = v[8 * i + j];
        store8(p + 8 * i, t);
    }
}

}  // namespace snacb
