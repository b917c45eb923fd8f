// k_chain: one DecoderBlock's NoiseBlock + three ResidualUnits (+ the next layer's Snake) fused in ONE
// persistent kernel for sm_100a.  A tile of NB*128 time rows of one stream stays on the SM through all
// four layers:
//
//   x1 = y + n[t] * (Wn y)                                   NoiseBlock   (snac layers.py NoiseBlock)
//   x_{l+1} = x_l + W_l snake2(dw_d(snake1(x_l))) + b_l      d = 1, 3, 9  (ResidualUnit)
//   out = snake_next(x_4)
//
//   * the fp32 residual stream lives in TENSOR MEMORY: the NoiseBlock epilogue writes x1 there (tcgen05.st)
//     and tcgen05.mma then accumulates every ResidualUnit's 1x1 conv directly on top of it;
//   * a single 16-bit copy of the tile lives in shared memory (128B-swizzled K-major, the UMMA A-operand
//     layout, filled by TMA).  Its content alternates between
//         S1_l = snake1_l(x_l)        written by the EPILOGUE that drains TMEM after layer l-1's MMAs: the Snake in
//                                     front of the depthwise conv is evaluated in fp32 on the fp32 residual, once per
//                                     element, where the value is in registers anyway (round 2: it used to be
//                                     re-derived from a 16-bit copy inside the prologue, which made the prologue the
//                                     MUFU / FMA-pipe hot spot while both pipes idled through the epilogues), and
//         A_l  = snake2_l(dw_l(S1_l)) written IN PLACE by the PROLOGUE: a 7-tap window slides in registers along one
//                                     dilation class (taps of row r are r + j*d, so a class only ever reads its own
//                                     rows; the 3 rows either side of a warp's span are fetched before a CTA barrier),
//     the MMAs consume A_l block by block, and the epilogue of each 128-row block starts as soon as that block's MMAs
//     have committed;
//   * FOLD (fp16 operands, Snake alphas inside a safe range -- checked when the handle is created, snacb.cu): no alpha
//     multiply is left in the prologue.  The epilogue emits S1'' = alpha1 * snake1(x) = x'' + sin^2 x'' with
//     x'' = alpha1 * x (one FFMA on the accumulator), the depthwise taps carry alpha2 / alpha1, its bias alpha2, so the
//     conv yields a'' = alpha2 * a and snake2 is a'' + sin^2 a'' with 1 / alpha2 folded into the 1x1 weights' K columns.
//     Otherwise (alpha = 0, |alpha| tiny or huge, or bf16 operands) the general form x + (alpha + 1e-9)^-1 sin^2(alpha x)
//     runs in fp32 -- for alpha = 0 that is exactly x, as in the reference;
//   * a tile carries a 40-row halo either side (3*(1+3+9) = 39 rows of receptive field); halo results are
//     garbage by construction and never stored.  Rows outside [0, T) are forced to zero after every layer
//     (the convs' zero padding);
//   * the last epilogue leaves snake_next(x_4) in the tile copy and TMA stores stream it out block by block while
//     the next tile's TMA loads refill the blocks behind them;
//   * HBM traffic: the ConvTranspose output is read once (+ halo), the block output written once.
//
// Earlier variants of this kernel that were measured and dropped (halo exchange through L2, half2-polynomial Snake,
// two warp groups half a layer apart) are kept, uncompiled, under csrc/experiments/ (DESIGN.md section 6).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "chain_span.cuh"
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace snacb {
using namespace ptx;

namespace {

template <typename HT> struct HalfFmtC;
template <> struct HalfFmtC<__half> { static constexpr uint32_t kFmt = 0; };
template <> struct HalfFmtC<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; };

__device__ __forceinline__ float2 unpack2c(uint32_t v, const __half*) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ float2 unpack2c(uint32_t v, const __nv_bfloat16*) {
    // bf16 -> fp32 is a shift: two ALU-pipe instructions, none on the FMA pipe
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<const __half2*>(&v); }
__device__ __forceinline__ uint32_t as_u32(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }

constexpr int kHalo = kChainHalo;            // 40 >= 39, multiple of 8

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2: one issue slot and one FMA-pipe slot for two IEEE fp32 operations, the same
// results as two FFMA): the TMEM-drain epilogues run their scale / bias and sin^2 accumulate on column pairs
__device__ __forceinline__ unsigned long long f2_pack(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& x, float& y) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int C, int NB, bool HALF, bool FOLD>
struct ChainCfg {
    static constexpr int kCH = C / 64;                      // 64-channel K chunks
    static constexpr int kRows = NB * 128;                  // tile rows incl. halo
    static constexpr int kROut = kRows - 2 * kHalo;         // rows stored per tile
    static constexpr int kPlane = NB * 16384;               // one chunk plane of the tile [NB][128 rows][128 B]
    static constexpr int kXBytes = kCH * kPlane;
    static constexpr bool kWRes = (C == 64);                // all four 1x1 weights resident
    static constexpr bool kWSingle = (C == 128);            // ONE layer resident: the next layer's weights are fetched into
                                                            // the same buffer as soon as this layer's MMAs have retired and
                                                            // land during the epilogue + prologue that follow (two CTAs per SM:
                                                            // 32 KB instead of 64 KB of weights per CTA)
    static constexpr bool kWChunked = (C == 256);           // weights streamed one 64-channel K chunk at a time
    static constexpr int kWLayer = C * C * 2;               // one layer's weights [kCH][C rows][128 B]
    static constexpr int kWChunk = C * 128;                 // one K chunk of them
    static constexpr int kWBytes = kWRes ? 4 * kWLayer : (kWChunked ? 2 * kWChunk : (kWSingle ? kWLayer : 2 * kWLayer));
    // prologue parameters per layer and channel pair:
    //   FOLD        8 words : half2 alpha2 * dw bias, 7 x half2 dw tap * alpha2 / alpha1
    //   fp16        12 words: half2 dw bias, 7 x half2 dw tap, float2 alpha2, float2 1 / (alpha2 + 1e-9)
    //   bf16        20 words: 7 x float2 dw tap, float2 dw bias, float2 alpha2, float2 1 / (alpha2 + 1e-9)
    static constexpr int kPrmWords = FOLD ? 8 : (HALF ? 12 : 20);
    static constexpr int kPrmBytes = 3 * (C / 2) * kPrmWords * 4;
    static constexpr int kEpiBytes = 12 * C * 4;            // epilogue vectors: 4 layer boundaries x 3, see the kernel
    static constexpr int kNzBytes = kRows * 4;              // noise value of every tile row
    static constexpr int kOffX = 0;
    static constexpr int kOffW = kOffX + kXBytes;
    static constexpr int kOffPrm = kOffW + kWBytes;
    static constexpr int kOffEpi = kOffPrm + kPrmBytes;
    static constexpr int kOffNz = kOffEpi + kEpiBytes;
    static constexpr int kSpanWarps = (C == 256) ? kChainWarps : 8;         // warps of the launch configuration (kNW*)
    static constexpr int kSpanElems = 3 * kSpanWarps * kChainSpans;
    static constexpr int kSpanBytes = 3 * kSpanElems * 4;   // the launch's span tables (halo-top, carry-top, short last tile), packed
    static constexpr int kROutC = kRows - kHalo;            // rows a carry-top tile owns (no halo above them)
    static constexpr int kBarBytes = 256;
    static constexpr int kOffSpan = kOffNz + kNzBytes;
    static constexpr int kOffBar = kOffSpan + kSpanBytes;
    static constexpr int kSmem = kOffBar + kBarBytes + 1024;
    static constexpr int kTmemCols = NB * C;
    static_assert(!FOLD || HALF, "the alpha-folded formulation is fp16 only");
    static_assert(NB >= 2 && NB <= 8, "blocks per tile");
    static_assert(kTmemCols == 512 || kTmemCols == 256 || kTmemCols == 128, "TMEM columns");
    static_assert(kSmem <= 232448, "shared memory budget");
    static_assert(kWBytes >= 16384, "tail pre-reads may run up to 90 rows past the tile");
};

// ---------------------------------------------------------------------------------------------------------
// One span of the in-place prologue: rows r_oct + k*D, k < 8*nq, of one 64-channel chunk (lane = channel pair).
// The tile copy holds S1 = snake1(x) (FOLD: alpha1 * snake1(x)); the span turns it into the 1x1 conv's operand.
// h0..h2 / t0..t2: the three rows before / after the span (fetched before the barrier, other warps rewrite them).
// ---------------------------------------------------------------------------------------------------------
template <int D, int ROWS, bool FOLD>
__device__ __forceinline__ void span_half(uint8_t* plane, int r_oct, const int nq, const uint32_t h0, const uint32_t h1,
                                          const uint32_t h2, const uint32_t t0, const uint32_t t1, const uint32_t t2,
                                          const uint32_t (&swz)[8], const uint32_t* prm, const int k0) {
    const uint4 q0 = *reinterpret_cast<const uint4*>(prm), q1 = *reinterpret_cast<const uint4*>(prm + 4);
    const __half2 bd = as_h2(q0.x);
    const __half2 w[7] = {as_h2(q0.y), as_h2(q0.z), as_h2(q0.w), as_h2(q1.x), as_h2(q1.y), as_h2(q1.z), as_h2(q1.w)};
    float2 al2 = make_float2(0.f, 0.f), ia2 = make_float2(0.f, 0.f);
    if (!FOLD) {
        const uint4 q2 = *reinterpret_cast<const uint4*>(prm + 8);
        al2 = make_float2(__uint_as_float(q2.x), __uint_as_float(q2.y));
        ia2 = make_float2(__uint_as_float(q2.z), __uint_as_float(q2.w));
    }
    uint32_t xs[14];                               // xs[i] = S1 of the class row (i - 3) steps from the octet's first row
    xs[0] = h0; xs[1] = h1; xs[2] = h2;
    uint8_t* ob = plane + r_oct * 128;             // r_oct = 0 (mod 8): (row & 7) of step k is (k*D) & 7
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uint32_t raw = 0u;                         // class starts may lie above the tile
        if (r_oct + j * D >= 0) raw = *reinterpret_cast<const uint32_t*>(ob + j * D * 128 + swz[(j * D) & 7]);
        xs[3 + j] = raw;
    }
    int qo = 0;
    if (k0 >= 0) {
        // Top span of a carry-top tile, first octet (kept out of the loop below, whose code must not change): the class
        // starts k0 steps above row 0.  Rows above the tile read as 0, except the three class rows right above the first
        // row inside it, which the previous tile left in the carry (h0..h2)
#pragma unroll
        for (int k = 0; k < 8; ++k)
            xs[6 + k] = (r_oct + (k + 3) * D >= 0)
                            ? *reinterpret_cast<const uint32_t*>(ob + (k + 3) * D * 128 + swz[((k + 3) * D) & 7]) : 0u;
        if (k0 > 0) {
#pragma unroll
            for (int i = 0; i < 11; ++i)
                xs[i] = (i == k0) ? h0 : ((i == k0 + 1) ? h1 : ((i == k0 + 2) ? h2 : (i < 3 ? 0u : xs[i])));
        }
        if (nq == 1) { xs[11] = t0; xs[12] = t1; xs[13] = t2; }
        uint32_t o[8];
        dw_snake_half<8, FOLD>(xs, w, bd, al2, ia2, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = r_oct + k * D;
            if (static_cast<unsigned>(r) < static_cast<unsigned>(ROWS))
                *reinterpret_cast<uint32_t*>(ob + k * D * 128 + swz[(k * D) & 7]) = o[k];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) xs[i] = xs[8 + i];
        r_oct += 8 * D;
        ob += 8 * D * 128;
        qo = 1;
    }
#pragma unroll 1
    for (; qo < nq; ++qo) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            xs[6 + k] = *reinterpret_cast<const uint32_t*>(ob + (k + 3) * D * 128 + swz[((k + 3) * D) & 7]);
        if (qo == nq - 1) { xs[11] = t0; xs[12] = t1; xs[13] = t2; }
        uint32_t o[8];
        dw_snake_half<8, FOLD>(xs, w, bd, al2, ia2, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = r_oct + k * D;
            if (static_cast<unsigned>(r) < static_cast<unsigned>(ROWS))
                *reinterpret_cast<uint32_t*>(ob + k * D * 128 + swz[(k * D) & 7]) = o[k];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) xs[i] = xs[8 + i];
        r_oct += 8 * D;
        ob += 8 * D * 128;
    }
}

// same span, fp32 math (bf16 operands)
template <int D, int ROWS>
__device__ __forceinline__ void span_bf16(uint8_t* plane, int r_oct, const int nq, const uint32_t h0, const uint32_t h1,
                                          const uint32_t h2, const uint32_t t0, const uint32_t t1, const uint32_t t2,
                                          const uint32_t (&swz)[8], const uint32_t* prm, const int k0) {
    const float4* p4 = reinterpret_cast<const float4*>(prm);
    const float4 q0 = p4[0], q1 = p4[1], q2 = p4[2], q3 = p4[3], q4 = p4[4];
    const float2 w[7] = {make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y), make_float2(q1.z, q1.w),
                         make_float2(q2.x, q2.y), make_float2(q2.z, q2.w), make_float2(q3.x, q3.y)};
    const float2 bd = make_float2(q3.z, q3.w);
    const float2 al2 = make_float2(q4.x, q4.y), ia2 = make_float2(q4.z, q4.w);
    uint32_t xs[14];
    xs[0] = h0; xs[1] = h1; xs[2] = h2;
    uint8_t* ob = plane + r_oct * 128;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uint32_t raw = 0u;
        if (r_oct + j * D >= 0) raw = *reinterpret_cast<const uint32_t*>(ob + j * D * 128 + swz[(j * D) & 7]);
        xs[3 + j] = raw;
    }
    int qo = 0;
    if (k0 >= 0) {                                 // top span of a carry-top tile: see span_half
#pragma unroll
        for (int k = 0; k < 8; ++k)
            xs[6 + k] = (r_oct + (k + 3) * D >= 0)
                            ? *reinterpret_cast<const uint32_t*>(ob + (k + 3) * D * 128 + swz[((k + 3) * D) & 7]) : 0u;
        if (k0 > 0) {
#pragma unroll
            for (int i = 0; i < 11; ++i)
                xs[i] = (i == k0) ? h0 : ((i == k0 + 1) ? h1 : ((i == k0 + 2) ? h2 : (i < 3 ? 0u : xs[i])));
        }
        if (nq == 1) { xs[11] = t0; xs[12] = t1; xs[13] = t2; }
        uint32_t o[8];
        dw_snake_bf16<8>(xs, w, bd, al2, ia2, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = r_oct + k * D;
            if (static_cast<unsigned>(r) < static_cast<unsigned>(ROWS))
                *reinterpret_cast<uint32_t*>(ob + k * D * 128 + swz[(k * D) & 7]) = o[k];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) xs[i] = xs[8 + i];
        r_oct += 8 * D;
        ob += 8 * D * 128;
        qo = 1;
    }
#pragma unroll 1
    for (; qo < nq; ++qo) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            xs[6 + k] = *reinterpret_cast<const uint32_t*>(ob + (k + 3) * D * 128 + swz[((k + 3) * D) & 7]);
        if (qo == nq - 1) { xs[11] = t0; xs[12] = t1; xs[13] = t2; }
        uint32_t o[8];
        dw_snake_bf16<8>(xs, w, bd, al2, ia2, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = r_oct + k * D;
            if (static_cast<unsigned>(r) < static_cast<unsigned>(ROWS))
                *reinterpret_cast<uint32_t*>(ob + k * D * 128 + swz[(k * D) & 7]) = o[k];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) xs[i] = xs[8 + i];
        r_oct += 8 * D;
        ob += 8 * D * 128;
    }
}

enum { EPI_C_NOISE = 0, EPI_C_MID = 1, EPI_C_FINAL = 2 };

// snake_h2_poly / kEpiPolyMask (the epilogues' FMA-pipe share of sin^2): chain_span.cuh, shared with kernels_chain_ws.cu

}  // namespace

// NW symmetric warps (prologue + epilogue); thread 0 also issues TMA / MMA.  16 warps: one CTA per SM; 8 warps: two.
// JIT: the race-detector build of the same kernel (per-warp delays after the prologue's barrier, see below); the product
// instantiations carry none of it
template <int C, int NB, int NW, typename HT, bool FOLD, bool JIT = false>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1)
k_chain(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmOe,
        const __grid_constant__ CUtensorMap tmOm, const __grid_constant__ CUtensorMap tmWn,
        const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
        const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ ChainArgs a, const int num_strips) {
    constexpr bool kHalfMath = std::is_same<HT, __half>::value;
    using Cfg = ChainCfg<C, NB, kHalfMath, FOLD>;
    constexpr int CH = Cfg::kCH;
    constexpr int kThreads = NW * 32;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem + Cfg::kOffX;
    uint8_t* sW = smem + Cfg::kOffW;
    uint32_t* sPrm = reinterpret_cast<uint32_t*>(smem + Cfg::kOffPrm);
    float* sEpi = reinterpret_cast<float*>(smem + Cfg::kOffEpi);
    float* sNz = reinterpret_cast<float*>(smem + Cfg::kOffNz);
    uint32_t* sSpan = reinterpret_cast<uint32_t*>(smem + Cfg::kOffSpan);     // [3 tables][3][kSpanWarps][kChainSpans]: r_first | n_oct << 16 | kc << 24 | top << 28
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
    uint64_t* ld_bar = bars;          // tile landed (TMA)
    uint64_t* w_bar = bars + 1;       // [2] weight buffers landed
    uint64_t* mma_bar = bars + 3;     // [NB] the layer's MMAs of one 128-row block complete
    uint64_t* wfree_bar = bars + 3 + NB;   // [2] chunked weights: the MMAs reading a buffer have retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + NB);
    volatile int* s_next = reinterpret_cast<volatile int*>(tmem_slot + 2);   // [2] next tile of this CTA, by tile parity
    volatile int* s_cfg = s_next + 2;   // [2] the tile's type (see the tile loop): rows_e | own_end << 11 | halo_top << 22 |
                                        // table << 23 | carry_out << 25; re-read where it is used instead of living in
                                        // registers through the whole tile

    const long long t_kernel0 = clock64();
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int strip_rows = Cfg::kROut + (a.strip_tiles - 1) * Cfg::kROutC;     // rows a full strip owns
    const CUtensorMap* wmaps[4] = {&tmWn, &tmW0, &tmW1, &tmW2};

    // ------------------------------------------------------------------ one-time setup
    if (tid == 0) {
        prefetch_tmap(&tmY); prefetch_tmap(&tmOe); prefetch_tmap(&tmOm);
        prefetch_tmap(&tmWn); prefetch_tmap(&tmW0); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
        mbar_init(ld_bar, 1); mbar_init(&w_bar[0], 1); mbar_init(&w_bar[1], 1);
        for (int b = 0; b < NB; ++b) mbar_init(&mma_bar[b], 1);
        mbar_init(&wfree_bar[0], 1); mbar_init(&wfree_bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
    for (int i = tid; i < 3 * (C / 2); i += kThreads) {            // per-layer prologue parameters of one channel pair
        const int l = i / (C / 2), ch = 2 * (i % (C / 2));
        const ChainLayer& L = a.res[l];
        uint32_t* d = sPrm + i * Cfg::kPrmWords;
        if (FOLD) {
            const float a1x = L.alpha1[ch], a1y = L.alpha1[ch + 1], a2x = L.alpha2[ch], a2y = L.alpha2[ch + 1];
            d[0] = as_u32(__floats2half2_rn(L.dw_b[ch] * a2x, L.dw_b[ch + 1] * a2y));
#pragma unroll
            for (int j = 0; j < 7; ++j)
                d[1 + j] = as_u32(__floats2half2_rn(L.dw_w[j * C + ch] * (a2x / a1x), L.dw_w[j * C + ch + 1] * (a2y / a1y)));
        } else if (kHalfMath) {
            d[0] = as_u32(__floats2half2_rn(L.dw_b[ch], L.dw_b[ch + 1]));
#pragma unroll
            for (int j = 0; j < 7; ++j) d[1 + j] = as_u32(__floats2half2_rn(L.dw_w[j * C + ch], L.dw_w[j * C + ch + 1]));
            d[8] = __float_as_uint(L.alpha2[ch]); d[9] = __float_as_uint(L.alpha2[ch + 1]);
            d[10] = __float_as_uint(L.inv2[ch]); d[11] = __float_as_uint(L.inv2[ch + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                d[2 * j] = __float_as_uint(L.dw_w[j * C + ch]); d[2 * j + 1] = __float_as_uint(L.dw_w[j * C + ch + 1]);
            }
            d[14] = __float_as_uint(L.dw_b[ch]); d[15] = __float_as_uint(L.dw_b[ch + 1]);
            d[16] = __float_as_uint(L.alpha2[ch]); d[17] = __float_as_uint(L.alpha2[ch + 1]);
            d[18] = __float_as_uint(L.inv2[ch]); d[19] = __float_as_uint(L.inv2[ch + 1]);
        }
    }
    // the span table is indexed by (layer, warp) at run time: from shared memory, not from the constant bank (a dynamically
    // indexed kernel parameter costs a constant-cache miss per layer: ~1 k cycles of the 'pre' phase, measured)
    static_assert(NW <= Cfg::kSpanWarps, "span table");
    for (int i = tid; i < Cfg::kSpanElems; i += kThreads) {
        const int l_ = i / (Cfg::kSpanWarps * kChainSpans), w_ = (i / kChainSpans) % Cfg::kSpanWarps, k_ = i % kChainSpans;
        auto pk = [](const ChainSpan& f) {
            return (static_cast<uint32_t>(f.r_first) & 0xFFFFu) | (static_cast<uint32_t>(f.n_oct) << 16) |
                   (static_cast<uint32_t>(f.kc) << 24) | (static_cast<uint32_t>(f.pad) << 28);
        };
        sSpan[i] = pk(a.spans[l_][w_][k_]);
        sSpan[Cfg::kSpanElems + i] = pk(a.spans_carry[l_][w_][k_]);
        sSpan[2 * Cfg::kSpanElems + i] = pk(a.spans_last[l_][w_][k_]);
    }
    // epilogue vectors, three per layer boundary i (0: NoiseBlock -> unit d=1, 1: d=1 -> d=3, 2: d=3 -> d=9, 3: d=9 -> out):
    //   [3i]     bias: sum of the 1x1 biases so far (0 for i = 0); FOLD, i < 3: that bias times alpha1 of the coming unit
    //   [3i + 1] alpha of the Snake the epilogue applies (the coming unit's snake1, the next layer's Snake for i = 3)
    //   [3i + 2] 1 / (alpha + 1e-9) of it (unused by FOLD for i < 3)
    for (int c = tid; c < C; c += kThreads) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float b = i == 0 ? 0.f : a.bias_cum[(i - 1) * C + c];
            const float al = i < 3 ? a.res[i].alpha1[c] : a.alpha_next[c];
            const float ia = i < 3 ? a.res[i].inv1[c] : a.inv_next[c];
            sEpi[(3 * i) * C + c] = (FOLD && i < 3) ? b * al : b;
            sEpi[(3 * i + 1) * C + c] = al;
            sEpi[(3 * i + 2) * C + c] = ia;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // weight staging: resident (C = 64) or two rotating buffers, layer l of a tile uses buffer l & 1
    auto load_w = [&](int l, int buf) {       // thread 0
        uint8_t* dst = sW + (Cfg::kWRes ? l : (Cfg::kWSingle ? 0 : buf)) * Cfg::kWLayer;
        uint64_t* bar = &w_bar[(Cfg::kWRes || Cfg::kWSingle) ? 0 : buf];
        if (!Cfg::kWRes) mbar_expect_tx(bar, Cfg::kWLayer);
#pragma unroll
        for (int kc = 0; kc < CH; ++kc) tma_load_2d_hint(dst + kc * (C * 128), wmaps[l], kc * 64, 0, bar, kL2EvictLast);
    };
    // chunked weights (C = 256): chunk G of the CTA's running sequence = layer (G / CH) % 4, K chunk G % CH, buffer G & 1
    auto load_wc = [&](int G) {               // thread 0
        const int buf = G & 1;
        mbar_expect_tx(&w_bar[buf], Cfg::kWChunk);
        tma_load_2d_hint(sW + buf * Cfg::kWChunk, wmaps[(G / CH) & 3], (G % CH) * 64, 0, &w_bar[buf], kL2EvictLast);
    };
    // a tile = (strip << 5) | j: tile j of a strip; strip = stream * sps + the strip's index inside the stream's row range
    auto tile_coords = [&](int tile, int& s, int& t_start) {
        const int strip = tile >> 5, j = tile & 31;
        s = strip / a.sps;
        t_start = (a.t_n > 0 ? a.t_lo : 0) + sm_off(a.map, s, a.rpf) + (strip - s * a.sps) * strip_rows +
                  (j == 0 ? -kHalo : Cfg::kROut + (j - 1) * Cfg::kROutC);
    };
    auto load_block = [&](int s, int t_start, int b) {      // thread 0; ld_bar's expect_tx covers the whole tile
        const int sl = sm_slot(a.map, s);
#pragma unroll
        for (int kc = 0; kc < CH; ++kc)
            tma_load_3d_hint(sX + kc * Cfg::kPlane + b * 16384, &tmY, kc * 64, t_start + b * 128, sl, ld_bar, kL2EvictFirst);
    };
    int tile = static_cast<int>(blockIdx.x) << 5;
    if (tid == 0 && (tile >> 5) < num_strips) {
        if (Cfg::kWRes) {
            mbar_expect_tx(&w_bar[0], 4 * Cfg::kWLayer);
            for (int l = 0; l < 4; ++l) load_w(l, 0);
        } else if (Cfg::kWChunked) {
            load_wc(0);
            load_wc(1);
        } else if (Cfg::kWSingle) {
            load_w(0, 0);
        } else {
            load_w(0, 0);
            load_w(1, 1);
        }
        int s, t_start;
        tile_coords(tile, s, t_start);
        mbar_expect_tx(ld_bar, CH * Cfg::kPlane);
        for (int b = 0; b < NB; ++b) load_block(s, t_start, b);
    }

    uint32_t mma_par = 0;
    constexpr uint32_t idescW = umma_idesc_f16(128, C, HalfFmtC<HT>::kFmt);

    // swizzled byte offset of this lane's channel pair inside a 128-byte row, for each (row & 7)
    uint32_t swz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) swz[j] = ((((lane >> 2) ^ j) & 7) << 4) + ((lane & 3) << 2);
    const uint32_t sx_addr = smem_u32(sX);

    // issue one layer's 1x1 conv for the whole tile (thread 0): TMEM[blk] (+)= A[blk] * W^T, one commit per block
    auto issue_layer = [&](int l, int n, bool has_next, int nb_live) {
        if (Cfg::kWChunked) {
            // K chunk by K chunk through two weight buffers; a buffer is refilled (two chunks ahead) as soon as the
            // MMAs reading it have retired.  All blocks complete with the last chunk.
            const int gbase = (n * 4 + l) * CH;
#pragma unroll
            for (int kc = 0; kc < CH; ++kc) {
                const int G = gbase + kc, buf = G & 1;
                mbar_wait(&w_bar[buf], (G >> 1) & 1);
                tc_fence_after();
                const uint32_t w_addr = smem_u32(sW + buf * Cfg::kWChunk);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if (b >= nb_live) continue;            // short last tile: blocks past its right halo
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_f16_ss(tmem_base + b * C, umma_desc_sw128(sx_addr + kc * Cfg::kPlane + b * 16384 + k * 32),
                                   umma_desc_sw128(w_addr + k * 32), idescW, (l > 0 || kc > 0 || k > 0) ? 1u : 0u);
                }
                mma_commit(&wfree_bar[buf]);
                if (kc & 1) {
                    const bool more = (kc + 1 < CH) || (l < 3) || has_next;
                    if (more) {
                        mbar_wait(&wfree_bar[0], ((G - 1) >> 1) & 1); load_wc(G + 1);
                        mbar_wait(&wfree_bar[1], (G >> 1) & 1); load_wc(G + 2);
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) mma_commit(&mma_bar[b]);
            return;
        }
        const int buf = l & 1;
        if (Cfg::kWRes) { if (n == 0 && l == 0) mbar_wait(&w_bar[0], 0); }
        else if (Cfg::kWSingle) mbar_wait(&w_bar[0], (4 * n + l) & 1);       // one load per layer into the one buffer
        else mbar_wait(&w_bar[buf], (2 * n + (l >> 1)) & 1);
        tc_fence_after();
        const uint32_t w_addr = smem_u32(sW + (Cfg::kWRes ? l : (Cfg::kWSingle ? 0 : buf)) * Cfg::kWLayer);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b < nb_live) {                             // short last tile: blocks past its right halo carry no MMAs
#pragma unroll
                for (int kc = 0; kc < CH; ++kc)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_f16_ss(tmem_base + b * C, umma_desc_sw128(sx_addr + kc * Cfg::kPlane + b * 16384 + k * 32),
                                   umma_desc_sw128(w_addr + kc * (C * 128) + k * 32), idescW,
                                   (l > 0 || kc > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit(&mma_bar[b]);                       // every block's barrier completes a phase per layer
        }
    };
    // after ALL of layer l's MMAs completed: its weight buffer is free -> prefetch the layer two ahead (thread 0);
    // single buffer: the NEXT layer, which lands while the epilogue and the next prologue run
    auto prefetch_w = [&](int l, bool has_next) {
        if (Cfg::kWRes || Cfg::kWChunked) return;
        if (Cfg::kWSingle) {
            if (l + 1 < 4 || has_next) load_w((l + 1) & 3, 0);
            return;
        }
        const int l2 = (l + 2) & 3;
        if (l + 2 < 4 || has_next) load_w(l2, l & 1);
    };

    // epilogue after layer boundary `bnd`; each 128-row block is drained as soon as its MMAs have committed
    //   NOISE (bnd 0): x1 = y + n[t] * TMEM -> TMEM (fp32 residual stream); tile copy = snake1_0(x1)
    //   MID (bnd 1, 2): tile copy = snake1_bnd(TMEM + cumulative bias)
    //   FINAL (bnd 3):  tile copy = snake_next(TMEM + cumulative bias)
    auto epilogue = [&](auto mode_tag, const int bnd, int t_start, const int cfg) {
        const int rows_e = cfg & 0x7FF, own_end = (cfg >> 11) & 0x7FF, own_begin = ((cfg >> 22) & 1) ? kHalo : 0;
        constexpr int MODE = decltype(mode_tag)::value;
        constexpr bool kFoldHere = FOLD && MODE != EPI_C_FINAL;
        const float* vb = sEpi + (3 * bnd) * C;            // bias (FOLD: scaled bias)
        const float* va = vb + C;                          // alpha (FOLD: the scale)
        const float* vi = va + C;                          // 1 / (alpha + 1e-9)
        const int q = warp & 3, g = warp >> 2;
        constexpr int kPieces = NB * (C / 32);
        const HT* tag = nullptr;
#pragma unroll 1
        for (int it = g; it < kPieces; it += NW / 4) {
            const int blk = it / (C / 32), cg = it % (C / 32);
            // the last epilogue only feeds the TMA stores: 32-row groups entirely inside the halo are skipped
            if (MODE == EPI_C_FINAL && (blk * 128 + q * 32 + 32 <= own_begin || blk * 128 + q * 32 >= own_end)) continue;
            if (blk * 128 + q * 32 >= rows_e) continue;        // short last tile: rows past its right halo
            mbar_wait(&mma_bar[blk], mma_par);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + blk * C + cg * 32;
            uint32_t raw[32];
            tmem_ld32(taddr, raw);
            const int i = blk * 128 + q * 32 + lane;
            const int t = t_start + i;
            const bool valid = static_cast<unsigned>(t) < static_cast<unsigned>(a.T);
            uint8_t* row = sX + (cg >> 1) * Cfg::kPlane + i * 128;
            uint4 yv[4];
            float nz = 0.f;
            if (MODE == EPI_C_NOISE) {
#pragma unroll
                for (int c = 0; c < 4; ++c) yv[c] = *reinterpret_cast<const uint4*>(row + ((((cg & 1) * 4 + c) ^ (i & 7)) << 4));
                nz = sNz[i];
            }
            tmem_ld_wait();
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float v0 = __uint_as_float(raw[j]), v1 = __uint_as_float(raw[j + 1]);
                float v2 = __uint_as_float(raw[j + 2]), v3 = __uint_as_float(raw[j + 3]);
                if (MODE == EPI_C_NOISE) {
                    const uint32_t* yw = reinterpret_cast<const uint32_t*>(yv);
                    const float2 y0 = unpack2c(yw[j / 2], tag), y1 = unpack2c(yw[j / 2 + 1], tag);
                    const unsigned long long nz2 = f2_pack(nz, nz);
                    f2_unpack(ffma2(nz2, f2_pack(v0, v1), f2_pack(y0.x, y0.y)), v0, v1);
                    f2_unpack(ffma2(nz2, f2_pack(v2, v3), f2_pack(y1.x, y1.y)), v2, v3);
                    raw[j] = __float_as_uint(v0); raw[j + 1] = __float_as_uint(v1);
                    raw[j + 2] = __float_as_uint(v2); raw[j + 3] = __float_as_uint(v3);
                }
                const float4 al = *reinterpret_cast<const float4*>(va + cg * 32 + j);
                if (kFoldHere) {
                    // x'' = alpha1 * (x + b);  S1'' = x'' + sin^2 x''   (column pairs: FFMA2, bit-identical to two FFMA)
                    unsigned long long p01 = f2_pack(v0, v1), p23 = f2_pack(v2, v3);
                    if (MODE == EPI_C_NOISE) { p01 = fmul2(p01, f2_pack(al.x, al.y)); p23 = fmul2(p23, f2_pack(al.z, al.w)); }
                    else {
                        const float4 b = *reinterpret_cast<const float4*>(vb + cg * 32 + j);
                        p01 = ffma2(p01, f2_pack(al.x, al.y), f2_pack(b.x, b.y));
                        p23 = ffma2(p23, f2_pack(al.z, al.w), f2_pack(b.z, b.w));
                    }
                    // The epilogue is bound by the XU pipe (32 MUFU.SIN per 32 x 32 piece = 256 cycles per sub-partition
                    // against ~160 issue slots): kEpiPolyMask moves some column pairs' sin^2 to the FMA pipe as a packed
                    // half2 polynomial, until the two pipes balance
                    f2_unpack(p01, v0, v1); f2_unpack(p23, v2, v3);
                    if ((kEpiPolyMask >> (j / 2)) & 1u) {
                        o[j / 2] = as_u32(snake_h2_poly(__floats2half2_rn(v0, v1)));
                    } else {
                        const float s0 = __sinf(v0), s1 = __sinf(v1);
                        const unsigned long long s01 = f2_pack(s0, s1);
                        f2_unpack(ffma2(s01, s01, p01), v0, v1);
                        o[j / 2] = pack2(v0, v1, tag);
                    }
                    if ((kEpiPolyMask >> (j / 2 + 1)) & 1u) {
                        o[j / 2 + 1] = as_u32(snake_h2_poly(__floats2half2_rn(v2, v3)));
                    } else {
                        const float s2 = __sinf(v2), s3 = __sinf(v3);
                        const unsigned long long s23 = f2_pack(s2, s3);
                        f2_unpack(ffma2(s23, s23, p23), v2, v3);
                        o[j / 2 + 1] = pack2(v2, v3, tag);
                    }
                    continue;
                } else {
                    if (MODE != EPI_C_NOISE) {
                        const float4 b = *reinterpret_cast<const float4*>(vb + cg * 32 + j);
                        v0 += b.x; v1 += b.y; v2 += b.z; v3 += b.w;
                    }
                    const float4 ia = *reinterpret_cast<const float4*>(vi + cg * 32 + j);
                    v0 = snake_f<true>(v0, al.x, ia.x); v1 = snake_f<true>(v1, al.y, ia.y);
                    v2 = snake_f<true>(v2, al.z, ia.z); v3 = snake_f<true>(v3, al.w, ia.w);
                }
                o[j / 2] = pack2(v0, v1, tag);
                o[j / 2 + 1] = pack2(v2, v3, tag);
            }
            if (MODE == EPI_C_NOISE) tmem_st32(taddr, raw);
            // rows outside [0, T) are the convs' zero padding: only the tiles at a stream's two ends have any, so the
            // selects are kept out of the common path (warp-uniform branch)
            if (MODE == EPI_C_FINAL || __all_sync(0xFFFFFFFFu, valid)) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int chunk = ((cg & 1) * 4 + c) ^ (i & 7);
                    *reinterpret_cast<uint4*>(row + chunk * 16) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int chunk = ((cg & 1) * 4 + c) ^ (i & 7);
                    *reinterpret_cast<uint4*>(row + chunk * 16) =
                        valid ? make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]) : make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
        if (MODE == EPI_C_NOISE) tmem_st_wait();
    };

    long long t_last = clock64();
    auto tick = [&](int slot) {
        if (a.prof != nullptr && tid == 0 && blockIdx.x == 0) {
            const long long t = clock64();
            a.prof[slot] += static_cast<unsigned long long>(t - t_last);
            t_last = t;
        }
    };

    for (int n = 0; (tile >> 5) < num_strips; ++n) {
        int s, t_start;
        tile_coords(tile, s, t_start);
        // claim the tile after this one (persistent CTAs, dynamic order: tiles cost the same but SMs do not run alike)
        if (tid == 0) {
            // Tile types.  The first tile of a strip is a HALO-TOP tile: 40 rows of context above its owned rows are
            // recomputed (and garbage towards the top, never stored).  Every further tile of the strip is a CARRY-TOP tile:
            // it owns its rows from row 0 on -- the three class rows above each dilation class come from the previous tile
            // through the carry -- so only the halo BELOW a tile is computed twice.  The last tile of a stream's last strip
            // may own fewer rows (a.last_rows): it runs the shorter span schedule, drains / rewrites rows [0, rows_e) only
            // (rows_e covers its right halo, in whole 32-row pieces) and issues MMAs for the blocks that hold them.
            // Every row any tile computes goes through the same arithmetic on the same inputs.
            const int strip = tile >> 5, j = tile & 31;
            const int si = strip % a.sps;
            const int ntl = (si == a.sps - 1) ? a.last_strip_tiles : a.strip_tiles;
            s_next[n & 1] = (j + 1 < ntl) ? tile + 1 : ((static_cast<int>(gridDim.x) + atomicAdd(a.tile_counter, 1)) << 5);
            const bool short_t = a.last_rows > 0 && si == a.sps - 1 && j == ntl - 1;
            const int own_begin = j == 0 ? kHalo : 0;
            const int own_end = short_t ? own_begin + a.last_rows : Cfg::kRows - kHalo;
            int rows_e = short_t ? ((own_end + kHalo + 31) & ~31) : Cfg::kRows;
            if (rows_e > Cfg::kRows) rows_e = Cfg::kRows;
            s_cfg[n & 1] = rows_e | (own_end << 11) | ((j == 0 ? 1 : 0) << 22) | ((short_t ? 2 : (j == 0 ? 0 : 1)) << 23) |
                           ((j + 1 < ntl ? 1 : 0) << 25);
        }

        // ---------------------------------------------------------------- noise values (overlaps the tile load)
        {
            unsigned long long key = 0;
            if (a.noise == nullptr)
                key = splitmix64(a.seed * 0x100000001B3ull + static_cast<unsigned long long>(100 + a.noise_stage));
            for (int i = tid; i < Cfg::kRows; i += kThreads) {
                const int t = t_start + i;
                float v = 0.f;
                if (static_cast<unsigned>(t) < static_cast<unsigned>(a.T))
                    v = a.noise ? a.noise[static_cast<size_t>(s) * a.T + t]
                                : counter_normal(key, noise_counter(a.stream_keys ? a.stream_keys[s] : a.stream_offset + s, t + a.t0 + sm_org(a.map, s, a.rpf)));
                sNz[i] = v;
            }
        }
        mbar_wait(ld_bar, n & 1);
        __syncthreads();
        const int next_tile = s_next[n & 1];
        const bool has_next = (next_tile >> 5) < num_strips;
        tick(0);
        // ---------------------------------------------------------------- NoiseBlock: TMEM = Wn y, then x1 = y + n TMEM
        if (tid == 0) issue_layer(0, n, has_next, ((s_cfg[n & 1] & 0x7FF) + 127) >> 7);
        epilogue(std::integral_constant<int, EPI_C_NOISE>{}, 0, t_start, s_cfg[n & 1]);
        if (tid == 0) { mbar_wait(&mma_bar[NB - 1], mma_par); prefetch_w(0, has_next); }
        mma_par ^= 1u;
        tc_fence_before();
        __syncthreads();
        tick(1);

        // ---------------------------------------------------------------- three ResidualUnits
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            const int d = (l == 0) ? 1 : (l == 1 ? 3 : 9);
            // ---- spans of this warp: pre-read the 3 rows before and after each span (owned by other warps)
            uint32_t hd[kChainSpans][3], tl[kChainSpans][3];
            int r_first[kChainSpans], n_oct[kChainSpans], kcs[kChainSpans];
            const int cfg_l = s_cfg[n & 1];
            const uint32_t* spanT = sSpan + ((cfg_l >> 23) & 3) * Cfg::kSpanElems;
            const bool carry_top = ((cfg_l >> 22) & 1) == 0;
            const int c_off = (l == 0) ? 0 : (l == 1 ? 3 : 12);                 // this layer's rows inside the 39-row carry
            auto preread = [&]() {
#pragma unroll
            for (int sp = 0; sp < kChainSpans; ++sp) {
                const uint32_t spw = spanT[(l * Cfg::kSpanWarps + warp) * kChainSpans + sp];
                const ChainSpan spn{static_cast<short>(spw & 0xFFFFu), static_cast<short>((spw >> 16) & 0xFFu), static_cast<short>((spw >> 24) & 0xFu), 0};
                r_first[sp] = spn.r_first; n_oct[sp] = spn.n_oct; kcs[sp] = spn.kc;
#pragma unroll
                for (int j = 0; j < 3; ++j) { hd[sp][j] = 0u; tl[sp][j] = 0u; }
                if (spn.n_oct > 0) {                       // warp-uniform: unused span slots cost one branch
                    const uint8_t* lane_base = sX + spn.kc * Cfg::kPlane + ((lane & 3) << 2);
                    const int c16 = lane >> 2;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int rh = spn.r_first - (3 - j) * d;
                        const int rt = spn.r_first + (8 * spn.n_oct + j) * d;
                        if (rh >= 0) hd[sp][j] = *reinterpret_cast<const uint32_t*>(lane_base + rh * 128 + (((c16 ^ rh) & 7) << 4));
                        tl[sp][j] = *reinterpret_cast<const uint32_t*>(lane_base + rt * 128 + (((c16 ^ rt) & 7) << 4));
                    }
                    // carry-top tile (rare path, kept apart from the loads above): the head rows of a span that lie above
                    // row 0 come from the carry the previous tile of the strip left -- for the FIRST span of a class
                    // (top = 1 + its steps above row 0) the three class rows above its first row inside the tile
                    if (carry_top && spn.r_first - 3 * d < 0) {
                        const int topk = static_cast<int>(spw >> 28);
                        kcs[sp] = spn.kc | (topk << 8);
                        // carry buffers of this CTA by tile parity: this tile reads what the previous one wrote
                        const __half* crd = static_cast<const __half*>(a.carry) +
                                            (static_cast<size_t>(blockIdx.x) * 2 + ((n & 1) ^ 1)) * 39 * C + c_off * C + spn.kc * 64 + 2 * lane;
                        const int r_nn = spn.r_first + (topk > 0 ? topk - 1 : 0) * d;
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const int rh = r_nn - (3 - j) * d;      // topk = 0: r_nn = r_first, only its rows above row 0
                            if (rh < 0) hd[sp][j] = *reinterpret_cast<const uint32_t*>(crd + (rh + 3 * d) * C);
                        }
                    }
                }
            }
            };
            preread();
            if ((cfg_l >> 25) & 1) {
                // the next tile of the strip starts at this tile's row kRows - 40: leave it the 3 d rows of S1 above that
                for (int idx = tid; idx < 3 * d * CH * 8; idx += kThreads) {
                    const int row = idx / (CH * 8), kc = (idx >> 3) % CH, c16 = idx & 7;
                    const int r = Cfg::kRows - kHalo - 3 * d + row;
                    const uint4 v = *reinterpret_cast<const uint4*>(sX + kc * Cfg::kPlane + r * 128 + (((c16 ^ r) & 7) << 4));
                    __half* carry_wr = static_cast<__half*>(a.carry) + (static_cast<size_t>(blockIdx.x) * 2 + (n & 1)) * 39 * C;
                    *reinterpret_cast<uint4*>(carry_wr + (c_off + row) * C + kc * 64 + c16 * 8) = v;
                }
            }
            __syncthreads();
            // Race detector in place of compute-sanitizer racecheck (closed on this pool): with a jitter seed every warp
            // starts its in-place rewrite at a different, pseudo-random time (up to half a prologue apart), so a warp that
            // read a row another warp owns AFTER the barrier would see it rewritten in some runs and not in others; the
            // output must stay bit-identical for every seed (tests/test_gpu_parity.py::test_chain_schedule_survives_jitter).
            // (Carried in the product kernel it cost 3 % of the C = 64 chain: a separate instantiation, selected by the seed.)
            if (JIT && a.jitter != 0u) {
                const unsigned hsh = static_cast<unsigned>(splitmix64((static_cast<unsigned long long>(a.jitter) << 32) ^
                                                                      (static_cast<unsigned long long>(tile) << 8) ^ (warp << 2) ^ l));
                const long long until = clock64() + (hsh & 4095u);
                while (clock64() < until) {}
                // negative control (seed bit 31): fetch the neighbour rows AFTER the barrier and the delay, i.e. commit the
                // very read-after-overwrite the schedule forbids -- the test checks that this DOES change the output
                if (a.jitter & 0x80000000u) preread();
            }
            tick(2 + 4 * l);
#pragma unroll 1
            for (int sp = 0; sp < kChainSpans; ++sp) {
                static_assert(kChainSpans == 4, "span select");
                const int nq = sp == 0 ? n_oct[0] : (sp == 1 ? n_oct[1] : (sp == 2 ? n_oct[2] : n_oct[3]));
                if (nq == 0) continue;
                const int kck = sp == 0 ? kcs[0] : (sp == 1 ? kcs[1] : (sp == 2 ? kcs[2] : kcs[3]));
                const int kc = kck & 0xFF, k0 = (kck >> 8) - 1;           // k0 >= 0: top span of a carry-top tile
                const int r0 = sp == 0 ? r_first[0] : (sp == 1 ? r_first[1] : (sp == 2 ? r_first[2] : r_first[3]));
                uint32_t hh[3], tt[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    hh[j] = sp == 0 ? hd[0][j] : (sp == 1 ? hd[1][j] : (sp == 2 ? hd[2][j] : hd[3][j]));
                    tt[j] = sp == 0 ? tl[0][j] : (sp == 1 ? tl[1][j] : (sp == 2 ? tl[2][j] : tl[3][j]));
                }
                const uint32_t* prm = sPrm + ((l * (C / 2)) + kc * 32 + lane) * Cfg::kPrmWords;
                uint8_t* plane = sX + kc * Cfg::kPlane;
                if (kHalfMath) {
                    if (d == 1) span_half<1, Cfg::kRows, FOLD>(plane, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], swz, prm, k0);
                    else if (d == 3) span_half<3, Cfg::kRows, FOLD>(plane, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], swz, prm, k0);
                    else span_half<9, Cfg::kRows, FOLD>(plane, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], swz, prm, k0);
                } else {
                    if (d == 1) span_bf16<1, Cfg::kRows>(plane, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], swz, prm, k0);
                    else if (d == 3) span_bf16<3, Cfg::kRows>(plane, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], swz, prm, k0);
                    else span_bf16<9, Cfg::kRows>(plane, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], swz, prm, k0);
                }
            }
            tick(3 + 4 * l);
            fence_proxy_async_smem();
            __syncthreads();
            tick(4 + 4 * l);
            if (tid == 0) issue_layer(l + 1, n, has_next, ((s_cfg[n & 1] & 0x7FF) + 127) >> 7);
            if (l < 2) epilogue(std::integral_constant<int, EPI_C_MID>{}, l + 1, t_start, s_cfg[n & 1]);
            else epilogue(std::integral_constant<int, EPI_C_FINAL>{}, 3, t_start, s_cfg[n & 1]);
            if (tid == 0) { mbar_wait(&mma_bar[NB - 1], mma_par); prefetch_w(l + 1, has_next); }
            mma_par ^= 1u;
            if (l == 2) fence_proxy_async_smem();          // the tile copy is the source of the TMA stores below
            tc_fence_before();
            __syncthreads();
            tick(5 + 4 * l);
        }

        // ---------------------------------------------------------------- stream the tile out, refill behind it
        if (tid == 0) {
            const int t_out = t_start + kHalo;
            const int own_end = (s_cfg[n & 1] >> 11) & 0x7FF;
            const bool halo_top = ((s_cfg[n & 1] >> 22) & 1) != 0;
            const int sl = sm_slot(a.map, s);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
#pragma unroll
                for (int kc = 0; kc < CH; ++kc) {
                    if (b * 128 >= own_end) continue;      // short last tile: nothing owned in this block (empty group)
                    const uint8_t* src = sX + kc * Cfg::kPlane + b * 16384;
                    if (b == 0 && halo_top) tma_store_3d(&tmOe, src + kHalo * 128, kc * 64, t_out, sl);
                    else if (b == NB - 1) tma_store_3d(&tmOe, src, kc * 64, t_start + b * 128, sl);
                    else tma_store_3d(&tmOm, src, kc * 64, t_start + b * 128, sl);
                }
                bulk_commit_group();
            }
            if (has_next) {
                int s2, t2;
                tile_coords(next_tile, s2, t2);
                mbar_expect_tx(ld_bar, CH * Cfg::kPlane);
                // block b may be refilled once the store of block b has read it (groups complete in order)
                if (NB > 7) { bulk_wait_group_read<7>(); load_block(s2, t2, NB - 8); }
                if (NB > 6) { bulk_wait_group_read<6>(); load_block(s2, t2, NB - 7); }
                if (NB > 5) { bulk_wait_group_read<5>(); load_block(s2, t2, NB - 6); }
                if (NB > 4) { bulk_wait_group_read<4>(); load_block(s2, t2, NB - 5); }
                if (NB > 3) { bulk_wait_group_read<3>(); load_block(s2, t2, NB - 4); }
                if (NB > 2) { bulk_wait_group_read<2>(); load_block(s2, t2, NB - 3); }
                bulk_wait_group_read<1>(); load_block(s2, t2, NB - 2);
                bulk_wait_group_read<0>(); load_block(s2, t2, NB - 1);
            }
        }
        tick(14);
        tile = next_tile;
    }
    if (tid == 0) bulk_wait_group<0>();
    if (a.prof != nullptr && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.prof[20 + 2 * blockIdx.x] = static_cast<unsigned long long>(clock64() - t_kernel0);
        a.prof[21 + 2 * blockIdx.x] = smid;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

namespace {

template <int C, int NB, int NW, typename HT, bool FOLD, bool JIT = false>
cudaError_t launch_chain_t(const ChainArgs& a, const CUtensorMap* tm, int sm_count, cudaStream_t st) {
    using Cfg = ChainCfg<C, NB, std::is_same<HT, __half>::value, FOLD>;
    static PerDeviceOnce once;
    int dev_;
    if (once.needed(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(k_chain<C, NB, NW, HT, FOLD, JIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_chain<C, NB, NW, HT, FOLD, JIT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        once.done(dev_);
    }
    if (a.strip_tiles < 1 || a.strip_tiles > 31 || a.sps < 1 || a.last_strip_tiles < 1 || a.last_strip_tiles > a.strip_tiles ||
        (a.strip_tiles > 1 && a.carry == nullptr))
        return cudaErrorInvalidValue;
    const long long strips = static_cast<long long>(a.S) * a.sps;
    if (strips == 0) return cudaSuccess;
    if (strips >= (1LL << 25)) return cudaErrorInvalidValue;            // a tile id is (strip << 5) | j
    const int slots = sm_count * (NW == 8 ? 2 : 1);
    const int grid = strips < slots ? static_cast<int>(strips) : slots;
    k_chain<C, NB, NW, HT, FOLD, JIT><<<grid, NW * 32, Cfg::kSmem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], a,
                                                                          static_cast<int>(strips));
    return cudaGetLastError();
}

// launch configurations: C = 64: 512-row tiles, 8 warps, two CTAs per SM (one CTA's barrier / MMA / TMA waits are
// the other's issue slots); C = 128: 256-row tiles, 8 warps, two CTAs per SM as well (64 KB tile copy + ONE layer of
// weights per CTA; with carry-top tiles a 256-row tile owns 216 rows = 84 %, what a 512-row tile with two halos owned);
// C = 256 (fp16 only: the fp32-math parameters of the bf16 variant do not fit): 256-row tiles (TMEM: 2 x 256 columns),
// 16 warps, weights streamed per K chunk
constexpr int kNB64 = 4, kNW64 = 8, kNB128 = 2, kNW128 = 8, kNB256 = 2, kNW256 = 16;

}  // namespace

bool chain_supported(int C, int half_fp16) { return C == 64 || C == 128 || (C == 256 && half_fp16); }
int chain_tile_rows(int C) { return (C == 64 ? kNB64 : (C == 128 ? kNB128 : kNB256)) * 128; }
int chain_warps(int C) { return C == 64 ? kNW64 : (C == 128 ? kNW128 : kNW256); }

// Spans of the in-place prologue (see the header comment): for dilation d the rows of a tile split into d classes
// r = r0 + k d.  Class starts are multiples of 8 (so that the swizzle phase of step k is static) no larger than the
// first row whose result is needed at that layer; negative starts skip their first few steps.
bool chain_build_spans(int C, ChainSpan (*spans)[kChainWarps][kChainSpans], int own_end_in, bool carry_top) {
    const int rows = chain_tile_rows(C), ch = C / 64, nw = chain_warps(C);
    const int own_end = (own_end_in > 0 && own_end_in < rows - kChainHalo) ? own_end_in : rows - kChainHalo;
    static const int dil[3] = {1, 3, 9};
    ChainSpan out[3][kChainWarps][kChainSpans];
    for (int l = 0; l < 3; ++l) {
        const int d = dil[l];
        struct Cls { int kc, r0, noct, k0; };
        std::vector<Cls> cls;
        const int top = (d == 1) ? 0 : (d == 3 ? 8 : 40);
        // rows at or beyond `hi` are not needed downstream: the last unit feeds only the stored rows (< own_end),
        // the unit before it additionally that unit's 27 rows of taps, the first one 9 more
        const int hi = own_end + (l == 2 ? 0 : (l == 1 ? 27 : 36));
        for (int kc = 0; kc < ch; ++kc)
            for (int m = 0; m < d; ++m) {
                int r0 = top - 8 * m, k0 = 0;
                if (carry_top) {
                    // the class of residue m must produce every row from its first one (m) on: start at the largest
                    // multiple of 8 that is <= m and = m (mod d); k0 = its steps above row 0 (nothing is stored for them)
                    r0 = m;
                    while (r0 & 7) { r0 -= d; ++k0; }
                }
                const int steps = (hi - r0 + d - 1) / d;
                if (steps <= 0) continue;
                cls.push_back({kc, r0, (steps + 7) / 8, k0});
            }
        int total = 0;
        for (auto& c : cls) total += c.noct;
        for (int w = 0; w < kChainWarps; ++w) {
            for (int k = 0; k < kChainSpans; ++k) out[l][w][k] = ChainSpan{0, 0, 0, 0};
            if (w >= nw) continue;
            const int lo = static_cast<int>(static_cast<long long>(w) * total / nw);
            const int hi = static_cast<int>(static_cast<long long>(w + 1) * total / nw);
            int nsp = 0, base = 0;
            for (auto& c : cls) {
                const int a0 = lo > base ? lo : base, a1 = hi < base + c.noct ? hi : base + c.noct;
                if (a1 > a0) {
                    if (nsp >= kChainSpans) {
                        if (own_end_in > 0 || carry_top) return false;
                        fprintf(stderr, "snacb: chain span table overflow\n"); abort();
                    }
                    // a carry-top class's FIRST span reads the three class rows above row 0 from the carry: pad = 1 + k0
                    const short top_flag = (carry_top && a0 == base) ? static_cast<short>(1 + c.k0) : static_cast<short>(0);
                    out[l][w][nsp++] = ChainSpan{static_cast<short>(c.r0 + 8 * d * (a0 - base)),
                                                 static_cast<short>(a1 - a0), static_cast<short>(c.kc), top_flag};
                }
                base += c.noct;
            }
        }
    }
    memcpy(spans, out, sizeof out);
    return true;
}

// Strips: cost model = tiles every CTA slot walks (a short last tile counts by its rows) + half a strip of tail.
void chain_plan_strips(int C, int t_n, int S, int slots, bool no_carry, int* strip_tiles, int* sps, int* last_strip_tiles,
                       int* last_rows) {
    const int rows = chain_tile_rows(C), own_h = rows - 2 * kChainHalo, own_c = rows - kChainHalo;
    auto plan = [&](int K, int* sps_o, int* lst_o, int* lrows_o) {
        const int SR = own_h + (K - 1) * own_c;
        const int n = (t_n + SR - 1) / SR, rem = t_n - (n - 1) * SR;
        const int lst = rem <= own_h ? 1 : 1 + (rem - own_h + own_c - 1) / own_c;
        int lrows = lst == 1 ? rem : rem - own_h - (lst - 2) * own_c;
        if (lrows == (lst == 1 ? own_h : own_c)) lrows = 0;
        *sps_o = n; *lst_o = lst; *lrows_o = lrows;
        const double last_cost = lrows ? (lrows + (lst == 1 ? 2 : 1) * kChainHalo + 32.0) / rows : 1.0;
        const double per_stream = (n - 1) * static_cast<double>(K) + (lst - 1) + (last_cost < 1.0 ? last_cost : 1.0);
        return per_stream * S / slots + 0.5 * K;
    };
    int bestK = 1, a, b, c;
    double best = plan(1, &a, &b, &c);
    if (!no_carry)
        for (int K = 2; K <= 24; ++K) {
            const double v = plan(K, &a, &b, &c);
            if (v < best - 1e-9) { best = v; bestK = K; }
        }
    plan(bestK, sps, last_strip_tiles, last_rows);
    *strip_tiles = bestK;
}

size_t chain_carry_bytes(int C, int sm_count) {
    return static_cast<size_t>(sm_count) * (chain_warps(C) == 8 ? 2 : 1) * 2 * 39 * C * 2;
}
size_t chain_carry_bytes_max(int sm_count) {
    size_t m = 0;
    for (int C : {64, 128, 256}) { const size_t b = chain_carry_bytes(C, sm_count); if (b > m) m = b; }
    return m;
}

// tm: [0] y load map, box (64, 128, 1); [1] out store map, box (64, 88, 1); [2] out store map, box (64, 128, 1);
//     [3..6] noise 1x1, res d=1, d=3, d=9 weight maps, box (64, C); all 128B-swizzled.
// fold: the alpha-folded fp16 formulation (the d=1, 3, 9 weight maps then carry 1 / alpha2 in their K columns)
cudaError_t launch_chain(int half_fp16, int fold, const ChainArgs& a, const CUtensorMap* tm, int sm_count, cudaStream_t st) {
    if (half_fp16) {
        if (fold && a.jitter != 0u) {                       // race-detector instantiations (SNACB_CHAIN_JITTER)
            if (a.C == 64) return launch_chain_t<64, kNB64, kNW64, __half, true, true>(a, tm, sm_count, st);
            if (a.C == 128) return launch_chain_t<128, kNB128, kNW128, __half, true, true>(a, tm, sm_count, st);
            if (a.C == 256) return launch_chain_t<256, kNB256, kNW256, __half, true, true>(a, tm, sm_count, st);
            return cudaErrorInvalidValue;
        }
        if (fold) {
            if (a.C == 64) return launch_chain_t<64, kNB64, kNW64, __half, true>(a, tm, sm_count, st);
            if (a.C == 128) return launch_chain_t<128, kNB128, kNW128, __half, true>(a, tm, sm_count, st);
            if (a.C == 256) return launch_chain_t<256, kNB256, kNW256, __half, true>(a, tm, sm_count, st);
            return cudaErrorInvalidValue;
        }
        if (a.C == 64) return launch_chain_t<64, kNB64, kNW64, __half, false>(a, tm, sm_count, st);
        if (a.C == 128) return launch_chain_t<128, kNB128, kNW128, __half, false>(a, tm, sm_count, st);
        if (a.C == 256) return launch_chain_t<256, kNB256, kNW256, __half, false>(a, tm, sm_count, st);
        return cudaErrorInvalidValue;
    }
    if (fold) return cudaErrorInvalidValue;
    if (a.C == 64) return launch_chain_t<64, kNB64, kNW64, __nv_bfloat16, false>(a, tm, sm_count, st);
    if (a.C == 128) return launch_chain_t<128, kNB128, kNW128, __nv_bfloat16, false>(a, tm, sm_count, st);
    return cudaErrorInvalidValue;
}

}  // namespace snacb
