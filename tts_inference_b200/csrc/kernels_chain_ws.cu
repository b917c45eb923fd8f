// k_chain_ws: the fused DecoderBlock chain of kernels_chain.cu (NoiseBlock + three ResidualUnits + the next layer's
// Snake; fp32 residual stream in TMEM; one 16-bit tile copy in shared memory rewritten in place) with its LOCK-STEP
// BROKEN: warp-specialised roles pipelined over the 128-row blocks of a tile.
//
// Why (profiles/r02_chain_ncu.txt): in the lock-step kernel every warp runs the prologue (depthwise conv + Snake2: bound
// by the FMA pipe and issue slots, XU half idle), then every warp runs the epilogue (TMEM drain + Snake1: bound by the
// XU / MIO queue, FMA pipe half idle), with the tensor pipe, TMA and all of the MMA latency exposed in between.  The two
// phases load COMPLEMENTARY pipes, so they are run at the same time on different blocks:
//
//   warps 0-6   P  prologue of layer l on block b:  A_l = snake2(dw_d(S1_l)) in place, then arrive on a_bar[b]
//   warps 7-14  E  epilogue after layer l on block b (two warps per TMEM lane quadrant): wait mma_bar[b], drain TMEM,
//                  S1_{l+1} = snake1(x_{l+1}) (NoiseBlock: also x1 = y + n Wn y back to TMEM), arrive on s1_bar[b]
//                  (out_bar[b] after the last layer)
//   warp 15     IO one thread: every tcgen05.mma (layer l, block b as soon as its operand is ready: ld_bar[b] for the
//                  NoiseBlock, a_bar[b] else), weight prefetches, TMA store of a finished block and the TMA load of the
//                  next tile's block behind it, tile claims -- as an event loop, so that nothing waits in a fixed order
//
// Dependencies (all through mbarriers; the protocol is emulated on the CPU with random timing in
// tests/test_host.py::test_ws_chain_protocol):
//   P(l, b) needs S1_l of block b, of the 3d rows above it and of the 3d rows below it:
//     * below: s1_bar[b + 1] (the rows are still untouched: P walks the blocks of a layer in order),
//     * above: block b - 1 was already rewritten in place -- every P warp copies the last 27 rows of a block to a carry
//       buffer (double-buffered by block parity) before the block's barrier,
//     * inside the block the 7-tap window slides in registers along a dilation class (rows r, r + d, ...); a class may be
//       cut between two warps at a multiple of 4 steps: both fetch the 3 rows across the cut before the block's barrier.
//   MMA(l, b) needs A_l(b) complete (a_bar[b], 8 arrivals) and the layer's weights; E(l, b) needs MMA(l, b).
//   TMEM: the NoiseBlock epilogue's tcgen05.st is ordered before the next tcgen05.mma on the block by
//   tcgen05.wait::st + fence::before_thread_sync -> s1_bar -> (P) -> a_bar -> fence::after_thread_sync in the IO thread.
//
// Every element goes through the arithmetic of k_chain in the same order: outputs are bit-identical
// (tests/test_gpu_parity.py::test_ws_chain_is_bit_identical).
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "chain_span.cuh"
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace snacb {
using namespace ptx;

namespace {

template <typename HT> struct HalfFmtW;
template <> struct HalfFmtW<__half> { static constexpr uint32_t kFmt = 0; };
template <> struct HalfFmtW<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; };

__device__ __forceinline__ float2 unpack2w(uint32_t v, const __half*) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ float2 unpack2w(uint32_t v, const __nv_bfloat16*) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}
__device__ __forceinline__ __half2 w_h2(uint32_t v) { return *reinterpret_cast<const __half2*>(&v); }
__device__ __forceinline__ uint32_t w_u32(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }

constexpr int kHaloW = kChainHalo;
constexpr int kNP = kChainWsP;               // prologue warps
constexpr int kNE = 8;                       // epilogue warps (two per TMEM lane quadrant)
constexpr int kWsThreads = (kNP + kNE + 1) * 32;
constexpr int kCarryRows = 27;               // 3 * 9: rows above a block that its taps reach

template <int C, int NB, bool HALF, bool FOLD>
struct WsCfg {
    static constexpr int kCH = C / 64;
    static constexpr int kRows = NB * 128;
    static constexpr int kROut = kRows - 2 * kHaloW;
    static constexpr int kPlane = NB * 16384;
    static constexpr int kXBytes = kCH * kPlane;
    static constexpr bool kWRes = (C == 64);                // all four 1x1 weights resident
    static constexpr int kWLayer = C * C * 2;
    static constexpr int kWBytes = kWRes ? 4 * kWLayer : 2 * kWLayer;
    static constexpr int kPrmWords = FOLD ? 8 : (HALF ? 12 : 20);          // see kernels_chain.cu
    static constexpr int kPrmBytes = 3 * (C / 2) * kPrmWords * 4;
    static constexpr int kEpiBytes = 12 * C * 4;
    static constexpr int kCarryOne = kCH * kCarryRows * 128;
    static constexpr int kCarryBytes = ((2 * kCarryOne + 1023) / 1024) * 1024;
    static constexpr int kSpanBytes = 3 * kNP * kChainSpans * 8;
    static constexpr int kBarBytes = 512;
    static constexpr int kOffX = 0;
    static constexpr int kOffW = kOffX + kXBytes;
    static constexpr int kOffCarry = kOffW + kWBytes;
    static constexpr int kOffPrm = kOffCarry + kCarryBytes;
    static constexpr int kOffEpi = kOffPrm + kPrmBytes;
    static constexpr int kOffSpan = kOffEpi + kEpiBytes;
    static constexpr int kOffBar = kOffSpan + kSpanBytes;
    static constexpr int kSmem = kOffBar + kBarBytes + 1024;
    static constexpr int kTmemCols = NB * C;
    static constexpr int kEpiPieces = C / 32 / 2;           // 32-column pieces of a block per E warp
    static_assert(!FOLD || HALF, "the alpha-folded formulation is fp16 only");
    static_assert(C == 64 || C == 128, "C = 256 keeps the lock-step kernel (two blocks per tile leave nothing to pipeline)");
    static_assert(kTmemCols == 512 || kTmemCols == 256, "TMEM columns");
    static_assert(kSmem <= 232448, "shared memory budget");
    static_assert(kWBytes >= 16384, "look-ahead loads of the last block run up to ~100 rows past the tile copy");
    static_assert((4 * NB + 8) * 8 + 64 <= kBarBytes, "barrier block");
};

enum { WS_NOISE = 0, WS_MID = 1, WS_FINAL = 2 };

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

}  // namespace

#ifdef SNACB_EXPERIMENTS
template <int C, int NB, typename HT, bool FOLD>
__global__ void __launch_bounds__(kWsThreads, 1)
k_chain_ws(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmOe,
           const __grid_constant__ CUtensorMap tmOm, const __grid_constant__ CUtensorMap tmWn,
           const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
           const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ ChainArgs a, const int num_tiles) {
    constexpr bool kHalfMath = std::is_same<HT, __half>::value;
    using Cfg = WsCfg<C, NB, kHalfMath, FOLD>;
    constexpr int CH = Cfg::kCH;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem + Cfg::kOffX;
    uint8_t* sW = smem + Cfg::kOffW;
    uint8_t* sCarry = smem + Cfg::kOffCarry;
    uint32_t* sPrm = reinterpret_cast<uint32_t*>(smem + Cfg::kOffPrm);
    float* sEpi = reinterpret_cast<float*>(smem + Cfg::kOffEpi);
    ChainSpan* sSpan = reinterpret_cast<ChainSpan*>(smem + Cfg::kOffSpan);   // [3][kNP][kChainSpans]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
    uint64_t* ld_bar = bars;                  // [NB] a block of the tile landed (TMA), once per tile
    uint64_t* mma_bar = bars + NB;            // [NB] the layer's MMAs of the block complete, 4 per tile
    uint64_t* a_bar = bars + 2 * NB;          // [NB] the block's operand is written (kNP arrivals), 3 per tile
    uint64_t* s1_bar = bars + 3 * NB;         // [NB] the block's S1 is written (kNE arrivals), 3 per tile (noise, mid, mid)
    uint64_t* out_bar = bars + 4 * NB;        // [NB] the block's output rows are written (kNE arrivals), once per tile
    uint64_t* w_bar = bars + 5 * NB;          // [2] weight buffers landed
    uint64_t* tile_bar = bars + 5 * NB + 2;   // the tile after next-to-start is announced in s_tile[], once per tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 * NB + 3);
    volatile int* s_tile = reinterpret_cast<volatile int*>(tmem_slot + 2);   // [2] tile index by local tile parity
    static_assert((5 * NB + 3) * 8 + 16 <= Cfg::kBarBytes, "barrier block");

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tiles_t = ((a.t_n > 0 ? a.t_n : a.T) + Cfg::kROut - 1) / Cfg::kROut;

    // ------------------------------------------------------------------ one-time setup (all warps)
    if (tid == 0) {
        for (int b = 0; b < NB; ++b) {
            mbar_init(&ld_bar[b], 1); mbar_init(&mma_bar[b], 1); mbar_init(&a_bar[b], kNP);
            mbar_init(&s1_bar[b], kNE); mbar_init(&out_bar[b], kNE);
        }
        mbar_init(&w_bar[0], 1); mbar_init(&w_bar[1], 1); mbar_init(tile_bar, 1);
        fence_barrier_init();
    }
    if (warp == kNP + kNE) {
        if (lane == 0) {
            prefetch_tmap(&tmY); prefetch_tmap(&tmOe); prefetch_tmap(&tmOm);
            prefetch_tmap(&tmWn); prefetch_tmap(&tmW0); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
        }
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    for (int i = tid; i < 3 * (C / 2); i += kWsThreads) {          // per-layer prologue parameters of one channel pair
        const int l = i / (C / 2), ch = 2 * (i % (C / 2));
        const ChainLayer& L = a.res[l];
        uint32_t* d = sPrm + i * Cfg::kPrmWords;
        if (FOLD) {
            const float a1x = L.alpha1[ch], a1y = L.alpha1[ch + 1], a2x = L.alpha2[ch], a2y = L.alpha2[ch + 1];
            d[0] = w_u32(__floats2half2_rn(L.dw_b[ch] * a2x, L.dw_b[ch + 1] * a2y));
#pragma unroll
            for (int j = 0; j < 7; ++j)
                d[1 + j] = w_u32(__floats2half2_rn(L.dw_w[j * C + ch] * (a2x / a1x), L.dw_w[j * C + ch + 1] * (a2y / a1y)));
        } else if (kHalfMath) {
            d[0] = w_u32(__floats2half2_rn(L.dw_b[ch], L.dw_b[ch + 1]));
#pragma unroll
            for (int j = 0; j < 7; ++j) d[1 + j] = w_u32(__floats2half2_rn(L.dw_w[j * C + ch], L.dw_w[j * C + ch + 1]));
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
    for (int i = tid; i < 3 * kNP * kChainSpans; i += kWsThreads)
        sSpan[i] = a.spans[i / (kNP * kChainSpans)][(i / kChainSpans) % kNP][i % kChainSpans];
    for (int c = tid; c < C; c += kWsThreads) {                    // epilogue vectors, layout of kernels_chain.cu
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
    if (tid == 0) s_tile[0] = static_cast<int>(blockIdx.x);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int tile, int& s, int& t_start) {
        s = tile / tiles_t;
        t_start = (a.t_n > 0 ? a.t_lo : 0) + (tile % tiles_t) * Cfg::kROut - kHaloW;
    };

    if (warp < kNP) {
        // ============================================================================================ P: prologue warps
        const uint32_t lane_off = (static_cast<uint32_t>(lane >> 2) << 4) | (static_cast<uint32_t>(lane & 3) << 2);
        const int c16 = lane >> 2;
        int tile = static_cast<int>(blockIdx.x);
        const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
        long long t_last = clock64();
        auto tick = [&](int slot) {
            if (prof) { const long long t = clock64(); a.prof[slot] += static_cast<unsigned long long>(t - t_last); t_last = t; }
        };
        for (int n = 0; tile < num_tiles; ++n) {
#pragma unroll 1
            for (int l = 0; l < 3; ++l) {
                const int d = (l == 0) ? 1 : (l == 1 ? 3 : 9);
                const uint32_t s1_par = static_cast<uint32_t>((3 * n + l) & 1);     // s1_bar completes 3x per tile
#pragma unroll 1
                for (int b = 0; b < NB; ++b) {
                    // S1 of this block and of the rows below it (the first rows of block b + 1) must be written
                    if (b == 0) mbar_wait(&s1_bar[0], s1_par);
                    if (b + 1 < NB) mbar_wait(&s1_bar[b + 1], s1_par);
                    tick(0 + 4 * l);                           // waiting for S1
                    // ---- pre-read the 3 rows before / after each of this warp's spans
                    uint32_t hd[kChainSpans][3], tl[kChainSpans][3];
                    int r_first[kChainSpans], n_quad[kChainSpans], kcs[kChainSpans];
#pragma unroll
                    for (int sp = 0; sp < kChainSpans; ++sp) {
                        const ChainSpan spn = sSpan[(l * kNP + warp) * kChainSpans + sp];
                        r_first[sp] = spn.r_first; n_quad[sp] = spn.n_oct; kcs[sp] = spn.kc;
#pragma unroll
                        for (int j = 0; j < 3; ++j) { hd[sp][j] = 0u; tl[sp][j] = 0u; }
                        if (spn.n_oct > 0) {
                            const uint8_t* lane_base = sX + spn.kc * Cfg::kPlane + b * 16384 + ((lane & 3) << 2);
                            const uint8_t* carry = sCarry + ((b + 1) & 1) * Cfg::kCarryOne + spn.kc * (kCarryRows * 128) + ((lane & 3) << 2);
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const int rh = spn.r_first - (3 - j) * d;
                                const int rt = spn.r_first + (4 * spn.n_oct + j) * d;
                                if (rh >= 0) hd[sp][j] = *reinterpret_cast<const uint32_t*>(lane_base + rh * 128 + (((c16 ^ rh) & 7) << 4));
                                else if (b > 0) hd[sp][j] = *reinterpret_cast<const uint32_t*>(carry + (kCarryRows + rh) * 128 + (((c16 ^ rh) & 7) << 4));
                                tl[sp][j] = *reinterpret_cast<const uint32_t*>(lane_base + rt * 128 + (((c16 ^ rt) & 7) << 4));
                            }
                        }
                    }
                    // ---- keep the last rows of this block for block b + 1: it reads them after they were rewritten
                    if (b + 1 < NB) {
                        uint8_t* dst = sCarry + (b & 1) * Cfg::kCarryOne;
                        for (int i = tid; i < CH * kCarryRows * 8; i += kNP * 32) {
                            const int kc = i / (kCarryRows * 8), rem = i % (kCarryRows * 8);
                            *reinterpret_cast<uint4*>(dst + kc * (kCarryRows * 128) + rem * 16) =
                                *reinterpret_cast<const uint4*>(sX + kc * Cfg::kPlane + b * 16384 + (128 - kCarryRows) * 128 + rem * 16);
                        }
                    }
                    named_bar_sync(1, kNP * 32);
                    tick(1 + 4 * l);                           // pre-reads, carry copy, block barrier
#pragma unroll 1
                    for (int sp = 0; sp < kChainSpans; ++sp) {
                        static_assert(kChainSpans == 4, "span select");
                        const int nq = sp == 0 ? n_quad[0] : (sp == 1 ? n_quad[1] : (sp == 2 ? n_quad[2] : n_quad[3]));
                        if (nq == 0) continue;
                        const int kc = sp == 0 ? kcs[0] : (sp == 1 ? kcs[1] : (sp == 2 ? kcs[2] : kcs[3]));
                        const int r0 = sp == 0 ? r_first[0] : (sp == 1 ? r_first[1] : (sp == 2 ? r_first[2] : r_first[3]));
                        uint32_t hh[3], tt[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            hh[j] = sp == 0 ? hd[0][j] : (sp == 1 ? hd[1][j] : (sp == 2 ? hd[2][j] : hd[3][j]));
                            tt[j] = sp == 0 ? tl[0][j] : (sp == 1 ? tl[1][j] : (sp == 2 ? tl[2][j] : tl[3][j]));
                        }
                        const uint32_t* prm = sPrm + ((l * (C / 2)) + kc * 32 + lane) * Cfg::kPrmWords;
                        uint8_t* blk = sX + kc * Cfg::kPlane + b * 16384;
                        if (kHalfMath) {
                            if (d == 1) ws_span_half<1, FOLD>(blk, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], lane_off, prm);
                            else if (d == 3) ws_span_half<3, FOLD>(blk, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], lane_off, prm);
                            else ws_span_half<9, FOLD>(blk, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], lane_off, prm);
                        } else {
                            if (d == 1) ws_span_bf16<1>(blk, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], lane_off, prm);
                            else if (d == 3) ws_span_bf16<3>(blk, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], lane_off, prm);
                            else ws_span_bf16<9>(blk, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], lane_off, prm);
                        }
                    }
                    fence_proxy_async_smem();                  // the tensor core reads the operand through the async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_bar[b]);
                    tick(2 + 4 * l);                           // spans
                }
            }
            // the next tile (announced by the IO thread while this one was in flight)
            mbar_wait(tile_bar, static_cast<uint32_t>(n & 1));
            tile = s_tile[(n + 1) & 1];
            tick(12);
        }
        if (prof) a.prof[15] = static_cast<unsigned long long>(clock64());
    } else if (warp < kNP + kNE) {
        // ============================================================================================ E: epilogue warps
        const int ew = warp - kNP;
        const int q = warp & 3;                     // TMEM lane quadrant this warp may access
        const int hsel = ew >> 2;                   // which half of the 32-column pieces
        const HT* tag = nullptr;
        unsigned long long key = 0;
        if (a.noise == nullptr)
            key = splitmix64(a.seed * 0x100000001B3ull + static_cast<unsigned long long>(100 + a.noise_stage));
        int tile = static_cast<int>(blockIdx.x);
        const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == kNP * 32;
        long long t_last = clock64();
        auto tick = [&](int slot) {
            if (prof) { const long long t = clock64(); a.prof[slot] += static_cast<unsigned long long>(t - t_last); t_last = t; }
        };
        for (int n = 0; tile < num_tiles; ++n) {
            int s, t_start;
            tile_coords(tile, s, t_start);
            auto epilogue = [&](auto mode_tag, const int bnd, const uint32_t mma_par) {
                constexpr int MODE = decltype(mode_tag)::value;
                constexpr bool kFoldHere = FOLD && MODE != WS_FINAL;
                const float* vb = sEpi + (3 * bnd) * C;
                const float* va = vb + C;
                const float* vi = va + C;
#pragma unroll 1
                for (int blk = 0; blk < NB; ++blk) {
                    const int i = blk * 128 + q * 32 + lane;          // tile row of this thread
                    const int t = t_start + i;
                    const bool valid = static_cast<unsigned>(t) < static_cast<unsigned>(a.T);
                    // the last epilogue only feeds the TMA stores: 32-row groups entirely inside the halo are skipped
                    const bool skip = MODE == WS_FINAL && (blk * 128 + q * 32 + 32 <= kHaloW || blk * 128 + q * 32 >= Cfg::kRows - kHaloW);
                    float nz = 0.f;
                    if (MODE == WS_NOISE && valid)
                        nz = a.noise ? a.noise[static_cast<size_t>(s) * a.T + t]
                                     : counter_normal(key, noise_counter(a.stream_keys ? a.stream_keys[s] : a.stream_offset + s, t + a.t0));
                    mbar_wait(&mma_bar[blk], mma_par);
                    tc_fence_after();
                    tick(17 + 2 * bnd);                        // waiting for the MMAs
#pragma unroll 1
                    for (int pc = 0; pc < Cfg::kEpiPieces && !skip; ++pc) {
                        const int cg = hsel * Cfg::kEpiPieces + pc;
                        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + blk * C + cg * 32;
                        uint32_t raw[32];
                        tmem_ld32(taddr, raw);
                        uint8_t* row = sX + (cg >> 1) * Cfg::kPlane + i * 128;
                        uint4 yv[4];
                        if (MODE == WS_NOISE) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) yv[c] = *reinterpret_cast<const uint4*>(row + ((((cg & 1) * 4 + c) ^ (i & 7)) << 4));
                        }
                        tmem_ld_wait();
                        uint32_t o[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float v0 = __uint_as_float(raw[j]), v1 = __uint_as_float(raw[j + 1]);
                            float v2 = __uint_as_float(raw[j + 2]), v3 = __uint_as_float(raw[j + 3]);
                            if (MODE == WS_NOISE) {
                                const uint32_t* yw = reinterpret_cast<const uint32_t*>(yv);
                                const float2 y0 = unpack2w(yw[j / 2], tag), y1 = unpack2w(yw[j / 2 + 1], tag);
                                v0 = fmaf(nz, v0, y0.x); v1 = fmaf(nz, v1, y0.y); v2 = fmaf(nz, v2, y1.x); v3 = fmaf(nz, v3, y1.y);
                                raw[j] = __float_as_uint(v0); raw[j + 1] = __float_as_uint(v1);
                                raw[j + 2] = __float_as_uint(v2); raw[j + 3] = __float_as_uint(v3);
                            }
                            const float4 al = *reinterpret_cast<const float4*>(va + cg * 32 + j);
                            if (kFoldHere) {
                                if (MODE == WS_NOISE) { v0 *= al.x; v1 *= al.y; v2 *= al.z; v3 *= al.w; }
                                else {
                                    const float4 bb = *reinterpret_cast<const float4*>(vb + cg * 32 + j);
                                    v0 = fmaf(v0, al.x, bb.x); v1 = fmaf(v1, al.y, bb.y); v2 = fmaf(v2, al.z, bb.z); v3 = fmaf(v3, al.w, bb.w);
                                }
                                // the same pipe split as k_chain's epilogue (kEpiPolyMask), so that the two kernels stay bit-identical
                                if ((kEpiPolyMask >> (j / 2)) & 1u) {
                                    o[j / 2] = cs_u32(snake_h2_poly(__floats2half2_rn(v0, v1)));
                                } else {
                                    const float s0 = __sinf(v0), s1 = __sinf(v1);
                                    o[j / 2] = pack2(fmaf(s0, s0, v0), fmaf(s1, s1, v1), tag);
                                }
                                if ((kEpiPolyMask >> (j / 2 + 1)) & 1u) {
                                    o[j / 2 + 1] = cs_u32(snake_h2_poly(__floats2half2_rn(v2, v3)));
                                } else {
                                    const float s2 = __sinf(v2), s3 = __sinf(v3);
                                    o[j / 2 + 1] = pack2(fmaf(s2, s2, v2), fmaf(s3, s3, v3), tag);
                                }
                                continue;
                            } else {
                                if (MODE != WS_NOISE) {
                                    const float4 bb = *reinterpret_cast<const float4*>(vb + cg * 32 + j);
                                    v0 += bb.x; v1 += bb.y; v2 += bb.z; v3 += bb.w;
                                }
                                const float4 ia = *reinterpret_cast<const float4*>(vi + cg * 32 + j);
                                v0 = snake_f<true>(v0, al.x, ia.x); v1 = snake_f<true>(v1, al.y, ia.y);
                                v2 = snake_f<true>(v2, al.z, ia.z); v3 = snake_f<true>(v3, al.w, ia.w);
                            }
                            o[j / 2] = pack2(v0, v1, tag);
                            o[j / 2 + 1] = pack2(v2, v3, tag);
                        }
                        if (MODE == WS_NOISE) tmem_st32(taddr, raw);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int chunk = ((cg & 1) * 4 + c) ^ (i & 7);
                            *reinterpret_cast<uint4*>(row + chunk * 16) =
                                (valid || MODE == WS_FINAL) ? make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3])
                                                            : make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                    if (MODE == WS_NOISE) tmem_st_wait();
                    if (MODE == WS_FINAL) fence_proxy_async_smem();        // the rows are the source of a TMA store
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(MODE == WS_FINAL ? &out_bar[blk] : &s1_bar[blk]);
                    tick(18 + 2 * bnd);                        // drain + Snake
                }
            };
            // mma_bar[b] completes four times per tile: parity = layer & 1
            epilogue(std::integral_constant<int, WS_NOISE>{}, 0, 0u);
            epilogue(std::integral_constant<int, WS_MID>{}, 1, 1u);
            epilogue(std::integral_constant<int, WS_MID>{}, 2, 0u);
            epilogue(std::integral_constant<int, WS_FINAL>{}, 3, 1u);
            mbar_wait(tile_bar, static_cast<uint32_t>(n & 1));
            tile = s_tile[(n + 1) & 1];
        }
    } else if (lane == 0) {
        // ============================================================================================ IO thread
        const CUtensorMap* wmaps[4] = {&tmWn, &tmW0, &tmW1, &tmW2};
        constexpr uint32_t idescW = umma_idesc_f16(128, C, HalfFmtW<HT>::kFmt);
        const uint32_t sx_addr = smem_u32(sX);
        auto load_w = [&](int l, int buf) {
            uint8_t* dst = sW + (Cfg::kWRes ? l : buf) * Cfg::kWLayer;
            uint64_t* bar = &w_bar[Cfg::kWRes ? 0 : buf];
            if (!Cfg::kWRes) mbar_expect_tx(bar, Cfg::kWLayer);
#pragma unroll
            for (int kc = 0; kc < CH; ++kc) tma_load_2d_hint(dst + kc * (C * 128), wmaps[l], kc * 64, 0, bar, kL2EvictLast);
        };
        auto load_block = [&](int s, int t_start, int b) {
            mbar_expect_tx(&ld_bar[b], CH * 16384);
#pragma unroll
            for (int kc = 0; kc < CH; ++kc)
                tma_load_3d_hint(sX + kc * Cfg::kPlane + b * 16384, &tmY, kc * 64, t_start + b * 128, s, &ld_bar[b], kL2EvictFirst);
        };
        auto issue_mma = [&](int l, int b) {
            const uint32_t w_addr = smem_u32(sW + (Cfg::kWRes ? l : (l & 1)) * Cfg::kWLayer);
#pragma unroll
            for (int kc = 0; kc < CH; ++kc)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_f16_ss(tmem_base + b * C, umma_desc_sw128(sx_addr + kc * Cfg::kPlane + b * 16384 + k * 32),
                               umma_desc_sw128(w_addr + kc * (C * 128) + k * 32), idescW, (l > 0 || kc > 0 || k > 0) ? 1u : 0u);
            mma_commit(&mma_bar[b]);
        };
        int tile = static_cast<int>(blockIdx.x);
        if (tile < num_tiles) {
            if (Cfg::kWRes) {
                mbar_expect_tx(&w_bar[0], 4 * Cfg::kWLayer);
                for (int l = 0; l < 4; ++l) load_w(l, 0);
            } else {
                load_w(0, 0);
                load_w(1, 1);
            }
            int s, t_start;
            tile_coords(tile, s, t_start);
            for (int b = 0; b < NB; ++b) load_block(s, t_start, b);
        }
        // Event loop.  MMA stream: (mt, ml, mb) = local tile, layer, block of the next MMA to issue; it may run ahead
        // into the next tile's NoiseBlock while the store stream (st, sb) still drains the current tile.
        int mt = 0, ml = 0, mb = 0;                 // MMA stream position
        int st = 0, sb = 0;                         // store stream position (local tile, block)
        int pf_l = -1, pf_t = 0;                    // pending weight prefetch: after layer pf_l of local tile pf_t completed
        int tile_of[4] = {tile, num_tiles, num_tiles, num_tiles};   // global tile index of local tile n at [n & 3] (the store
                                                    // stream may still need tile n when tile n + 2 is claimed)
        bool next_claimed = false;                  // tile mt + 1 has been claimed
        int n_local = tile < num_tiles ? 1 : 0;     // local tiles that exist so far (grows with successful claims)
        bool more = true;                           // the last claim was inside the tile range
        long long spin_t0 = clock64();
        while (n_local > 0 && (mt < n_local || st < n_local)) {
            bool progressed = false;
            // ---- MMA stream
            if (mt < n_local) {
                bool ready;
                if (ml == 0) ready = mbar_test_wait(&ld_bar[mb], static_cast<uint32_t>(mt & 1));
                else ready = mbar_test_wait(&a_bar[mb], static_cast<uint32_t>((3 * mt + ml - 1) & 1));
                if (ready && mb == 0) {
                    // the layer's weights: resident (C = 64, once) or buffer ml & 1, landed (2 * mt + (ml >> 1))-th time
                    if (Cfg::kWRes) { if (mt == 0 && ml == 0) ready = mbar_test_wait(&w_bar[0], 0u); }
                    else ready = mbar_test_wait(&w_bar[ml & 1], static_cast<uint32_t>((2 * mt + (ml >> 1)) & 1));
                }
                if (ready) {
                    tc_fence_after();
                    issue_mma(ml, mb);
                    progressed = true;
                    if (ml == 1 && mb == 0 && !next_claimed) {
                        // every P and E warp is inside local tile mt by now: announce the tile after it
                        const int nt = more ? static_cast<int>(gridDim.x) + atomicAdd(a.tile_counter, 1) : num_tiles;
                        more = nt < num_tiles;
                        tile_of[(mt + 1) & 3] = nt;
                        s_tile[(mt + 1) & 1] = nt;
                        mbar_arrive(tile_bar);              // release: s_tile is visible to whoever acquires the phase
                        if (more) ++n_local;
                        next_claimed = true;
                    }
                    if (++mb == NB) {
                        mb = 0;
                        if (!Cfg::kWRes) { pf_l = ml; pf_t = mt; }      // prefetch once this layer's MMAs have completed
                        if (++ml == 4) { ml = 0; ++mt; next_claimed = false; }
                    }
                }
            }
            // ---- weight prefetch: layer pf_l's buffer is free once its last block's MMAs completed
            if (pf_l >= 0 && mbar_test_wait(&mma_bar[NB - 1], static_cast<uint32_t>(pf_l & 1))) {
                const bool has_next = pf_t + 1 < n_local;
                if (pf_l + 2 < 4 || has_next) load_w((pf_l + 2) & 3, pf_l & 1);
                pf_l = -1;
                progressed = true;
            }
            // ---- store stream: a finished block goes out, the next tile's block comes in behind it
            if (st < n_local && st <= mt && !(st == mt && ml < 3) && mbar_test_wait(&out_bar[sb], static_cast<uint32_t>(st & 1))) {
                int s, t_start;
                tile_coords(tile_of[st & 3], s, t_start);
#pragma unroll
                for (int kc = 0; kc < CH; ++kc) {
                    const uint8_t* src = sX + kc * Cfg::kPlane + sb * 16384;
                    if (sb == 0) tma_store_3d(&tmOe, src + kHaloW * 128, kc * 64, t_start + kHaloW, s);
                    else if (sb == NB - 1) tma_store_3d(&tmOe, src, kc * 64, t_start + sb * 128, s);
                    else tma_store_3d(&tmOm, src, kc * 64, t_start + sb * 128, s);
                }
                bulk_commit_group();
                if (st + 1 < n_local) {
                    int s2, t2;
                    tile_coords(tile_of[(st + 1) & 3], s2, t2);
                    bulk_wait_group_read<0>();              // the store has read the block: it may be refilled
                    load_block(s2, t2, sb);
                }
                if (++sb == NB) { sb = 0; ++st; }
                progressed = true;
            }
            if (progressed) { spin_t0 = clock64(); continue; }
            __nanosleep(32);                                // nothing ready: leave the issue slots to the working warps
            if (a.prof != nullptr && blockIdx.x == 0) a.prof[16] += 1;
            if (clock64() - spin_t0 > SNACB_WAIT_LIMIT_CYCLES) {
                printf("snacb: k_chain_ws IO thread stuck (block %d: mma %d/%d/%d store %d/%d of %d tiles)\n", blockIdx.x, mt, ml, mb, st, sb, n_local);
                __trap();
            }
        }
        bulk_wait_group<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kNP + kNE) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

#endif  // SNACB_EXPERIMENTS

namespace {

#ifdef SNACB_EXPERIMENTS
template <int C, int NB, typename HT, bool FOLD>
cudaError_t launch_ws_t(const ChainArgs& a, const CUtensorMap* tm, int sm_count, cudaStream_t st) {
    using Cfg = WsCfg<C, NB, std::is_same<HT, __half>::value, FOLD>;
    static PerDeviceOnce once;
    int dev_;
    if (once.needed(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_ws<C, NB, HT, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_chain_ws<C, NB, HT, FOLD>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        once.done(dev_);
    }
    const int tiles = a.S * (((a.t_n > 0 ? a.t_n : a.T) + Cfg::kROut - 1) / Cfg::kROut);
    if (tiles == 0) return cudaSuccess;
    const int grid = tiles < sm_count ? tiles : sm_count;
    k_chain_ws<C, NB, HT, FOLD><<<grid, kWsThreads, Cfg::kSmem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], a, tiles);
    return cudaGetLastError();
}
#endif  // SNACB_EXPERIMENTS

constexpr int kWsNB64 = 8, kWsNB128 = 4;

}  // namespace

// bf16 operands with C = 128: the fp32-math prologue parameters do not fit beside the carry buffer -> lock-step kernel
bool chain_ws_supported(int C, int half_fp16) { return C == 64 || (C == 128 && half_fp16); }
int chain_ws_tile_rows(int C) { return (C == 64 ? kWsNB64 : kWsNB128) * 128; }

// Spans of the block-pipelined prologue: inside a 128-row block the rows of dilation d split into d classes r = m + k d;
// a class is cut at multiples of 4 steps ("quads"), and the concatenated quads of all (chunk, class) pairs are dealt out
// evenly to the kChainWsP prologue warps.  ChainSpan: r_first = block-relative first row, n_oct = QUADS, kc = chunk.
void chain_ws_build_spans(int C, ChainSpan (*spans)[kChainWarps][kChainSpans]) {
    const int ch = C / 64;
    static const int dil[3] = {1, 3, 9};
    for (int l = 0; l < 3; ++l) {
        const int d = dil[l];
        struct Cls { int kc, m, quads; };
        std::vector<Cls> cls;
        for (int kc = 0; kc < ch; ++kc)
            for (int m = 0; m < d; ++m) {
                const int steps = (128 - m + d - 1) / d;
                cls.push_back({kc, m, (steps + 3) / 4});
            }
        int total = 0;
        for (auto& c : cls) total += c.quads;
        for (int w = 0; w < kChainWarps; ++w) {
            for (int k = 0; k < kChainSpans; ++k) spans[l][w][k] = ChainSpan{0, 0, 0, 0};
            if (w >= kChainWsP) continue;
            const int lo = static_cast<int>(static_cast<long long>(w) * total / kChainWsP);
            const int hi = static_cast<int>(static_cast<long long>(w + 1) * total / kChainWsP);
            int nsp = 0, base = 0;
            for (auto& c : cls) {
                const int a0 = lo > base ? lo : base, a1 = hi < base + c.quads ? hi : base + c.quads;
                if (a1 > a0) {
                    if (nsp >= kChainSpans) { fprintf(stderr, "snacb: chain_ws span table overflow\n"); abort(); }
                    spans[l][w][nsp++] = ChainSpan{static_cast<short>(c.m + 4 * d * (a0 - base)), static_cast<short>(a1 - a0),
                                                   static_cast<short>(c.kc), 0};
                }
                base += c.quads;
            }
        }
    }
}

#ifdef SNACB_EXPERIMENTS
bool chain_ws_built() { return true; }
cudaError_t launch_chain_ws(int half_fp16, int fold, const ChainArgs& a, const CUtensorMap* tm, int sm_count, cudaStream_t st) {
    if (half_fp16) {
        if (fold) {
            if (a.C == 64) return launch_ws_t<64, kWsNB64, __half, true>(a, tm, sm_count, st);
            if (a.C == 128) return launch_ws_t<128, kWsNB128, __half, true>(a, tm, sm_count, st);
            return cudaErrorInvalidValue;
        }
        if (a.C == 64) return launch_ws_t<64, kWsNB64, __half, false>(a, tm, sm_count, st);
        if (a.C == 128) return launch_ws_t<128, kWsNB128, __half, false>(a, tm, sm_count, st);
        return cudaErrorInvalidValue;
    }
    if (fold) return cudaErrorInvalidValue;
    if (a.C == 64) return launch_ws_t<64, kWsNB64, __nv_bfloat16, false>(a, tm, sm_count, st);
    return cudaErrorInvalidValue;
}
#else
// The kernel is an experiment that lost its A/B: compiled only with SNACB_EXPERIMENTS=1 at build time.
bool chain_ws_built() { return false; }
cudaError_t launch_chain_ws(int, int, const ChainArgs&, const CUtensorMap*, int, cudaStream_t) { return cudaErrorNotSupported; }
#endif

}  // namespace snacb
