// k_chain2: the fused DecoderBlock chain of kernels_chain.cu (NoiseBlock + three ResidualUnits + next Snake, fp32
// residual stream in TMEM, one 16-bit tile copy in shared memory rewritten in place) with the lock-step of its layer
// sequence broken: the CTA's 16 warps form TWO GROUPS that own the upper / lower half of the tile's rows and run
// half a layer apart, so that one group's Snake / depthwise prologue (MUFU + FMA pipes) overlaps the other group's
// MMA wait, TMEM drain, TMA store and refill.  fp16 operands only; the arithmetic of every
// element is the one of k_chain in the same order (outputs are bit-identical, tests/test_gpu_parity.py).
//
//   group 0: rows [0, R/2)   walks every dilation class upwards and ENDS exactly at the group boundary
//   group 1: rows [R/2, R)   starts every class exactly at the boundary
//
// What crosses the boundary, per ResidualUnit l (dilation d):
//   * group 1 needs x_l of the 3d rows above the boundary, which group 0 overwrites in place first: group 0 copies
//     its last 27 rows to a stash (double-buffered by layer parity) before it starts its prologue;
//   * group 0 needs x_l of the 3d rows below the boundary, which group 1's previous epilogue produces while group 0
//     is already inside its prologue: these are the LAST three taps of the classes that end at the boundary, read
//     late, behind an mbarrier (bnd) that group 1's epilogue arrives on after the first 32 rows of its first block;
//   * group 1 must not overwrite those rows before group 0 has read them: it starts its prologue only after group 0
//     finished its own (tok).  The same token is what keeps the two groups in anti-phase.
// Weight buffers are refilled by group 1's leader once BOTH groups' MMAs of the layer have retired (tcgen05.commit
// of either leader on a count-2 mbarrier).  Each group loads / stores its own half of the tile with its own
// barriers, so group 0 runs into the next tile while group 1 finishes the current one.
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace snacb {
using namespace ptx;

namespace {

__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<const __half2*>(&v); }
__device__ __forceinline__ uint32_t as_u32(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }

constexpr int kHalo = kChainHalo;            // 40 >= 39, multiple of 8
constexpr int kStashRows = 27;               // 3 * 9: rows above the group boundary that group 1's taps reach
constexpr int kGroupThreads = 256;
enum { SPAN_TAIL_LATE = 1, SPAN_HEAD_STASH = 2 };   // ChainSpan::pad

template <int C, int NB>
struct Chain2Cfg {
    static constexpr int kCH = C / 64;                      // 64-channel K chunks
    static constexpr int kRows = NB * 128;                  // tile rows incl. halo
    static constexpr int kHalfRows = kRows / 2;             // group boundary
    static constexpr int kROut = kRows - 2 * kHalo;         // rows stored per tile
    static constexpr int kPlane = NB * 16384;               // one chunk plane of the tile [NB][128 rows][128 B]
    static constexpr int kXBytes = kCH * kPlane;
    static constexpr bool kWRes = (C == 64);                // all four 1x1 weights resident
    static constexpr int kWLayer = C * C * 2;               // one layer's weights [kCH][C rows][128 B]
    static constexpr int kWBytes = kWRes ? 4 * kWLayer : 2 * kWLayer;
    static constexpr int kPrmWords = 8;                     // per layer and channel pair
    static constexpr int kPrmBytes = 3 * (C / 2) * kPrmWords * 4;
    static constexpr int kEpiBytes = 8 * C * 4;
    static constexpr int kNzBytes = kRows * 4;
    static constexpr int kStashOne = kCH * kStashRows * 128;
    static constexpr int kSpanBytes = 3 * kChainWarps * kChainSpans * 8;
    static constexpr int kOffX = 1024;                      // 1 KB of slack above the tile: class starts reach 8 rows up
    static constexpr int kOffW = kOffX + kXBytes;
    static constexpr int kOffPrm = kOffW + kWBytes;
    static constexpr int kOffEpi = kOffPrm + kPrmBytes;
    static constexpr int kOffNz = kOffEpi + kEpiBytes;
    static constexpr int kOffSpan = kOffNz + kNzBytes;
    static constexpr int kOffStash = kOffSpan + kSpanBytes;
    static constexpr int kOffBar = kOffStash + 2 * kStashOne;
    static constexpr int kSmem = kOffBar + 256 + 1024;
    static constexpr int kTmemCols = NB * C;
    static_assert(NB == 2 || NB == 4 || NB == 8, "blocks per tile");
    static_assert(kTmemCols == 512 || kTmemCols == 256 || kTmemCols == 128, "TMEM columns");
    static_assert(kSmem <= 232448, "shared memory budget");
    static_assert(kWBytes >= 16384, "tail pre-reads may run up to 100 rows past the tile");
    static_assert(C / 32 >= 2, "both warps of a lane quarter own a piece of the boundary block");
};

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }

// mbarrier wait for the prologue's inner loop: no printf on the timeout path (keeps the ABI call out of the loop)
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > SNACB_WAIT_LIMIT_CYCLES) __trap();
}

// One span of the in-place prologue (see kernels_chain.cu::span_half for the fp16 formulation): rows r0 + k*D,
// k < 8*nq, of one 64-channel chunk, lane = channel pair.  r0 is arbitrary here (the swizzle phases are derived
// from it), and `late` spans fetch their three tail rows from the other group's half behind `bnd_bar`.
template <int D, int ROWS>
__device__ __forceinline__ void span2(uint8_t* plane, int r_oct, const int nq, const uint32_t h0, const uint32_t h1,
                                      const uint32_t h2, uint32_t t0, uint32_t t1, uint32_t t2, const uint32_t lane,
                                      const uint32_t* prm, const bool late, uint64_t* bnd_bar, const uint32_t bnd_par) {
    const uint4 q0 = *reinterpret_cast<const uint4*>(prm), q1 = *reinterpret_cast<const uint4*>(prm + 4);
    const __half2 bd = as_h2(q0.x);
    const __half2 w[7] = {as_h2(q0.y), as_h2(q0.z), as_h2(q0.w), as_h2(q1.x), as_h2(q1.y), as_h2(q1.z), as_h2(q1.w)};
    auto snake = [&](__half2 xh) -> __half2 {      // xh + sin^2(xh)
        const float2 t = __half22float2(xh);
        const __half2 sh = __floats2half2_rn(__sinf(t.x), __sinf(t.y));
        return __hfma2(sh, sh, xh);
    };
    // swizzled byte offset of this lane's channel pair inside a 128-byte row, for step k: row = r_oct + k*D (mod 8)
    uint32_t swz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        swz[j] = ((((lane >> 2) ^ static_cast<uint32_t>(r_oct + j)) & 7u) << 4) + ((lane & 3u) << 2);
    // car[0..5]: snake1 of the six rows in front of the octet's fourth tap (rows -3 .. +2 relative to its first row)
    __half2 car[6];
    car[0] = snake(as_h2(h0)); car[1] = snake(as_h2(h1)); car[2] = snake(as_h2(h2));
    uint8_t* ob = plane + r_oct * 128;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uint32_t raw = 0u;                         // class starts may lie above the tile
        if (r_oct + j * D >= 0) raw = *reinterpret_cast<const uint32_t*>(ob + j * D * 128 + swz[(j * D) & 7]);
        car[3 + j] = snake(as_h2(raw));
    }
#pragma unroll 1
    for (int qo = 0; qo < nq; ++qo) {
        uint32_t raw[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            raw[k] = *reinterpret_cast<const uint32_t*>(ob + (k + 3) * D * 128 + swz[((k + 3) * D) & 7]);
        if (qo == nq - 1) {
            if (late) {
                mbar_wait_lean(bnd_bar, bnd_par);
                t0 = *reinterpret_cast<const uint32_t*>(ob + 8 * D * 128 + swz[(8 * D) & 7]);
                t1 = *reinterpret_cast<const uint32_t*>(ob + 9 * D * 128 + swz[(9 * D) & 7]);
                t2 = *reinterpret_cast<const uint32_t*>(ob + 10 * D * 128 + swz[(10 * D) & 7]);
            }
            raw[5] = t0; raw[6] = t1; raw[7] = t2;
        }
        // The octet runs in four batches instead of step by step: a warp issues in order, and 16 back-to-back MUFUs
        // pipeline through the XU where eight separate dependent chains stall on each one's latency.
        __half2 v[14];                             // snake1 of rows -3 .. +10
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = car[j];
        {
            float sx[8], sy[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float2 t = __half22float2(as_h2(raw[k]));
                sx[k] = __sinf(t.x); sy[k] = __sinf(t.y);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const __half2 sh = __floats2half2_rn(sx[k], sy[k]);
                v[6 + k] = __hfma2(sh, sh, as_h2(raw[k]));
            }
        }
        __half2 acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = bd;
#pragma unroll
        for (int j = 0; j < 7; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = __hfma2(w[j], v[k + j], acc[k]);
        {
            float sx[8], sy[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float2 t = __half22float2(acc[k]);
                sx[k] = __sinf(t.x); sy[k] = __sinf(t.y);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const __half2 sh = __floats2half2_rn(sx[k], sy[k]);
                acc[k] = __hfma2(sh, sh, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = r_oct + k * D;
            if (static_cast<unsigned>(r) < static_cast<unsigned>(ROWS))
                *reinterpret_cast<uint32_t*>(ob + k * D * 128 + swz[(k * D) & 7]) = as_u32(acc[k]);
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) car[j] = v[8 + j];
        r_oct += 8 * D;
        ob += 8 * D * 128;
    }
}

enum { EPI2_NOISE = 0, EPI2_MID = 1, EPI2_FINAL = 2 };

}  // namespace

template <int C, int NB>
__global__ void __launch_bounds__(2 * kGroupThreads)
k_chain2(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmOe,
         const __grid_constant__ CUtensorMap tmOm, const __grid_constant__ CUtensorMap tmWn,
         const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
         const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ ChainArgs a, const int num_tiles) {
    using Cfg = Chain2Cfg<C, NB>;
    using HT = __half;
    constexpr int CH = Cfg::kCH;
    constexpr int NBG = NB / 2;                      // blocks per group
    constexpr int kThreads = 2 * kGroupThreads;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem + Cfg::kOffX;
    uint8_t* sW = smem + Cfg::kOffW;
    uint32_t* sPrm = reinterpret_cast<uint32_t*>(smem + Cfg::kOffPrm);
    float* sEpi = reinterpret_cast<float*>(smem + Cfg::kOffEpi);
    float* sNz = reinterpret_cast<float*>(smem + Cfg::kOffNz);
    ChainSpan* sSpan = reinterpret_cast<ChainSpan*>(smem + Cfg::kOffSpan);   // [3][kChainWarps][kChainSpans]
    uint8_t* sStash = smem + Cfg::kOffStash;                                 // [2][CH][27 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
    uint64_t* ld_bar = bars;          // [2] a group's half of the tile landed (TMA)
    uint64_t* w_bar = bars + 2;       // [2] weight buffers landed
    uint64_t* wfree_bar = bars + 4;   // [2] both groups' MMAs reading a weight buffer have retired (count 2)
    uint64_t* bnd_bar = bars + 6;     // group 1's epilogue wrote the first 32 rows below the boundary (count 2)
    uint64_t* tok_bar = bars + 7;     // group 0 finished a prologue (count 8: one arrival per warp)
    uint64_t* mma_bar = bars + 8;     // [NB] the layer's MMAs of one 128-row block complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + NB);
    volatile int* s_next = reinterpret_cast<volatile int*>(tmem_slot + 2);   // [4] ring: tile n+1 of this CTA at [(n+1) & 3]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const uint32_t lane = tid & 31;
    const int g = warp >> 3;                          // group
    const int gtid = tid & (kGroupThreads - 1);
    const bool leader = gtid == 0;
    const int b_first = g * NBG;                      // first block of the group
    const int tiles_t = ((a.t_n > 0 ? a.t_n : a.T) + Cfg::kROut - 1) / Cfg::kROut;
    const CUtensorMap* wmaps[4] = {&tmWn, &tmW0, &tmW1, &tmW2};

    // ------------------------------------------------------------------ one-time setup
    if (tid == 0) {
        prefetch_tmap(&tmY); prefetch_tmap(&tmOe); prefetch_tmap(&tmOm);
        prefetch_tmap(&tmWn); prefetch_tmap(&tmW0); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
        mbar_init(&ld_bar[0], 1); mbar_init(&ld_bar[1], 1);
        mbar_init(&w_bar[0], 1); mbar_init(&w_bar[1], 1);
        mbar_init(&wfree_bar[0], 2); mbar_init(&wfree_bar[1], 2);
        mbar_init(bnd_bar, 2);
        mbar_init(tok_bar, 8);
        for (int b = 0; b < NB; ++b) mbar_init(&mma_bar[b], 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
    for (int i = tid; i < 3 * (C / 2); i += kThreads) {            // per-layer prologue parameters of one channel pair
        const int l = i / (C / 2), ch = 2 * (i % (C / 2));
        const ChainLayer& L = a.res[l];
        uint32_t* d = sPrm + i * Cfg::kPrmWords;
        // half2 words: alpha2 * dw bias, then dw taps 0..6 times alpha2 / alpha1 (kernels_chain.cu::span_half)
        const float a1x = L.alpha1[ch], a1y = L.alpha1[ch + 1], a2x = L.alpha2[ch], a2y = L.alpha2[ch + 1];
        d[0] = as_u32(__floats2half2_rn(L.dw_b[ch] * a2x, L.dw_b[ch + 1] * a2y));
#pragma unroll
        for (int j = 0; j < 7; ++j)
            d[1 + j] = as_u32(__floats2half2_rn(L.dw_w[j * C + ch] * (a2x / a1x), L.dw_w[j * C + ch + 1] * (a2y / a1y)));
    }
    for (int i = tid; i < 3 * kChainWarps * kChainSpans; i += kThreads)
        sSpan[i] = a.spans[i / (kChainWarps * kChainSpans)][(i / kChainSpans) % kChainWarps][i % kChainSpans];
    // epilogue vectors, as in k_chain: [0] scale of the NoiseBlock output, [1],[2] scale and scaled bias after unit d=1,
    // [3],[4] after d=3, [5] bias after d=9, [6],[7] alpha / 1/alpha of the next Snake
    for (int c = tid; c < C; c += kThreads) {
        const float s0 = a.res[0].alpha1[c], s1 = a.res[1].alpha1[c], s2 = a.res[2].alpha1[c];
        sEpi[c] = s0;
        sEpi[C + c] = s1; sEpi[2 * C + c] = a.bias_cum[c] * s1;
        sEpi[3 * C + c] = s2; sEpi[4 * C + c] = a.bias_cum[C + c] * s2;
        sEpi[5 * C + c] = a.bias_cum[2 * C + c];
        sEpi[6 * C + c] = a.alpha_next[c]; sEpi[7 * C + c] = a.inv_next[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto load_w = [&](int l, int buf) {       // one thread
        uint8_t* dst = sW + (Cfg::kWRes ? l : buf) * Cfg::kWLayer;
        uint64_t* bar = &w_bar[Cfg::kWRes ? 0 : buf];
        if (!Cfg::kWRes) mbar_expect_tx(bar, Cfg::kWLayer);
#pragma unroll
        for (int kc = 0; kc < CH; ++kc) tma_load_2d_hint(dst + kc * (C * 128), wmaps[l], kc * 64, 0, bar, kL2EvictLast);
    };
    auto tile_coords = [&](int tile, int& s, int& t_start) {
        s = tile / tiles_t;
        t_start = (a.t_n > 0 ? a.t_lo : 0) + (tile % tiles_t) * Cfg::kROut - kHalo;
    };
    auto load_block = [&](int s, int t_start, int b) {      // group leader; ld_bar[g]'s expect_tx covers the group's half
#pragma unroll
        for (int kc = 0; kc < CH; ++kc)
            tma_load_3d_hint(sX + kc * Cfg::kPlane + b * 16384, &tmY, kc * 64, t_start + b * 128, s, &ld_bar[g], kL2EvictFirst);
    };
    int tile = blockIdx.x;
    if (tile < num_tiles) {
        if (tid == 0) {
            if (Cfg::kWRes) {
                mbar_expect_tx(&w_bar[0], 4 * Cfg::kWLayer);
                for (int l = 0; l < 4; ++l) load_w(l, 0);
            } else {
                load_w(0, 0);
                load_w(1, 1);
            }
        }
        if (leader) {
            int s, t_start;
            tile_coords(tile, s, t_start);
            mbar_expect_tx(&ld_bar[g], Cfg::kXBytes / 2);
            for (int b = b_first; b < b_first + NBG; ++b) load_block(s, t_start, b);
        }
    }

    uint32_t mma_par = 0;
    constexpr uint32_t idescW = umma_idesc_f16(128, C, 0);
    const uint32_t sx_addr = smem_u32(sX);

    // issue one layer's 1x1 conv for the group's blocks (group leader): TMEM[blk] (+)= A[blk] * W^T, one commit per block
    auto issue_layer = [&](int l, int n) {
        const int buf = l & 1;
        if (Cfg::kWRes) { if (n == 0 && l == 0) mbar_wait(&w_bar[0], 0); }
        else mbar_wait(&w_bar[buf], (2 * n + (l >> 1)) & 1);
        tc_fence_after();
        const uint32_t w_addr = smem_u32(sW + (Cfg::kWRes ? l : buf) * Cfg::kWLayer);
#pragma unroll
        for (int bb = 0; bb < NBG; ++bb) {
            const int b = b_first + bb;
#pragma unroll
            for (int kc = 0; kc < CH; ++kc)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_f16_ss(tmem_base + b * C, umma_desc_sw128(sx_addr + kc * Cfg::kPlane + b * 16384 + k * 32),
                               umma_desc_sw128(w_addr + kc * (C * 128) + k * 32), idescW,
                               (l > 0 || kc > 0 || k > 0) ? 1u : 0u);
            mma_commit(&mma_bar[b]);
        }
        if (!Cfg::kWRes) mma_commit(&wfree_bar[buf]);
    };
    // group 1's leader, after its own epilogue of layer l: once both groups' MMAs of the layer have retired the weight
    // buffer is free -> prefetch the layer two ahead
    auto refill_w = [&](int l, int n, bool has_next) {
        if (Cfg::kWRes) return;
        mbar_wait(&wfree_bar[l & 1], ((4 * n + l) >> 1) & 1);
        if (l + 2 < 4 || has_next) load_w((l + 2) & 3, l & 1);
    };

    // epilogue of one layer over the group's blocks, block by block as their MMAs commit (see k_chain)
    auto epilogue = [&](auto mode_tag, const float* scale, const float* bias, int t_start) {
        constexpr int MODE = decltype(mode_tag)::value;
        const int q = warp & 3, gh = (warp >> 2) & 1;
        constexpr int kCG = C / 32;
        constexpr int kPieces = NBG * kCG;
        const HT* tag = nullptr;
#pragma unroll 1
        for (int it = gh; it < kPieces; it += 2) {
            const int blk = b_first + it / kCG, cg = it % kCG;
            const bool bnd_piece = (MODE != EPI2_FINAL) && g == 1 && q == 0 && it < kCG && it + 2 >= kCG;
            // the last epilogue only feeds the TMA stores: 32-row groups entirely inside the halo are skipped
            if (MODE == EPI2_FINAL && (blk * 128 + q * 32 + 32 <= kHalo || blk * 128 + q * 32 >= Cfg::kRows - kHalo)) continue;
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
            if (MODE == EPI2_NOISE) {
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
                if (MODE == EPI2_NOISE) {
                    const uint32_t* yw = reinterpret_cast<const uint32_t*>(yv);
                    const float2 y0 = __half22float2(as_h2(yw[j / 2])), y1 = __half22float2(as_h2(yw[j / 2 + 1]));
                    v0 = fmaf(nz, v0, y0.x); v1 = fmaf(nz, v1, y0.y); v2 = fmaf(nz, v2, y1.x); v3 = fmaf(nz, v3, y1.y);
                    raw[j] = __float_as_uint(v0); raw[j + 1] = __float_as_uint(v1);
                    raw[j + 2] = __float_as_uint(v2); raw[j + 3] = __float_as_uint(v3);
                    const float4 sc = *reinterpret_cast<const float4*>(scale + cg * 32 + j);
                    v0 *= sc.x; v1 *= sc.y; v2 *= sc.z; v3 *= sc.w;
                } else if (MODE == EPI2_MID) {
                    const float4 sc = *reinterpret_cast<const float4*>(scale + cg * 32 + j);
                    const float4 b = *reinterpret_cast<const float4*>(bias + cg * 32 + j);
                    v0 = fmaf(v0, sc.x, b.x); v1 = fmaf(v1, sc.y, b.y); v2 = fmaf(v2, sc.z, b.z); v3 = fmaf(v3, sc.w, b.w);
                } else {
                    const float4 b = *reinterpret_cast<const float4*>(bias + cg * 32 + j);
                    v0 += b.x; v1 += b.y; v2 += b.z; v3 += b.w;
                    const float4 al = *reinterpret_cast<const float4*>(sEpi + 6 * C + cg * 32 + j);
                    const float4 ia = *reinterpret_cast<const float4*>(sEpi + 7 * C + cg * 32 + j);
                    v0 = snake_f<true>(v0, al.x, ia.x); v1 = snake_f<true>(v1, al.y, ia.y);
                    v2 = snake_f<true>(v2, al.z, ia.z); v3 = snake_f<true>(v3, al.w, ia.w);
                }
                o[j / 2] = pack2(v0, v1, tag);
                o[j / 2 + 1] = pack2(v2, v3, tag);
            }
            if (MODE == EPI2_NOISE) tmem_st32(taddr, raw);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int chunk = ((cg & 1) * 4 + c) ^ (i & 7);
                *reinterpret_cast<uint4*>(row + chunk * 16) =
                    (valid || MODE == EPI2_FINAL) ? make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3])
                                                  : make_uint4(0u, 0u, 0u, 0u);
            }
            if (bnd_piece) {                  // this warp's last piece of the rows right below the boundary
                __syncwarp();
                if (lane == 0) mbar_arrive(bnd_bar);
            }
        }
        if (MODE == EPI2_NOISE) tmem_st_wait();
    };

    // debug phase timing (SNACB_CHAIN_PROF=1): clock64 sums of CTA 0's two group leaders, slots g*32 + phase
    long long t_last = clock64();
    auto tick = [&](int slot) {
        if (a.prof != nullptr && leader && blockIdx.x == 0) {
            const long long t = clock64();
            a.prof[g * 32 + slot] += static_cast<unsigned long long>(t - t_last);
            t_last = t;
        }
    };
    int next_tile = 0;
    bool has_next = false;
    for (int n = 0; tile < num_tiles; ++n) {
        int s, t_start;
        tile_coords(tile, s, t_start);
        // group 0's leader claims the tile after this one (persistent CTAs, dynamic order) for both groups
        if (g == 0 && leader) s_next[(n + 1) & 3] = static_cast<int>(gridDim.x) + atomicAdd(a.tile_counter, 1);

        // ---------------------------------------------------------------- noise values of the group's rows
        {
            unsigned long long key = 0;
            if (a.noise == nullptr)
                key = splitmix64(a.seed * 0x100000001B3ull + static_cast<unsigned long long>(100 + a.noise_stage));
            for (int i = g * Cfg::kHalfRows + gtid; i < (g + 1) * Cfg::kHalfRows; i += kGroupThreads) {
                const int t = t_start + i;
                float v = 0.f;
                if (static_cast<unsigned>(t) < static_cast<unsigned>(a.T))
                    v = a.noise ? a.noise[static_cast<size_t>(s) * a.T + t]
                                : counter_normal(key, noise_counter(a.stream_keys ? a.stream_keys[s] : a.stream_offset + s, t));
                sNz[i] = v;
            }
        }
        mbar_wait(&ld_bar[g], n & 1);
        group_sync(g);
        tick(0);
        if (g == 0) { next_tile = s_next[(n + 1) & 3]; has_next = next_tile < num_tiles; }

        // ---------------------------------------------------------------- NoiseBlock: TMEM = Wn y, then x1 = y + n TMEM
        if (leader) issue_layer(0, n);
        epilogue(std::integral_constant<int, EPI2_NOISE>{}, sEpi, nullptr, t_start);
        if (g == 1 && leader) refill_w(0, n, true);
        mma_par ^= 1u;
        tc_fence_before();
        group_sync(g);
        tick(1);

        // ---------------------------------------------------------------- three ResidualUnits
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            const int d = (l == 0) ? 1 : (l == 1 ? 3 : 9);
            const uint32_t lpar = static_cast<uint32_t>(3 * n + l) & 1u;       // parity of tok / bnd / stash for this unit
            uint8_t* stash = sStash + lpar * Cfg::kStashOne;
            if (g == 0) {
                // x_l of the last 27 rows above the boundary, for group 1 (this group overwrites them below)
                for (int i = gtid; i < CH * kStashRows * 8; i += kGroupThreads) {
                    const int kc = i / (kStashRows * 8), rem = i % (kStashRows * 8);
                    *reinterpret_cast<uint4*>(stash + kc * (kStashRows * 128) + rem * 16) = *reinterpret_cast<const uint4*>(
                        sX + kc * Cfg::kPlane + (Cfg::kHalfRows - kStashRows) * 128 + rem * 16);
                }
            }
            // ---- spans of this warp: pre-read the 3 rows before and after each span (owned by other warps of the group)
            uint32_t hd[kChainSpans][3], tl[kChainSpans][3];
            int r_first[kChainSpans], n_oct[kChainSpans], kcs[kChainSpans], flg[kChainSpans];
            const uint32_t c16 = lane >> 2, sub = (lane & 3u) << 2;
#pragma unroll
            for (int sp = 0; sp < kChainSpans; ++sp) {
                const ChainSpan spn = sSpan[(l * kChainWarps + warp) * kChainSpans + sp];
                r_first[sp] = spn.r_first; n_oct[sp] = spn.n_oct; kcs[sp] = spn.kc; flg[sp] = spn.pad;
#pragma unroll
                for (int j = 0; j < 3; ++j) { hd[sp][j] = 0u; tl[sp][j] = 0u; }
                if (spn.n_oct > 0) {                       // warp-uniform
                    const uint8_t* lane_base = sX + spn.kc * Cfg::kPlane + sub;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int rh = spn.r_first - (3 - j) * d;
                        const int rt = spn.r_first + (8 * spn.n_oct + j) * d;
                        if (rh >= 0 && !(spn.pad & SPAN_HEAD_STASH))
                            hd[sp][j] = *reinterpret_cast<const uint32_t*>(lane_base + rh * 128 + (((c16 ^ rh) & 7) << 4));
                        if (!(spn.pad & SPAN_TAIL_LATE))
                            tl[sp][j] = *reinterpret_cast<const uint32_t*>(lane_base + rt * 128 + (((c16 ^ rt) & 7) << 4));
                    }
                }
            }
            group_sync(g);
            tick(2 + 5 * l);
            if (g == 1) {
                // group 0 has finished its prologue of this unit: its late reads of our boundary rows are done, and its
                // stash is complete
                mbar_wait(tok_bar, lpar);
                if (l == 0) { next_tile = s_next[(n + 1) & 3]; has_next = next_tile < num_tiles; }
#pragma unroll
                for (int sp = 0; sp < kChainSpans; ++sp) {
                    if (n_oct[sp] > 0 && (flg[sp] & SPAN_HEAD_STASH)) {
                        const uint8_t* sb = stash + kcs[sp] * (kStashRows * 128) + sub;
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const int rh = r_first[sp] - (3 - j) * d;
                            hd[sp][j] = *reinterpret_cast<const uint32_t*>(
                                sb + (rh - (Cfg::kHalfRows - kStashRows)) * 128 + (((c16 ^ rh) & 7) << 4));
                        }
                    }
                }
            }
            tick(3 + 5 * l);
#pragma unroll 1
            for (int sp = 0; sp < kChainSpans; ++sp) {
                static_assert(kChainSpans == 4, "span select");
                const int nq = sp == 0 ? n_oct[0] : (sp == 1 ? n_oct[1] : (sp == 2 ? n_oct[2] : n_oct[3]));
                if (nq == 0) continue;
                const int kc = sp == 0 ? kcs[0] : (sp == 1 ? kcs[1] : (sp == 2 ? kcs[2] : kcs[3]));
                const int r0 = sp == 0 ? r_first[0] : (sp == 1 ? r_first[1] : (sp == 2 ? r_first[2] : r_first[3]));
                const int fl = sp == 0 ? flg[0] : (sp == 1 ? flg[1] : (sp == 2 ? flg[2] : flg[3]));
                uint32_t hh[3], tt[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    hh[j] = sp == 0 ? hd[0][j] : (sp == 1 ? hd[1][j] : (sp == 2 ? hd[2][j] : hd[3][j]));
                    tt[j] = sp == 0 ? tl[0][j] : (sp == 1 ? tl[1][j] : (sp == 2 ? tl[2][j] : tl[3][j]));
                }
                const uint32_t* prm = sPrm + ((l * (C / 2)) + kc * 32 + lane) * Cfg::kPrmWords;
                uint8_t* plane = sX + kc * Cfg::kPlane;
                const bool late = (fl & SPAN_TAIL_LATE) != 0;
                if (d == 1) span2<1, Cfg::kRows>(plane, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], lane, prm, late, bnd_bar, lpar);
                else if (d == 3) span2<3, Cfg::kRows>(plane, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], lane, prm, late, bnd_bar, lpar);
                else span2<9, Cfg::kRows>(plane, r0, nq, hh[0], hh[1], hh[2], tt[0], tt[1], tt[2], lane, prm, late, bnd_bar, lpar);
            }
            tick(4 + 5 * l);
            if (g == 0) {                     // hand the prologue token to group 1
                __syncwarp();
                if (lane == 0) mbar_arrive(tok_bar);
            }
            fence_proxy_async_smem();
            group_sync(g);
            tick(5 + 5 * l);
            if (leader) issue_layer(l + 1, n);
            if (l < 2) epilogue(std::integral_constant<int, EPI2_MID>{}, sEpi + (1 + 2 * l) * C, sEpi + (2 + 2 * l) * C, t_start);
            else epilogue(std::integral_constant<int, EPI2_FINAL>{}, nullptr, sEpi + 5 * C, t_start);
            if (g == 1 && leader) refill_w(l + 1, n, has_next);
            mma_par ^= 1u;
            if (l == 2) fence_proxy_async_smem();          // the tile copy is the source of the TMA stores below
            tc_fence_before();
            group_sync(g);
            tick(6 + 5 * l);
        }

        // ---------------------------------------------------------------- stream the group's half out, refill behind it
        if (leader) {
#pragma unroll
            for (int bb = 0; bb < NBG; ++bb) {
                const int b = b_first + bb;
#pragma unroll
                for (int kc = 0; kc < CH; ++kc) {
                    const uint8_t* src = sX + kc * Cfg::kPlane + b * 16384;
                    if (b == 0) tma_store_3d(&tmOe, src + kHalo * 128, kc * 64, t_start + kHalo, s);
                    else if (b == NB - 1) tma_store_3d(&tmOe, src, kc * 64, t_start + b * 128, s);
                    else tma_store_3d(&tmOm, src, kc * 64, t_start + b * 128, s);
                }
                bulk_commit_group();
            }
            if (has_next) {
                int s2, t2;
                tile_coords(next_tile, s2, t2);
                mbar_expect_tx(&ld_bar[g], Cfg::kXBytes / 2);
                // block b may be refilled once the store of block b has read it (groups complete in order)
                if (NBG > 3) { bulk_wait_group_read<3>(); load_block(s2, t2, b_first + NBG - 4); }
                if (NBG > 2) { bulk_wait_group_read<2>(); load_block(s2, t2, b_first + NBG - 3); }
                if (NBG > 1) { bulk_wait_group_read<1>(); load_block(s2, t2, b_first + NBG - 2); }
                bulk_wait_group_read<0>(); load_block(s2, t2, b_first + NBG - 1);
            }
        }
        tick(17);
        tile = next_tile;
    }
    if (leader) bulk_wait_group<0>();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

namespace {

template <int C, int NB>
cudaError_t launch_chain2_t(const ChainArgs& a, const CUtensorMap* tm, int sm_count, cudaStream_t st) {
    using Cfg = Chain2Cfg<C, NB>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_chain2<C, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_chain2<C, NB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const int tiles = a.S * (((a.t_n > 0 ? a.t_n : a.T) + Cfg::kROut - 1) / Cfg::kROut);
    if (tiles == 0) return cudaSuccess;
    const int grid = tiles < sm_count ? tiles : sm_count;
    k_chain2<C, NB><<<grid, 2 * kGroupThreads, Cfg::kSmem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], a, tiles);
    return cudaGetLastError();
}

constexpr int kNB2_64 = 4, kNB2_128 = 4;

}  // namespace

bool chain2_supported(int C) { return C == 64 || C == 128; }
int chain2_tile_rows(int C) { return (C == 64 ? kNB2_64 : kNB2_128) * 128; }

// Spans of the two-group prologue.  For dilation d the rows of a group's half split into d classes r = r0 + k d.
//   group 0 (warps 0-7): class m ends exactly at row R/2 - 1 - m; its start r0 is the last one no larger than the first
//     row whose result the layer must produce (negative starts skip their first few steps); the span that ends a class
//     reads its three tail rows late (SPAN_TAIL_LATE);
//   group 1 (warps 8-15): class m starts exactly at row R/2 + m and takes its three head rows from group 0's stash
//     (SPAN_HEAD_STASH); it runs to the last row the layer must produce, rounded up to whole octets.
// Octets are dealt evenly to the 8 warps of each group.
void chain2_build_spans(int C, ChainSpan (*spans)[kChainWarps][kChainSpans]) {
    const int rows = chain2_tile_rows(C), ch = C / 64, half = rows / 2;
    static const int dil[3] = {1, 3, 9};
    static const int need_lo[3] = {4, 13, 40};            // first row whose result is consumed downstream, per layer
    for (int l = 0; l < 3; ++l) {
        const int d = dil[l];
        const int need_hi = rows - need_lo[l];            // one past the last such row
        for (int w = 0; w < kChainWarps; ++w)
            for (int k = 0; k < kChainSpans; ++k) spans[l][w][k] = ChainSpan{0, 0, 0, 0};
        for (int g = 0; g < 2; ++g) {
            struct Cls { int kc, r0, noct; };
            std::vector<Cls> cls;
            for (int kc = 0; kc < ch; ++kc)
                for (int m = 0; m < d; ++m) {
                    if (g == 0) {
                        const int e = half - 1 - m;                                   // last row of the class in this half
                        int noct = 1;
                        while (e - (8 * noct - 1) * d > need_lo[l]) ++noct;
                        cls.push_back({kc, e - (8 * noct - 1) * d, noct});
                    } else {
                        const int r0 = half + m;
                        const int steps = (need_hi - r0 + d - 1) / d;
                        cls.push_back({kc, r0, (steps + 7) / 8});
                    }
                }
            int total = 0;
            for (auto& c : cls) total += c.noct;
            for (int wg = 0; wg < 8; ++wg) {
                const int w = g * 8 + wg;
                const int lo = static_cast<int>(static_cast<long long>(wg) * total / 8);
                const int hi = static_cast<int>(static_cast<long long>(wg + 1) * total / 8);
                int nsp = 0, base = 0;
                for (auto& c : cls) {
                    const int a0 = lo > base ? lo : base, a1 = hi < base + c.noct ? hi : base + c.noct;
                    if (a1 > a0) {
                        if (nsp >= kChainSpans) { fprintf(stderr, "snacb: chain2 span table overflow\n"); abort(); }
                        short flags = 0;
                        if (g == 0 && a1 == base + c.noct) flags |= SPAN_TAIL_LATE;
                        if (g == 1 && a0 == base) flags |= SPAN_HEAD_STASH;
                        spans[l][w][nsp++] = ChainSpan{static_cast<short>(c.r0 + 8 * d * (a0 - base)),
                                                       static_cast<short>(a1 - a0), static_cast<short>(c.kc), flags};
                    }
                    base += c.noct;
                }
            }
        }
    }
}

// tm: as for launch_chain
cudaError_t launch_chain2(const ChainArgs& a, const CUtensorMap* tm, int sm_count, cudaStream_t st) {
    if (a.C == 64) return launch_chain2_t<64, kNB2_64>(a, tm, sm_count, st);
    if (a.C == 128) return launch_chain2_t<128, kNB2_128>(a, tm, sm_count, st);
    return cudaErrorInvalidValue;
}

}  // namespace snacb
