// k_convt_res: ConvTranspose1d(k = 2s, stride s) for the layers whose whole weight matrix fits shared memory
// (block 3: Cin 128 -> Cout 64, s = 2, 64 KB), hand-written for sm_100a.
//
//   y[m*s + p][o] = b[o] + sum_tap sum_c  x[m + sh(p) - tap][c] * Wp[p*Cout + o][tap*Cin + c],   sh(p) = [p >= s/2]
//
// The generic k_gemm_tc streams the weight tile and one row-shifted copy of the activation tile per (phase, tap)
// through L2 for every 128-row tile and is L2-bandwidth-bound there (SM <-> L2 ~ 43 B/clk/SM against 190 B/clk/SM
// asked).  Here:
//   * the packed weights [s*Cout][2*Cin] are loaded ONCE per persistent CTA and stay resident;
//   * one TMA load brings the 136 input rows m0-1 .. m0+134 of a tile; every (phase, tap) reads them through a
//     ROW-SHIFTED UMMA descriptor (start address + delta*128 B; the swizzle is a function of the absolute smem
//     address, see ptx.cuh) -- no second copy;
//   * 32 tcgen05.mma (M=128, N=Cout, K=16) per tile into one of two TMEM accumulator stages [128 lanes x s*Cout];
//   * epilogue: 8 warps drain TMEM (+ bias) into a swizzled [256 out rows][128 B] staging tile, one TMA store per
//     tile writes it (direct 16-byte stores at a 256-byte stride were LSU-bound).
// Stream edges are TMA out-of-bounds zero fill (x is already Snake'd, snake(0) = 0, so this is the conv's padding).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace snacb {
using namespace ptx;

namespace {

template <typename HT> struct HalfFmtT;
template <> struct HalfFmtT<__half> { static constexpr uint32_t kFmt = 0; };
template <> struct HalfFmtT<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; };

template <int CIN, int COUT, int S>
struct ConvtCfg {
    static constexpr int kCH = CIN / 64;                 // K chunks per tap
    static constexpr int kN = S * COUT;                  // accumulator columns
    static constexpr int kBoxRows = 136;                 // 128 + halo, multiple of 8 (swizzle atom)
    static constexpr int kAChunk = kBoxRows * 128;
    static constexpr int kAStage = kCH * kAChunk;
    static constexpr int kStages = 3;
    static constexpr int kOutBytes = 128 * S * COUT * 2;  // output staging tile [128*S rows][COUT] 16-bit
    static constexpr int kWChunk = kN * 128;             // [kN rows][64 k]
    static constexpr int kWBytes = 2 * kCH * kWChunk;    // taps x chunks
    static constexpr int kOffA = 0;
    static constexpr int kOffW = kStages * kAStage;
    static constexpr int kOffOut = kOffW + kWBytes;
    static constexpr int kOffBias = kOffOut + kOutBytes;
    static constexpr int kOffBar = kOffBias + COUT * 4;
    static constexpr int kSmem = kOffBar + 128 + 1024;
    static constexpr int kTmemCols = 2 * kN;
    static constexpr int kEpiWarps = 8;
    static constexpr int kThreads = 64 + kEpiWarps * 32;
    static_assert(kN <= 256 && (kTmemCols & (kTmemCols - 1)) == 0, "accumulator width");
    static_assert(kAChunk % 1024 == 0 && kOffOut % 1024 == 0 && kSmem <= 232448, "shared memory");
    static_assert(COUT == 64, "one output row = one 128-byte swizzle row");
    static_assert(S == 2, "tap shifts are written out for stride 2");
};

}  // namespace

template <int CIN, int COUT, int S, typename HT>
__global__ void __launch_bounds__((ConvtCfg<CIN, COUT, S>::kThreads), 1)
k_convt_res(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ CUtensorMap tmO, const GemmArgs a, const int num_tiles) {
    using Cfg = ConvtCfg<CIN, COUT, S>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem + Cfg::kOffA;
    uint8_t* sW = smem + Cfg::kOffW;
    uint8_t* sO = smem + Cfg::kOffOut;
    float* sBias = reinterpret_cast<float*>(smem + Cfg::kOffBias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
    uint64_t* full = bars;                    // [kStages]
    uint64_t* empty = bars + Cfg::kStages;    // [kStages]
    uint64_t* tfull = bars + 2 * Cfg::kStages;   // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t_lo = a.t_n > 0 ? a.t_lo : 0, t_n = a.t_n > 0 ? a.t_n : a.Tin;
    const int tiles_t = (t_n + 127) / 128;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmW);
        prefetch_tmap(&tmO);
        for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], Cfg::kEpiWarps); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
    for (int c = threadIdx.x; c < COUT; c += Cfg::kThreads) sBias[c] = a.bias[c];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_expect_tx(wbar, Cfg::kWBytes);
            for (int j = 0; j < 2 * Cfg::kCH; ++j)
                tma_load_2d_hint(sW + j * Cfg::kWChunk, &tmW, j * 64, 0, wbar, kL2EvictLast);
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int s = tile / tiles_t, m0 = t_lo + sm_off(a.map, s, a.rpf) + (tile % tiles_t) * 128;
                const int sl = sm_slot(a.map, s);
                mbar_wait(&empty[stage], phase ^ 1u);
                mbar_expect_tx(&full[stage], Cfg::kAStage);
                for (int kc = 0; kc < Cfg::kCH; ++kc)
                    tma_load_3d(sA + stage * Cfg::kAStage + kc * Cfg::kAChunk, &tmA, kc * 64, m0 - 1, sl, &full[stage]);
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(128, COUT, HalfFmtT<HT>::kFmt);
        mbar_wait(wbar, 0);
        int stage = 0; uint32_t phase = 0;
        int as = 0; uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[as], aphase ^ 1u);
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_base = smem_u32(sA + stage * Cfg::kAStage);
                const uint32_t w_base = smem_u32(sW);
#pragma unroll
                for (int p = 0; p < S; ++p) {
                    const int sh = (p >= S / 2) ? 1 : 0;
#pragma unroll
                    for (int tap = 0; tap < 2; ++tap) {
                        const uint32_t delta = 1 + sh - tap;           // tile row 0 is input row m0 - 1
#pragma unroll
                        for (int kc = 0; kc < Cfg::kCH; ++kc)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                mma_f16_ss(tmem_base + as * Cfg::kN + p * COUT,
                                           umma_desc_sw128(a_base + kc * Cfg::kAChunk + delta * 128 + k * 32),
                                           umma_desc_sw128(w_base + (tap * Cfg::kCH + kc) * Cfg::kWChunk + p * COUT * 128 + k * 32),
                                           idesc, (tap > 0 || kc > 0 || k > 0) ? 1u : 0u);
                    }
                }
                mma_commit(&empty[stage]);
                mma_commit(&tfull[as]);
            }
            __syncwarp();
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: 2 warps per TMEM lane quadrant
        const int q = warp & 3, half = (warp - 2) >> 2;      // half = output phase p (64 of the 128 accumulator columns)
        const bool leader = (warp == 2 && lane == 0);
        int as = 0; uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int s = tile / tiles_t, m0 = t_lo + sm_off(a.map, s, a.rpf) + (tile % tiles_t) * 128;
                const int sl = sm_slot(a.map, s);
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::kN + half * COUT;
            uint32_t r0[32], r1[32];
            tmem_ld32(taddr, r0);
            tmem_ld32(taddr + 32, r1);
            // the staging tile is free once the previous tile's TMA store has read it
            if (leader) bulk_wait_group_read<0>();
            asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kEpiWarps * 32) : "memory");
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            const int orow = (q * 32 + lane) * S + half;                   // row of the 256-row output tile
            uint8_t* dst = sO + orow * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t* r = (c < 4) ? r0 : r1;
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = (c & 3) * 8 + 2 * e, ch = c * 8 + 2 * e;
                    o[e] = pack2(__uint_as_float(r[col]) + sBias[ch], __uint_as_float(r[col + 1]) + sBias[ch + 1],
                                 static_cast<const HT*>(nullptr));
                }
                *reinterpret_cast<uint4*>(dst + ((c ^ (orow & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, %0;" ::"n"(Cfg::kEpiWarps * 32) : "memory");
            if (leader) {
                tma_store_3d(&tmO, sO, 0, m0 * S, sl);                      // rows beyond T are clipped by TMA
                bulk_commit_group();
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
        if (leader) bulk_wait_group<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// =================================================================================================
// k_convt_ph: ConvTranspose1d whose weights do not fit shared memory as a whole (block 2: Cin 256 -> Cout 128, s = 4,
// 512 KB) but whose slice for ONE output phase does (128 KB).  The persistent CTAs are split into s groups, one per
// output phase p; a CTA keeps its phase's weights [Cout][2*Cin] resident and walks the 128-row input tiles of its
// group, loading each tile's 136 input rows once (one 64-channel chunk per ring stage, both taps read it through
// row-shifted descriptors).  L2 -> SM traffic per 128x128x512 tile: 68 KB (33 B/clk at the tensor rate) instead of the
// 192 KB the generic kernel streams; the s groups read the same activations at about the same time (L2 hits).
// =================================================================================================
namespace {
template <int CIN, int COUT, int S>
struct ConvtPhCfg {
    static constexpr int kCH = CIN / 64;
    static constexpr int kBoxRows = 136;
    static constexpr int kAChunk = kBoxRows * 128;
    static constexpr int kRing = 4;                      // A chunk stages: one tile deep (5 measured the same)
    static constexpr int kWChunk = COUT * 128;           // [COUT rows][64 k]
    static constexpr int kWBytes = 2 * kCH * kWChunk;    // taps x chunks of this phase
    static constexpr int kOffA = 0;
    static constexpr int kOffW = kRing * kAChunk;
    static constexpr int kEpiWarps = 4;                  // one per TMEM lane quadrant
    static constexpr int kStage = 32 * 128;              // a warp's staging tile: 32 rows x 64 channels, 128B-swizzled
    static constexpr int kOffStage = kOffW + kWBytes;
    static constexpr int kOffBias = kOffStage + kEpiWarps * kStage;
    static constexpr int kOffBar = kOffBias + COUT * 4;
    static constexpr int kSmem = kOffBar + 128 + 1024;
    static constexpr int kTmemCols = 2 * COUT;
    static constexpr int kThreads = 64 + kEpiWarps * 32;
    static_assert(COUT == 128 && (kOffW % 1024) == 0 && (kOffStage % 1024) == 0 && kSmem <= 232448, "configuration");
};
}  // namespace

template <int CIN, int COUT, int S, typename HT>
__global__ void __launch_bounds__((ConvtPhCfg<CIN, COUT, S>::kThreads), 1)
k_convt_ph(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
           const __grid_constant__ CUtensorMap tmO, const GemmArgs a, const int num_tiles) {
    using Cfg = ConvtPhCfg<CIN, COUT, S>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem + Cfg::kOffA;
    uint8_t* sW = smem + Cfg::kOffW;
    uint8_t* sStage = smem + Cfg::kOffStage;
    float* sBias = reinterpret_cast<float*>(smem + Cfg::kOffBias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
    uint64_t* full = bars;                    // [kRing]
    uint64_t* empty = bars + Cfg::kRing;      // [kRing]
    uint64_t* tfull = bars + 2 * Cfg::kRing;  // [2]
    uint64_t* tempty = tfull + 2;             // [2]
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t_lo = a.t_n > 0 ? a.t_lo : 0, t_n = a.t_n > 0 ? a.t_n : a.Tin;
    const int tiles_t = (t_n + 127) / 128;
    const int p = blockIdx.x % S;                                   // output phase of this CTA
    const int gi = blockIdx.x / S, gs = (static_cast<int>(gridDim.x) - p + S - 1) / S;   // index / size of its group
    const int sh = (p >= S / 2) ? 1 : 0;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmW);
        prefetch_tmap(&tmO);
        for (int i = 0; i < Cfg::kRing; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], Cfg::kEpiWarps); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
    for (int c = threadIdx.x; c < COUT; c += Cfg::kThreads) sBias[c] = a.bias[c];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_expect_tx(wbar, Cfg::kWBytes);
            for (int j = 0; j < 2 * Cfg::kCH; ++j)
                tma_load_2d_hint(sW + j * Cfg::kWChunk, &tmW, j * 64, p * COUT, wbar, kL2EvictLast);
            int stage = 0; uint32_t phase = 0;
            for (int tile = gi; tile < num_tiles; tile += gs) {
                const int s = tile / tiles_t, m0 = t_lo + sm_off(a.map, s, a.rpf) + (tile % tiles_t) * 128;
                const int sl = sm_slot(a.map, s);
                for (int kc = 0; kc < Cfg::kCH; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_expect_tx(&full[stage], Cfg::kAChunk);
                    tma_load_3d(sA + stage * Cfg::kAChunk, &tmA, kc * 64, m0 - 1, sl, &full[stage]);
                    if (++stage == Cfg::kRing) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(128, COUT, HalfFmtT<HT>::kFmt);
        mbar_wait(wbar, 0);
        int stage = 0; uint32_t phase = 0;
        int as = 0; uint32_t aphase = 0;
        for (int tile = gi; tile < num_tiles; tile += gs) {
            mbar_wait(&tempty[as], aphase ^ 1u);
            for (int kc = 0; kc < Cfg::kCH; ++kc) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_base = smem_u32(sA + stage * Cfg::kAChunk);
                    const uint32_t w_base = smem_u32(sW);
#pragma unroll
                    for (int tap = 0; tap < 2; ++tap) {
                        const uint32_t delta = 1 + sh - tap;               // stage row 0 is input row m0 - 1
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_f16_ss(tmem_base + as * COUT, umma_desc_sw128(a_base + delta * 128 + k * 32),
                                       umma_desc_sw128(w_base + (tap * Cfg::kCH + kc) * Cfg::kWChunk + k * 32), idesc,
                                       (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                    }
                    mma_commit(&empty[stage]);
                    if (kc == Cfg::kCH - 1) mma_commit(&tfull[as]);
                }
                __syncwarp();
                if (++stage == Cfg::kRing) { stage = 0; phase ^= 1u; }
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: one warp per TMEM lane quadrant.
        // A warp drains its 32 rows x 128 channels in two 64-channel halves; each half is packed into the warp's own
        // 4 KB staging tile (128B-swizzled: conflict-free 16-byte stores, lane = row) and leaves with ONE TMA store
        // (box 64 channels x 32 input rows of phase p; the output seen as [slot][Tin][s * Cout]).  Before: every lane
        // stored its row's 128 bytes itself -- 32 cache lines per store instruction, the L1-wavefront-bound epilogue that
        // held the tensor pipe at 54 %.  A 32-row group cut by the end of the row range keeps the per-row stores.
        const int q = warp & 3;
        uint8_t* stg = sStage + (warp - 2) * Cfg::kStage;
        HT* out = static_cast<HT*>(a.out);
        int as = 0; uint32_t aphase = 0;
        for (int tile = gi; tile < num_tiles; tile += gs) {
            const int s = tile / tiles_t, so = sm_off(a.map, s, a.rpf), mw = t_lo + so + (tile % tiles_t) * 128 + q * 32;
            const int sl = sm_slot(a.map, s);
            const int m = mw + lane;
            const int m_end = min(a.Tin, t_lo + so + t_n);
            const bool whole = mw + 32 <= m_end;                 // warp-uniform
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * COUT + half * 64;
                uint32_t r0[32], r1[32];
                tmem_ld32(taddr, r0);
                tmem_ld32(taddr + 32, r1);
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
                if (whole) {
                    if (lane == 0) bulk_wait_group_read<0>();    // the previous store has read the staging tile
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t* r = (c < 4) ? r0 : r1;
                        uint32_t o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int col = (c & 3) * 8 + 2 * e, ch = half * 64 + c * 8 + 2 * e;
                            o[e] = pack2(__uint_as_float(r[col]) + sBias[ch], __uint_as_float(r[col + 1]) + sBias[ch + 1],
                                         static_cast<const HT*>(nullptr));
                        }
                        *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&tmO, stg, p * COUT + half * 64, mw, sl);
                        bulk_commit_group();
                    }
                } else if (m < m_end) {
                    HT* dst = out + (static_cast<size_t>(sl) * a.Tin * S + static_cast<size_t>(m) * S + p) * COUT + half * 64;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + sBias[half * 64 + j];
                    store32(dst, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r1[j]) + sBias[half * 64 + 32 + j];
                    store32(dst + 32, v);
                }
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
        if (lane == 0) bulk_wait_group<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

namespace {
template <int CIN, int COUT, int S, typename HT>
cudaError_t launch_t(const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, int sm_count,
                     cudaStream_t st) {
    using Cfg = ConvtCfg<CIN, COUT, S>;
    static PerDeviceOnce once;
    int dev_;
    if (once.needed(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(k_convt_res<CIN, COUT, S, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        once.done(dev_);
    }
    const int t_n = a.t_n > 0 ? a.t_n : a.Tin;
    const int tiles = a.S * ((t_n + 127) / 128);
    if (tiles == 0) return cudaSuccess;
    const int grid = tiles < sm_count ? tiles : sm_count;
    k_convt_res<CIN, COUT, S, HT><<<grid, Cfg::kThreads, Cfg::kSmem, st>>>(tmA, tmW, tmO, a, tiles);
    return cudaGetLastError();
}
}  // namespace

namespace {
template <int CIN, int COUT, int S, typename HT>
cudaError_t launch_ph_t(const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, int sm_count,
                        cudaStream_t st) {
    using Cfg = ConvtPhCfg<CIN, COUT, S>;
    static PerDeviceOnce once;
    int dev_;
    if (once.needed(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(k_convt_ph<CIN, COUT, S, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        once.done(dev_);
    }
    const int t_n = a.t_n > 0 ? a.t_n : a.Tin;
    const int tiles = a.S * ((t_n + 127) / 128);
    if (tiles == 0) return cudaSuccess;
    int per = sm_count / S;                           // CTAs per phase group
    if (per > tiles) per = tiles;
    if (per < 1) per = 1;
    k_convt_ph<CIN, COUT, S, HT><<<per * S, Cfg::kThreads, Cfg::kSmem, st>>>(tmA, tmW, tmO, a, tiles);
    return cudaGetLastError();
}
}  // namespace

bool convt_ph_supported(int Cin, int Cout, int s) { return Cin == 256 && Cout == 128 && s == 4; }
// tmA: activation map box (64, 136, 1); tmW: packed weights [s*Cout][2*Cin], box (64, Cout); tmO: the output seen as
// [slot][Tin][s*Cout] (one row = the s phases of an input row), box (64, 32, 1); all 128B-swizzled
cudaError_t launch_convt_ph(int half_fp16, const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW,
                            const CUtensorMap& tmO, int sm_count, cudaStream_t st) {
    if (!convt_ph_supported(a.K, a.Cout, a.up)) return cudaErrorInvalidValue;
    return half_fp16 ? launch_ph_t<256, 128, 4, __half>(a, tmA, tmW, tmO, sm_count, st)
                     : launch_ph_t<256, 128, 4, __nv_bfloat16>(a, tmA, tmW, tmO, sm_count, st);
}


#ifdef SNACB_EXPERIMENTS
// =================================================================================================
// k_convt_ph2: k_convt_ph on CTA PAIRS (clusters of two, tcgen05 cta_group::2).  k_convt_ph's M = 128, N = 128 MMAs read
// 4 KB of activations and 4 KB of weights per 64-cycle instruction -- the SM's whole 128 B/clk shared-memory port.  Here a
// pair of CTAs computes a 256-row tile (128 input rows each) per MMA, and each CTA holds only HALF of the phase's weights
// (64 output channels: 64 KB instead of 128 KB): per instruction an SM reads 4 KB of A and 2 KB of W.  Both CTAs run their
// own TMA producer and epilogue; only the leader (rank 0) issues MMAs.  Cross-CTA signalling: both CTAs' TMA loads of a stage
// complete on the LEADER's full barrier (cp.async.bulk.tensor .cta_group::2 with the barrier address mapped into the leader),
// tcgen05.commit multicasts the stage-free and accumulator-ready arrivals to both CTAs, the peer's epilogue warps release the
// accumulator on the leader's tempty barrier (mapa + mbarrier.arrive.shared::cluster).
// EXPERIMENT (SNACB_EXPERIMENTS build, SNACB_CONVT_2CTA=1): correct -- same output as the generic GEMM kernel -- but measured
// SLOWER than k_convt_ph: 596 us against 505 us at B = 1024 (945 us with a relay warp in the peer instead of the direct
// peer -> leader TMA completion; 2 or 4 accumulator stages, 6 or 8 ring stages make no difference).  DESIGN.md section 6.2.
// =================================================================================================
namespace {
template <int CIN, int COUT, int S>
struct ConvtPh2Cfg {
    static constexpr int kCH = CIN / 64;
    static constexpr int kBoxRows = 136;
    static constexpr int kAChunk = kBoxRows * 128;
    static constexpr int kRing = 8;
    static constexpr int kWChunk = (COUT / 2) * 128;     // [COUT / 2 rows][64 k]: this CTA's half of the output channels
    static constexpr int kWBytes = 2 * kCH * kWChunk;
    static constexpr int kOffA = 0;
    static constexpr int kOffW = kRing * kAChunk;
    static constexpr int kEpiWarps = 4;
    static constexpr int kStage = 32 * 128;
    static constexpr int kOffStage = kOffW + kWBytes;
    static constexpr int kOffBias = kOffStage + kEpiWarps * kStage;
    static constexpr int kOffBar = kOffBias + COUT * 4;
    static constexpr int kSmem = kOffBar + 256 + 1024;
    static constexpr int kAcc = 4;                       // accumulator stages (the cross-CTA release of a stage is slow)
    static constexpr int kTmemCols = kAcc * COUT;
    static constexpr int kThreads = 64 + kEpiWarps * 32;
    static_assert(COUT == 128 && (kOffW % 1024) == 0 && (kOffStage % 1024) == 0 && kSmem <= 232448, "configuration");
};
}  // namespace

template <int CIN, int COUT, int S, typename HT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((ConvtPh2Cfg<CIN, COUT, S>::kThreads), 1)
k_convt_ph2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ CUtensorMap tmO, const GemmArgs a, const int num_tiles) {
    using Cfg = ConvtPh2Cfg<CIN, COUT, S>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem + Cfg::kOffA;
    uint8_t* sW = smem + Cfg::kOffW;
    uint8_t* sStage = smem + Cfg::kOffStage;
    float* sBias = reinterpret_cast<float*>(smem + Cfg::kOffBias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
    uint64_t* full = bars;                      // [kRing] leader: BOTH CTAs' A stages have landed (the peer's TMA signals it too)
    uint64_t* empty = bars + Cfg::kRing;        // [kRing] the pair's MMAs have read the stage (multicast commit)
    uint64_t* tfull = bars + 2 * Cfg::kRing;    // [kAcc] accumulator ready (multicast commit)
    uint64_t* tempty = tfull + Cfg::kAcc;       // [kAcc] leader: both CTAs' epilogues have drained the accumulator
    uint64_t* wbar = tempty + Cfg::kAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int t_lo = a.t_n > 0 ? a.t_lo : 0, t_n = a.t_n > 0 ? a.t_n : a.Tin;
    const int tiles_t = (t_n + 127) / 128;
    const int cid = blockIdx.x >> 1, nclusters = static_cast<int>(gridDim.x) >> 1;
    const int p = cid % S;                                          // output phase of this pair
    const int gi = cid / S, gs = (nclusters - p + S - 1) / S;       // index / size of its phase group (in pairs)
    const int sh = (p >= S / 2) ? 1 : 0;
    const int num_pairs = (num_tiles + 1) / 2;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmW);
        prefetch_tmap(&tmO);
        for (int i = 0; i < Cfg::kRing; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < Cfg::kAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * Cfg::kEpiWarps); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc2(tmem_slot, Cfg::kTmemCols); tmem_relinquish2(); }
    for (int c = threadIdx.x; c < COUT; c += Cfg::kThreads) sBias[c] = a.bias[c];
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {                     // this CTA's half of the phase's weights
        mbar_expect_tx(wbar, Cfg::kWBytes);
        for (int j = 0; j < 2 * Cfg::kCH; ++j)
            tma_load_2d_hint(sW + j * Cfg::kWChunk, &tmW, j * 64, p * COUT + static_cast<int>(rank) * (COUT / 2), wbar, kL2EvictLast);
    }
    mbar_wait(wbar, 0);
    cluster_sync_all();                         // both CTAs' barriers exist and both halves of W have landed
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_of = [&](int pair) { const int t = 2 * pair + static_cast<int>(rank); return t < num_tiles ? t : num_tiles - 1; };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            // both CTAs' loads of a stage complete on the LEADER's full barrier (it expects the bytes of the pair)
            int stage = 0; uint32_t phase = 0;
            for (int pair = gi; pair < num_pairs; pair += gs) {
                const int tile = tile_of(pair);
                const int s = tile / tiles_t, m0 = t_lo + sm_off(a.map, s, a.rpf) + (tile % tiles_t) * 128;
                const int sl = sm_slot(a.map, s);
                for (int kc = 0; kc < Cfg::kCH; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::kAChunk);
                    tma_load_3d_pair(sA + stage * Cfg::kAChunk, &tmA, kc * 64, m0 - 1, sl, &full[stage], 0);
                    if (++stage == Cfg::kRing) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        int stage = 0; uint32_t phase = 0;
        if (rank == 0) {
            // -------------------------------------------------------------- MMA issuer (leader)
            constexpr uint32_t idesc = umma_idesc_f16(256, COUT, HalfFmtT<HT>::kFmt);
            int as = 0; uint32_t aphase = 0;
            for (int pair = gi; pair < num_pairs; pair += gs) {
                mbar_wait(&tempty[as], aphase ^ 1u);
                for (int kc = 0; kc < Cfg::kCH; ++kc) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_base = smem_u32(sA + stage * Cfg::kAChunk);
                        const uint32_t w_base = smem_u32(sW);
#pragma unroll
                        for (int tap = 0; tap < 2; ++tap) {
                            const uint32_t delta = 1 + sh - tap;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                mma2_f16_ss(tmem_base + as * COUT, umma_desc_sw128(a_base + delta * 128 + k * 32),
                                            umma_desc_sw128(w_base + (tap * Cfg::kCH + kc) * Cfg::kWChunk + k * 32), idesc,
                                            (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                        }
                        mma2_commit(&empty[stage]);
                        if (kc == Cfg::kCH - 1) mma2_commit(&tfull[as]);
                    }
                    __syncwarp();
                    if (++stage == Cfg::kRing) { stage = 0; phase ^= 1u; }
                }
                if (++as == Cfg::kAcc) { as = 0; aphase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (both CTAs): as in k_convt_ph
        const int q = warp & 3;
        uint8_t* stg = sStage + (warp - 2) * Cfg::kStage;
        HT* out = static_cast<HT*>(a.out);
        int as = 0; uint32_t aphase = 0;
        for (int pair = gi; pair < num_pairs; pair += gs) {
            const bool live = 2 * pair + static_cast<int>(rank) < num_tiles;
            const int tile = tile_of(pair);
            const int s = tile / tiles_t, so = sm_off(a.map, s, a.rpf), mw = t_lo + so + (tile % tiles_t) * 128 + q * 32;
            const int sl = sm_slot(a.map, s);
            const int m = mw + lane;
            const int m_end = live ? min(a.Tin, t_lo + so + t_n) : 0;
            const bool whole = mw + 32 <= m_end;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * COUT + half * 64;
                uint32_t r0[32], r1[32];
                tmem_ld32(taddr, r0);
                tmem_ld32(taddr + 32, r1);
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (rank == 0) mbar_arrive(&tempty[as]);
                        else mbar_arrive_remote(&tempty[as], 0);
                    }
                }
                if (whole) {
                    if (lane == 0) bulk_wait_group_read<0>();
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t* r = (c < 4) ? r0 : r1;
                        uint32_t o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int col = (c & 3) * 8 + 2 * e, ch = half * 64 + c * 8 + 2 * e;
                            o[e] = pack2(__uint_as_float(r[col]) + sBias[ch], __uint_as_float(r[col + 1]) + sBias[ch + 1],
                                         static_cast<const HT*>(nullptr));
                        }
                        *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&tmO, stg, p * COUT + half * 64, mw, sl);
                        bulk_commit_group();
                    }
                } else if (m < m_end) {
                    HT* dst = out + (static_cast<size_t>(sl) * a.Tin * S + static_cast<size_t>(m) * S + p) * COUT + half * 64;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + sBias[half * 64 + j];
                    store32(dst, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r1[j]) + sBias[half * 64 + 32 + j];
                    store32(dst + 32, v);
                }
            }
            if (++as == Cfg::kAcc) { as = 0; aphase ^= 1u; }
        }
        if (lane == 0) bulk_wait_group<0>();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                         // nobody leaves while the peer may still signal or be read
    if (warp == 1) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

namespace {
template <int CIN, int COUT, int S, typename HT>
cudaError_t launch_ph2_t(const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, int sm_count,
                         cudaStream_t st) {
    using Cfg = ConvtPh2Cfg<CIN, COUT, S>;
    static PerDeviceOnce once;
    int dev_;
    if (once.needed(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(k_convt_ph2<CIN, COUT, S, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        once.done(dev_);
    }
    const int t_n = a.t_n > 0 ? a.t_n : a.Tin;
    const int tiles = a.S * ((t_n + 127) / 128);
    if (tiles == 0) return cudaSuccess;
    const int pairs = (tiles + 1) / 2;
    int per = (sm_count / S) / 2;                     // CTA pairs per phase group
    {
        // a persistent kernel must not launch more clusters than can be co-resident (floorswept GPCs hold fewer CTA pairs
        // than sm_count / 2): a second wave would run alone
        static int max_clusters[64];
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && max_clusters[dev] == 0) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(static_cast<unsigned>(2 * (sm_count / 2)));
            cfg.blockDim = dim3(Cfg::kThreads);
            cfg.dynamicSmemBytes = Cfg::kSmem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, k_convt_ph2<CIN, COUT, S, HT>, &cfg) != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = sm_count / 2; }
            max_clusters[dev] = n;
            if (getenv("SNACB_CONVT_2CTA_VERBOSE")) fprintf(stderr, "k_convt_ph2: %d co-resident CTA pairs on %d SMs\n", n, sm_count);
        }
        if (dev >= 0 && dev < 64 && per > max_clusters[dev] / S) per = max_clusters[dev] / S;
    }
    if (per > pairs) per = pairs;
    if (per < 1) per = 1;
    k_convt_ph2<CIN, COUT, S, HT><<<per * S * 2, Cfg::kThreads, Cfg::kSmem, st>>>(tmA, tmW, tmO, a, tiles);
    return cudaGetLastError();
}
}  // namespace

// tmW: packed weights [s*Cout][2*Cin], box (64, Cout / 2); else as launch_convt_ph
cudaError_t launch_convt_ph2(int half_fp16, const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW,
                             const CUtensorMap& tmO, int sm_count, cudaStream_t st) {
    if (!convt_ph_supported(a.K, a.Cout, a.up)) return cudaErrorInvalidValue;
    return half_fp16 ? launch_ph2_t<256, 128, 4, __half>(a, tmA, tmW, tmO, sm_count, st)
                     : launch_ph2_t<256, 128, 4, __nv_bfloat16>(a, tmA, tmW, tmO, sm_count, st);
}

#endif  // SNACB_EXPERIMENTS

bool convt_res_supported(int Cin, int Cout, int s) { return Cin == 128 && Cout == 64 && s == 2; }
int convt_res_box_rows() { return 136; }

// tmA: activation map box (64, 136, 1); tmW: packed ConvTranspose weights [s*Cout][2*Cin], box (64, s*Cout);
// tmO: output map box (64, 128*s, 1); all 128B-swizzled
cudaError_t launch_convt_res(int half_fp16, const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW,
                             const CUtensorMap& tmO, int sm_count, cudaStream_t st) {
    if (!convt_res_supported(a.K, a.Cout, a.up)) return cudaErrorInvalidValue;
    return half_fp16 ? launch_t<128, 64, 2, __half>(a, tmA, tmW, tmO, sm_count, st)
                     : launch_t<128, 64, 2, __nv_bfloat16>(a, tmA, tmW, tmO, sm_count, st);
}

}  // namespace snacb
