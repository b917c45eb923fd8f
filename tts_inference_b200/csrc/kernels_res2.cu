// k_resunit2: persistent, software-pipelined ResidualUnit kernel for sm_100a (16-bit activations).
//
//   out = x + W * snake2( dw_b + sum_j dw_w[j] * snake1(x[t + (j-3) d]) ) + b     (+ next Snake)
//
// One CTA per SM loops over 256-row (C<=256) or 128-row (C=512) time tiles of one stream each:
//   * warp 18 (producer): TMA loads of the x tile + dilation halo, one 64-channel chunk per stage
//     (stream edges = TMA out-of-bounds zero fill, which is exactly the conv's zero padding because
//     snake(0) = 0), and of the 1x1 weights (resident for C<=128, streamed per chunk otherwise);
//   * warps 0-17 (compute): lane = channel pair, each warp owns one segment of 16 (or 8) output rows of one
//     dilation class (rows r, r+d, r+2d, ...) so that the 7-tap window slides in registers: Snake1 is evaluated once per
//     input row, the depthwise conv and Snake2 run in fp32, the result is packed to the 16-bit operand
//     type straight into the 128B-swizzled K-major A tile;
//   * warp 19 (MMA): tcgen05.mma per chunk into TMEM accumulators (two stages when they fit), so the
//     tensor work of tile i overlaps the CUDA-core prologue of tile i+1;
//   * warps 20-23 (epilogue, one per TMEM lane quadrant) drain TMEM (tcgen05.ld), add bias + residual
//     (+ the next layer's Snake) and store the 16-bit result, concurrently with the next tile's prologue.
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace snacb {
using namespace ptx;

namespace {

template <typename HT> struct HalfFmt2;
template <> struct HalfFmt2<__half> { static constexpr uint32_t kFmt = 0; };
template <> struct HalfFmt2<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; };

__device__ __forceinline__ float2 unpack2(uint32_t v, const __half*) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ float2 unpack2(uint32_t v, const __nv_bfloat16*) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

constexpr int kComputeWarps = 18;     // prologue: 18 segments cover a tile for every dilation (see seg table)
constexpr int kEpiWarps = 8;          // epilogue: two warps per TMEM lane quadrant (warps 20..27, warp%4 = quadrant)
constexpr int kLockstepEpiWarps = 16; // C >= 256: the first 16 prologue warps drain TMEM themselves (4 quadrants x 4 column groups)

template <int C, int DIL>
struct Res2Cfg {
    static constexpr int kTileM = (C == 512) ? 128 : 256;
    static constexpr int kAccs = kTileM / 128;                       // 128-row accumulators per tile
    static constexpr int kChunks = C / 64;
    static constexpr bool kWRes = (C <= 128);                        // 1x1 weights resident in smem
    static constexpr int kRows = kTileM + 6 * DIL;                   // x rows per tile incl. halo (even)
    static constexpr int kHalfRows = kRows / 2;                      // TMA box rows (<= 256)
    static constexpr int kXsBytes = ((kRows * 128 + 1023) / 1024) * 1024;
    static constexpr int kNXS = (C <= 128) ? 3 : 2;                  // x chunk stages
    static constexpr int kABytes = kTileM * 128;                     // A operand stage (one 64-ch chunk)
    static constexpr int kNSA = 2;
    static constexpr int kWChunkBytes = C * 128;                     // [C (n)][64 (k)] 16-bit
    static constexpr int kNHalfW = C > 256 ? 2 : 1;                  // streamed weights: one stage = one <=256-row N half
    static constexpr int kWStageBytes = kWRes ? kWChunkBytes : (kWChunkBytes / kNHalfW);
    static constexpr int kNSW = kWRes ? kChunks : (C > 256 ? 4 : 2); // C = 512: 4 x 32 KB ring (3 loads in flight)
    static constexpr int kAccCols = kAccs * C;                       // TMEM columns of one accumulator stage
    static constexpr int kAccStages = (2 * kAccCols <= 512) ? 2 : 1;
    static constexpr int kTmemCols = kAccStages * kAccCols;          // 128 / 256 / 512 (power of two)
    static constexpr int kNHalf = C > 256 ? 2 : 1;
    static constexpr int kMmaN = C > 256 ? 256 : C;
    static constexpr int kSegLen = kTileM / 16;                      // rows per prologue segment: 16 or 8
    static constexpr bool kSplitEpi = (C <= 128);                    // dedicated epilogue warps (else prologue warps do it)
    static constexpr int kThreads = (kComputeWarps + 2 + (kSplitEpi ? kEpiWarps : 0)) * 32;
    static constexpr bool kParamsSmem = (C <= 256);
    static constexpr int kPrmBytes = kParamsSmem ? (C / 2) * 96 : 0; // 24 floats per channel pair
    // C = 512 (block 0): the RESIDUAL is added by the tensor core.  The x chunk that the prologue reads sits in shared
    // memory anyway; loaded 128B-swizzled it is a valid UMMA A operand, and an MMA against a 64 x 64 identity adds it --
    // exactly: 16-bit x times 1.0 into the fp32 accumulator -- onto the chunk's 64 output columns.  The epilogue then has no
    // global loads left (they were the long pole: row-per-lane 32-byte loads, 32 cache lines per instruction).
    static constexpr bool kResMma = (C == 512);
    static constexpr int kIdentBytes = kResMma ? 64 * 128 : 0;
    static constexpr int kEpiBytes = 3 * C * 4;                      // bias, alpha_next, inv_alpha_next
    static constexpr int kMapBytes = 128;                            // segment table (int per compute warp)
    static constexpr int kBarBytes = 256;
    static constexpr int kOffA = 0;
    static constexpr int kOffW = kOffA + kNSA * kABytes;
    static constexpr int kOffXs = kOffW + kNSW * kWStageBytes;
    static constexpr int kOffIdent = kOffXs + kNXS * kXsBytes;
    static constexpr int kOffPrm = kOffIdent + kIdentBytes;
    static constexpr int kOffEpi = kOffPrm + kPrmBytes;
    static constexpr int kOffMap = kOffEpi + kEpiBytes;
    static constexpr int kOffBar = ((kOffMap + kMapBytes + 15) / 16) * 16;
    static constexpr int kSmem = kOffBar + kBarBytes + 1024;
    static_assert(kRows % 2 == 0 && kHalfRows <= 256, "TMA box");
    static_assert(!kResMma || (kRows <= 256 && kAccs == 1 && (kOffIdent % 1024) == 0 && (kXsBytes % 1024) == 0), "residual MMA");
    static_assert(kSmem <= 232448, "shared memory budget");
};

template <int C, int DIL, int EPI, typename HT>
__global__ void __launch_bounds__((Res2Cfg<C, DIL>::kThreads), 1)
k_resunit2(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
           const __grid_constant__ CUtensorMap tmO, const ResUnitArgs a, const int num_tiles) {
    using Cfg = Res2Cfg<C, DIL>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // keep pointer provenance (no integer round trip) so that the compiler emits LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem + Cfg::kOffA;
    uint8_t* sW = smem + Cfg::kOffW;
    uint8_t* sX = smem + Cfg::kOffXs;
    float* sPrm = reinterpret_cast<float*>(smem + Cfg::kOffPrm);
    float* sEpi = reinterpret_cast<float*>(smem + Cfg::kOffEpi);
    int* sSeg = reinterpret_cast<int*>(smem + Cfg::kOffMap);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
    uint64_t* x_full = bars;              // [3]
    uint64_t* x_empty = bars + 3;         // [3]
    uint64_t* a_full = bars + 6;          // [2]
    uint64_t* a_empty = bars + 8;         // [2]
    uint64_t* w_full = bars + 10;         // [4]
    uint64_t* w_empty = bars + 14;        // [4]
    uint64_t* acc_full = bars + 18;       // [2]
    uint64_t* acc_empty = bars + 20;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int t_lo = a.t_n > 0 ? a.t_lo : 0;
    const int tiles_t = ((a.t_n > 0 ? a.t_n : a.T) + Cfg::kTileM - 1) / Cfg::kTileM;

    if (tid == 0) {
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmW);
        if (Cfg::kResMma) prefetch_tmap(&tmO);
        for (int i = 0; i < 3; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], kComputeWarps + (Cfg::kResMma ? 1 : 0)); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], kComputeWarps); mbar_init(&a_empty[i], 1);
            mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
            mbar_init(&w_full[i + 2], 1); mbar_init(&w_empty[i + 2], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], Cfg::kSplitEpi ? kEpiWarps : kLockstepEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == kComputeWarps + 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
    // parameters -> shared memory (compute threads)
    if (tid < kComputeWarps * 32) {
        if (Cfg::kParamsSmem) {
            for (int p = tid; p < C / 2; p += kComputeWarps * 32) {
                float* d = sPrm + p * 24;
                const int ch = 2 * p;
                d[0] = a.alpha1[ch]; d[1] = a.alpha1[ch + 1];
                d[2] = a.inv_alpha1[ch]; d[3] = a.inv_alpha1[ch + 1];
#pragma unroll
                for (int j = 0; j < 7; ++j) { d[4 + 2 * j] = a.dw_w[j * C + ch]; d[5 + 2 * j] = a.dw_w[j * C + ch + 1]; }
                d[18] = a.dw_b[ch]; d[19] = a.dw_b[ch + 1];
                d[20] = a.alpha2[ch]; d[21] = a.alpha2[ch + 1];
                d[22] = a.inv_alpha2[ch]; d[23] = a.inv_alpha2[ch + 1];
            }
        }
        for (int c = tid; c < C; c += kComputeWarps * 32) {
            sEpi[c] = a.pw_b[c];
            sEpi[C + c] = (EPI == EPI_RES_SNAKE) ? a.alpha_next[c] : 0.f;
            sEpi[2 * C + c] = (EPI == EPI_RES_SNAKE) ? a.inv_alpha_next[c] : 0.f;
        }
        // Prologue segments: the rows of a tile split by dilation class (c, c+DIL, c+2 DIL, ...), each class cut
        // into runs of kSegLen rows (the last run of a class is shifted back so that it ends with the class:
        // a few rows are then computed twice with identical results).  <= 18 segments for DIL in {1,3,9}.
        if (tid == 0) {
            int nseg = 0;
            for (int c = 0; c < DIL; ++c) {
                const int n = (Cfg::kTileM - c + DIL - 1) / DIL;
                const int k = (n + Cfg::kSegLen - 1) / Cfg::kSegLen;
                for (int j = 0; j < k; ++j) {
                    int start = j * Cfg::kSegLen;
                    if (start > n - Cfg::kSegLen) start = n - Cfg::kSegLen;
                    if (nseg < kComputeWarps) sSeg[nseg++] = c + start * DIL;
                }
            }
            for (; nseg < kComputeWarps; ++nseg) sSeg[nseg] = -1;
        }
    }
    if (Cfg::kResMma) {
        // identity [64 n][64 k], K-major, 128B-swizzled (the B operand of the residual MMA)
        uint8_t* sI = smem + Cfg::kOffIdent;
        for (int i = tid; i < 64 * 64; i += Cfg::kThreads) {
            const int n = i >> 6, k = i & 63;
            *reinterpret_cast<HT*>(sI + sw128_offset(n, k)) = static_cast<HT>(n == k ? 1.0f : 0.0f);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_tiles = (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (warp == kComputeWarps) {
        // ============================================================ producer (TMA)
        if (lane == 0) {
            if (Cfg::kWRes) {
                mbar_expect_tx(&w_full[0], Cfg::kChunks * Cfg::kWChunkBytes);
                for (int kc = 0; kc < Cfg::kChunks; ++kc)
                    tma_load_2d(sW + kc * Cfg::kWChunkBytes, &tmW, kc * 64, 0, &w_full[0]);
            }
            int g = 0;
            for (int n = 0; n < my_tiles; ++n) {
                const int tile = blockIdx.x + n * gridDim.x;
                const int s_ = tile / tiles_t, s = sm_slot(a.map, s_);             // s: the stream's buffer slot (addressing only)
                const int t0 = t_lo + sm_off(a.map, s_, a.rpf) + (tile % tiles_t) * Cfg::kTileM;
                for (int kc = 0; kc < Cfg::kChunks; ++kc, ++g) {
                    const int sx = g % Cfg::kNXS;
                    if (g >= Cfg::kNXS) mbar_wait(&x_empty[sx], ((g / Cfg::kNXS) - 1) & 1);
                    mbar_expect_tx(&x_full[sx], Cfg::kRows * 128);
                    uint8_t* dst = sX + sx * Cfg::kXsBytes;
                    if (Cfg::kResMma) {                    // one box of kRows rows, 128B-swizzled (tmX: box (64, kRows, 1))
                        tma_load_3d(dst, &tmX, kc * 64, t0 - 3 * DIL, s, &x_full[sx]);
                    } else {
                        tma_load_3d(dst, &tmX, kc * 64, t0 - 3 * DIL, s, &x_full[sx]);
                        tma_load_3d(dst + Cfg::kHalfRows * 128, &tmX, kc * 64, t0 - 3 * DIL + Cfg::kHalfRows, s, &x_full[sx]);
                    }
                    if (!Cfg::kWRes) {
#pragma unroll
                        for (int nh = 0; nh < Cfg::kNHalfW; ++nh) {
                            const int h = g * Cfg::kNHalfW + nh, sw = h % Cfg::kNSW;
                            if (h >= Cfg::kNSW) mbar_wait(&w_empty[sw], ((h / Cfg::kNSW) - 1) & 1);
                            mbar_expect_tx(&w_full[sw], Cfg::kWStageBytes);
                            tma_load_2d(sW + sw * Cfg::kWStageBytes, &tmW, kc * 64, nh * 256, &w_full[sw]);
                        }
                    }
                }
            }
        }
    } else if (warp == kComputeWarps + 1) {
        // ============================================================ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(128, Cfg::kMmaN, HalfFmt2<HT>::kFmt);
        if (Cfg::kWRes) { mbar_wait(&w_full[0], 0); }
        int g = 0;
        for (int n = 0; n < my_tiles; ++n) {
            const int as = n % Cfg::kAccStages;
            long long tp2 = clock64();
            mbar_wait(&acc_empty[as], ((n / Cfg::kAccStages) & 1) ^ 1);
            if (a.prof && blockIdx.x == 0 && lane == 0) a.prof[2] += clock64() - tp2;
            tc_fence_after();
            const uint32_t d_base = tmem_base + as * Cfg::kAccCols;
            for (int kc = 0; kc < Cfg::kChunks; ++kc, ++g) {
                const int sa = g % 2;
                // residual (kResMma): acc[:, kc*64 .. +64) += x[rows 3 DIL .. 3 DIL + 128 of the stage] * I.  It needs the x
                // stage only, not the prologue's A operand: chunks >= 1 issue it as soon as the stage has landed (the
                // accumulator was initialised by chunk 0's first 1x1 MMA), chunk 0 after its 1x1 MMAs.  The stage is
                // released to the producer when these MMAs retire (x_empty counts them beside the 18 prologue warps).
                auto residual_mma = [&]() {
                    const int sx = g % Cfg::kNXS;
                    mbar_wait(&x_full[sx], (g / Cfg::kNXS) & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        constexpr uint32_t idescI = umma_idesc_f16(128, 64, HalfFmt2<HT>::kFmt);
                        const uint32_t x_addr = smem_u32(sX + sx * Cfg::kXsBytes) + 3 * DIL * 128;
                        const uint32_t i_addr = smem_u32(smem + Cfg::kOffIdent);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_f16_ss(d_base + kc * 64, umma_desc_sw128(x_addr + k * 32), umma_desc_sw128(i_addr + k * 32), idescI, 1u);
                        mma_commit(&x_empty[sx]);
                    }
                    __syncwarp();
                };
                if (Cfg::kResMma && kc > 0) residual_mma();
                long long tp0 = clock64();
                mbar_wait(&a_full[sa], (g / 2) & 1);
                if (a.prof && blockIdx.x == 0 && lane == 0) a.prof[0] += clock64() - tp0;
                const uint32_t a_addr = smem_u32(sA + sa * Cfg::kABytes);
#pragma unroll
                for (int nh = 0; nh < Cfg::kNHalf; ++nh) {
                    // streamed weights: one ring stage per N half, released as soon as its MMAs retire
                    const int h = g * Cfg::kNHalfW + nh, sw = h % Cfg::kNSW;
                    long long tp1 = clock64();
                    if (!Cfg::kWRes) mbar_wait(&w_full[sw], (h / Cfg::kNSW) & 1);
                    if (a.prof && blockIdx.x == 0 && lane == 0) a.prof[1] += clock64() - tp1;
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t w_addr = Cfg::kWRes ? smem_u32(sW + kc * Cfg::kWChunkBytes + nh * (256 * 128))
                                                           : smem_u32(sW + sw * Cfg::kWStageBytes);
#pragma unroll
                        for (int ac = 0; ac < Cfg::kAccs; ++ac)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                mma_f16_ss(d_base + ac * C + nh * 256, umma_desc_sw128(a_addr + ac * 16384 + k * 32),
                                           umma_desc_sw128(w_addr + k * 32), idesc, (kc > 0 || k > 0) ? 1u : 0u);
                        if (!Cfg::kWRes) mma_commit(&w_empty[sw]);
                    }
                    __syncwarp();
                }
                if (Cfg::kResMma && kc == 0) residual_mma();
                if (lane == 0) {
                    mma_commit(&a_empty[sa]);
                    if (kc == Cfg::kChunks - 1) mma_commit(&acc_full[as]);
                }
                __syncwarp();
            }
        }
    } else if (Cfg::kSplitEpi && warp >= kComputeWarps + 2) {
        // ============================================================ 8 epilogue warps (concurrent with the prologue)
        const HT* x = static_cast<const HT*>(a.x);
        HT* out = static_cast<HT*>(a.out);
        const int q = warp & 3;
        const int eh = (warp - (kComputeWarps + 2)) >> 2;       // which half of the 32-column pieces this warp takes
        constexpr int kIters = Cfg::kAccs * (C / 32);          // 32-column pieces per tile and quadrant
        for (int n = 0; n < my_tiles; ++n) {
            const int tile = blockIdx.x + n * gridDim.x;
            const int s_ = tile / tiles_t, s = sm_slot(a.map, s_);             // s: the stream's buffer slot (addressing only)
                const int t0 = t_lo + sm_off(a.map, s_, a.rpf) + (tile % tiles_t) * Cfg::kTileM;
            const int as = n % Cfg::kAccStages;
            const size_t srow = static_cast<size_t>(s) * a.T;
            // residual of the first piece is fetched before waiting for the accumulator
            uint4 xr[4], xn[4];
            auto fetch = [&](int it, uint4 (&dst)[4]) {
                const int ac = it / (C / 32), col = (it % (C / 32)) * 32;
                const int t = t0 + ac * 128 + q * 32 + lane;
                if (t < a.T) {
                    const uint4* p = reinterpret_cast<const uint4*>(x + (srow + t) * C + col);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = p[j];
                }
            };
            fetch(eh, xr);
            mbar_wait(&acc_full[as], (n / Cfg::kAccStages) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int it = eh; it < kIters; it += 2) {
                const int ac = it / (C / 32), col = (it % (C / 32)) * 32;
                const int t = t0 + ac * 128 + q * 32 + lane;
                const bool valid = t < a.T;
                uint32_t raw[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::kAccCols + ac * C + col, raw);
                if (it + 2 < kIters) fetch(it + 2, xn);
                tmem_ld_wait();
                if (valid) {
                    uint4 o[4];
                    const uint32_t* xw = reinterpret_cast<const uint32_t*>(xr);
                    uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(sEpi + col + j);
                        const float2 x0 = unpack2(xw[j / 2], static_cast<const HT*>(nullptr));
                        const float2 x1 = unpack2(xw[j / 2 + 1], static_cast<const HT*>(nullptr));
                        float v0 = x0.x + (__uint_as_float(raw[j]) + b.x);
                        float v1 = x0.y + (__uint_as_float(raw[j + 1]) + b.y);
                        float v2 = x1.x + (__uint_as_float(raw[j + 2]) + b.z);
                        float v3 = x1.y + (__uint_as_float(raw[j + 3]) + b.w);
                        if (EPI == EPI_RES_SNAKE) {
                            const float4 al = *reinterpret_cast<const float4*>(sEpi + C + col + j);
                            const float4 ia = *reinterpret_cast<const float4*>(sEpi + 2 * C + col + j);
                            v0 = snake_f<true>(v0, al.x, ia.x);
                            v1 = snake_f<true>(v1, al.y, ia.y);
                            v2 = snake_f<true>(v2, al.z, ia.z);
                            v3 = snake_f<true>(v3, al.w, ia.w);
                        }
                        ow[j / 2] = pack2(v0, v1, static_cast<const HT*>(nullptr));
                        ow[j / 2 + 1] = pack2(v2, v3, static_cast<const HT*>(nullptr));
                    }
                    uint4* dst = reinterpret_cast<uint4*>(out + (srow + t) * C + col);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = o[j];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) xr[j] = xn[j];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
        }
    } else {
        // ============================================================ 18 prologue warps
        // lock-step epilogue (C >= 256): warp = (TMEM quadrant, accumulator / column group)
        auto epilogue = [&](int n) {
            const HT* x = static_cast<const HT*>(a.x);
            HT* out = static_cast<HT*>(a.out);
            const int tile = blockIdx.x + n * gridDim.x;
            const int s_ = tile / tiles_t, s = sm_slot(a.map, s_);             // s: the stream's buffer slot (addressing only)
                const int t0 = t_lo + sm_off(a.map, s_, a.rpf) + (tile % tiles_t) * Cfg::kTileM;
            const int as = n % Cfg::kAccStages;
            const int q = warp & 3, rest = warp >> 2;
            const int ac = (Cfg::kAccs == 2) ? (rest & 1) : 0;
            constexpr int kColGroups = (Cfg::kAccs == 2) ? 2 : 4;
            const int cg = (Cfg::kAccs == 2) ? (rest >> 1) : rest;
            constexpr int kColsPerWarp = C / kColGroups;
            const int t = t0 + ac * 128 + q * 32 + lane;
            const bool valid = t < a.T;
            const size_t grow = static_cast<size_t>(s) * a.T + t;
            // the residual of a 32-column piece is fetched one piece ahead (the first one before the accumulator is
            // ready): its L2 latency is the long pole of this epilogue, not the arithmetic
            U32x8 xa[2], xb[2];
            auto fetch = [&](int cc, U32x8 (&dst)[2]) {
                if (Cfg::kResMma) return;                  // the residual is in the accumulator already
                if (valid) {
                    const HT* p = x + grow * C + cg * kColsPerWarp + cc * 32;
                    dst[0] = ld_global_v8(p);
                    dst[1] = ld_global_v8(p + 16);
                }
            };
            fetch(0, xa);
            mbar_wait(&acc_full[as], (n / Cfg::kAccStages) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Cfg::kAccCols + ac * C;
#pragma unroll 1
            for (int cc = 0; cc < kColsPerWarp / 32; ++cc) {
                const int col = cg * kColsPerWarp + cc * 32;
                uint32_t raw[32];
                tmem_ld32(taddr + col, raw);
                if (cc + 1 < kColsPerWarp / 32) fetch(cc + 1, xb);
                tmem_ld_wait();
                if (valid || Cfg::kResMma) {
                    const uint32_t* xw = reinterpret_cast<const uint32_t*>(xa);
                    U32x8 o[2];
                    uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(sEpi + col + j);
                        float2 x0 = make_float2(0.f, 0.f), x1 = make_float2(0.f, 0.f);
                        if (!Cfg::kResMma) {
                            x0 = unpack2(xw[j / 2], static_cast<const HT*>(nullptr));
                            x1 = unpack2(xw[j / 2 + 1], static_cast<const HT*>(nullptr));
                        }
                        float v0 = x0.x + (__uint_as_float(raw[j]) + b.x);
                        float v1 = x0.y + (__uint_as_float(raw[j + 1]) + b.y);
                        float v2 = x1.x + (__uint_as_float(raw[j + 2]) + b.z);
                        float v3 = x1.y + (__uint_as_float(raw[j + 3]) + b.w);
                        if (EPI == EPI_RES_SNAKE) {
                            const float4 al = *reinterpret_cast<const float4*>(sEpi + C + col + j);
                            const float4 ia = *reinterpret_cast<const float4*>(sEpi + 2 * C + col + j);
                            v0 = snake_f<true>(v0, al.x, ia.x);
                            v1 = snake_f<true>(v1, al.y, ia.y);
                            v2 = snake_f<true>(v2, al.z, ia.z);
                            v3 = snake_f<true>(v3, al.w, ia.w);
                        }
                        ow[j / 2] = pack2(v0, v1, static_cast<const HT*>(nullptr));
                        ow[j / 2 + 1] = pack2(v2, v3, static_cast<const HT*>(nullptr));
                    }
                    if (!Cfg::kResMma) {
                        st_global_v8(out + grow * C + col, o[0]);
                        st_global_v8(out + grow * C + col + 16, o[1]);
                    } else {
                        // kResMma: the 32 rows x 32 columns leave through the warp's 2 KB slot of the A stages (idle until the
                        // next tile's prologue; 64-byte rows, 64B-swizzled: conflict-free 16-byte stores) and ONE TMA store
                        // (box 32 x 32; rows past T are clipped) instead of 32 cache lines per store instruction
                        uint8_t* stg = sA + warp * 2048;
                        if (lane == 0) bulk_wait_group_read<0>();      // the previous piece's store has read the slot
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
                                make_uint4(ow[4 * c], ow[4 * c + 1], ow[4 * c + 2], ow[4 * c + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&tmO, stg, col, t0 + ac * 128 + q * 32, s);
                            bulk_commit_group();
                        }
                    }
                }
                xa[0] = xb[0]; xa[1] = xb[1];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
            if (Cfg::kResMma) {                                // the A stages go back to the prologue (barrier below)
                if (lane == 0) bulk_wait_group_read<0>();
                __syncwarp();
            }
        };

        int g = 0;
        const int seg_base = sSeg[warp];
        // swizzled x stage (kResMma): the 16-byte chunk of this lane's channel pair is XORed with (row & 7); input row j of
        // the segment is stage row seg_base + j DIL, whose (row & 7) repeats with period 8 in j (DIL is odd)
        uint32_t xswz[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            xswz[j] = static_cast<uint32_t>(seg_base) * 128u + ((((lane >> 2) ^ (seg_base + j * DIL)) & 7) << 4) + ((lane & 3) << 2);
        for (int n = 0; n < my_tiles; ++n) {
            for (int kc = 0; kc < Cfg::kChunks; ++kc, ++g) {
                // ---- per-lane parameters of channels (kc*64 + 2*lane, +1)
                float2 al1, ia1, w[7], bd, al2, ia2;
                if (Cfg::kParamsSmem) {
                    const float4* p4 = reinterpret_cast<const float4*>(sPrm + (kc * 32 + lane) * 24);
                    const float4 p0 = p4[0], p1 = p4[1], p2 = p4[2], p3 = p4[3], p5 = p4[4], p6 = p4[5];
                    al1 = make_float2(p0.x, p0.y); ia1 = make_float2(p0.z, p0.w);
                    w[0] = make_float2(p1.x, p1.y); w[1] = make_float2(p1.z, p1.w);
                    w[2] = make_float2(p2.x, p2.y); w[3] = make_float2(p2.z, p2.w);
                    w[4] = make_float2(p3.x, p3.y); w[5] = make_float2(p3.z, p3.w);
                    w[6] = make_float2(p5.x, p5.y); bd = make_float2(p5.z, p5.w);
                    al2 = make_float2(p6.x, p6.y); ia2 = make_float2(p6.z, p6.w);
                } else {
                    const int ch = kc * 64 + 2 * lane;
                    al1 = *reinterpret_cast<const float2*>(a.alpha1 + ch);
                    ia1 = *reinterpret_cast<const float2*>(a.inv_alpha1 + ch);
#pragma unroll
                    for (int j = 0; j < 7; ++j) w[j] = *reinterpret_cast<const float2*>(a.dw_w + j * C + ch);
                    bd = *reinterpret_cast<const float2*>(a.dw_b + ch);
                    al2 = *reinterpret_cast<const float2*>(a.alpha2 + ch);
                    ia2 = *reinterpret_cast<const float2*>(a.inv_alpha2 + ch);
                }
                const int sx = g % Cfg::kNXS, sa = g % 2;
                const bool pr = a.prof && blockIdx.x == 0 && tid == 0;
                long long tq = clock64();
                mbar_wait(&x_full[sx], (g / Cfg::kNXS) & 1);
                if (pr) { const long long t = clock64(); a.prof[3] += t - tq; tq = t; }
                if (g >= 2) mbar_wait(&a_empty[sa], ((g / 2) - 1) & 1);
                if (pr) { const long long t = clock64(); a.prof[4] += t - tq; tq = t; }
                const uint8_t* xs = sX + sx * Cfg::kXsBytes;
                uint8_t* dstA = sA + sa * Cfg::kABytes;

                // ---- this warp's segment: kSegLen output rows seg_base + i*DIL; the 7-tap window of Snake1'd rows
                //      slides in registers (static structure: one base register + immediate offsets)
                if (seg_base >= 0) {
                    const uint32_t xaddr = smem_u32(xs + seg_base * 128 + lane * 4);   // x stage row of tap 0 of output 0
                    // all kSegLen+6 input rows of the segment are fetched up front (packed 16-bit pairs, one
                    // register each) so that their shared-memory latency is paid once, not per row
                    uint32_t xr[Cfg::kSegLen + 6];
                    if (Cfg::kResMma) {
                        const uint32_t xb0 = smem_u32(xs);
#pragma unroll
                        for (int j = 0; j < Cfg::kSegLen + 6; ++j)
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(xr[j]) : "r"(xb0 + xswz[j & 7] + j * DIL * 128));
                    } else {
#pragma unroll
                        for (int j = 0; j < Cfg::kSegLen + 6; ++j)
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(xr[j]) : "r"(xaddr + j * DIL * 128));
                    }
                    float2 win[7];
#pragma unroll
                    for (int j = 0; j < 6; ++j)
                        win[j + 1] = snake_pair(unpack2(xr[j], static_cast<const HT*>(nullptr)), al1, ia1);
#pragma unroll
                    for (int i = 0; i < Cfg::kSegLen; ++i) {
#pragma unroll
                        for (int j = 0; j < 6; ++j) win[j] = win[j + 1];
                        win[6] = snake_pair(unpack2(xr[i + 6], static_cast<const HT*>(nullptr)), al1, ia1);
                        float2 acc = bd;
#pragma unroll
                        for (int j = 0; j < 7; ++j) acc = ffma2(w[j], win[j], acc);
                        acc = snake_pair(acc, al2, ia2);
                        const int r = seg_base + i * DIL;
                        *reinterpret_cast<uint32_t*>(dstA + (r >> 7) * 16384 + sw128_offset(r & 127, 2 * lane)) =
                            pack2(acc.x, acc.y, static_cast<const HT*>(nullptr));
                    }
                }
                // per-warp hand-off (no CTA-wide barrier): make this warp's generic-proxy writes visible to the
                // tensor core, then one lane arrives for the warp on the A-stage and x-stage barriers
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&a_full[sa]); mbar_arrive(&x_empty[sx]); }
                if (pr) { const long long t = clock64(); a.prof[5] += t - tq; tq = t; }
            }
            if (!Cfg::kSplitEpi && warp < kLockstepEpiWarps) {
                const long long te = clock64();
                if (Cfg::kAccStages == 1) epilogue(n);
                else if (n > 0) epilogue(n - 1);
                if (a.prof && blockIdx.x == 0 && tid == 0) { a.prof[6] += clock64() - te; a.prof[7] += 1; }
            }
            // kResMma: the epilogue staged its output in the A stages -- no prologue warp (16 and 17 run no epilogue) may
            // write the next tile's operand before every store has read its slot
            if (Cfg::kResMma) asm volatile("bar.sync 2, %0;" ::"n"(kComputeWarps * 32) : "memory");
        }
        if (!Cfg::kSplitEpi && Cfg::kAccStages == 2 && my_tiles > 0 && warp < kLockstepEpiWarps) epilogue(my_tiles - 1);
        if (Cfg::kResMma && warp < kLockstepEpiWarps && lane == 0) bulk_wait_group<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kComputeWarps + 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int C, int DIL, int EPI, typename HT>
cudaError_t launch_t(const ResUnitArgs& a, const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmO, int sm_count,
                     cudaStream_t st) {
    using Cfg = Res2Cfg<C, DIL>;
    static PerDeviceOnce once;
    int dev_;
    if (once.needed(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(k_resunit2<C, DIL, EPI, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        once.done(dev_);
    }
    const int tiles = a.S * (((a.t_n > 0 ? a.t_n : a.T) + Cfg::kTileM - 1) / Cfg::kTileM);
    if (tiles == 0) return cudaSuccess;
    const int grid = tiles < sm_count ? tiles : sm_count;
    k_resunit2<C, DIL, EPI, HT><<<grid, Cfg::kThreads, Cfg::kSmem, st>>>(tmX, tmW, tmO, a, tiles);
    return cudaGetLastError();
}

template <int C, typename HT>
cudaError_t launch_c(const ResUnitArgs& a, const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmO, int sm_count,
                     cudaStream_t st) {
    if (a.dil == 1) return launch_t<C, 1, EPI_RES, HT>(a, tmX, tmW, tmO, sm_count, st);
    if (a.dil == 3) return launch_t<C, 3, EPI_RES, HT>(a, tmX, tmW, tmO, sm_count, st);
    if (a.dil == 9) return launch_t<C, 9, EPI_RES_SNAKE, HT>(a, tmX, tmW, tmO, sm_count, st);
    return cudaErrorInvalidValue;
}

template <typename HT>
cudaError_t launch_h(const ResUnitArgs& a, const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmO, int sm_count,
                     cudaStream_t st) {
    switch (a.C) {
        case 512: return launch_c<512, HT>(a, tmX, tmW, tmO, sm_count, st);
        case 256: return launch_c<256, HT>(a, tmX, tmW, tmO, sm_count, st);
        case 128: return launch_c<128, HT>(a, tmX, tmW, tmO, sm_count, st);
        case 64: return launch_c<64, HT>(a, tmX, tmW, tmO, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

void resunit2_geometry(int C, int dil, int* tile_m, int* box_rows) {
    const int tm = (C == 512) ? 128 : 256;
    *tile_m = tm;
    *box_rows = (C == 512) ? tm + 6 * dil : (tm + 6 * dil) / 2;     // C = 512: one 128B-swizzled box (resunit2_swizzled_x)
}

bool resunit2_swizzled_x(int C) { return C == 512; }

// dil 1 and 3 use the plain residual epilogue, dil 9 (last unit of a block) applies the next Snake.
// tmO (C = 512 only, else unused): the output, box (32, 32, 1), 64B-swizzled
cudaError_t launch_resunit2(int half_fp16, const ResUnitArgs& a, const CUtensorMap& tmX, const CUtensorMap& tmW,
                            const CUtensorMap& tmO, int sm_count, cudaStream_t st) {
    return half_fp16 ? launch_h<__half>(a, tmX, tmW, tmO, sm_count, st) : launch_h<__nv_bfloat16>(a, tmX, tmW, tmO, sm_count, st);
}

}  // namespace snacb
