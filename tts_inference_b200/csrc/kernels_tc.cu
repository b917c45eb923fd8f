// Tensor-core kernels of the SNAC decode path, hand-written for sm_100a:
//   tcgen05.mma (bf16 x bf16 -> fp32 accumulators in TMEM), TMA tile loads (cp.async.bulk.tensor,
//   128-byte swizzle), mbarrier pipelines, tcgen05.ld epilogues.
//
//   k_gemm_tc     persistent warp-specialised row-GEMM with taps: stem 1x1 (768->1024),
//                 ConvTranspose1d as a 2-tap GEMM per output phase, NoiseBlock 1x1.
//                 A (activations, channel-last bf16) and W both arrive by TMA; the time shift of a
//                 tap is a TMA coordinate, stream boundaries are TMA out-of-bounds zero fill.
//   k_resunit_tc  ResidualUnit: Snake -> depthwise dilated k7 -> Snake computed on CUDA cores into the
//                 swizzled A operand in shared memory, 1x1 conv on tcgen05, bias + residual (+ next
//                 Snake) in the epilogue.
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace snacb {
using namespace ptx;

constexpr int kTileM = 128;
constexpr int kChunkK = 64;                 // bf16 elements per 128-byte swizzle row
constexpr int kABytes = kTileM * 128;       // one A stage: 128 rows x 64 bf16

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~static_cast<uintptr_t>(1023));
}

__device__ __forceinline__ float2 unpack2t(uint32_t v, const __half*) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ float2 unpack2t(uint32_t v, const __nv_bfloat16*) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

template <typename HT> struct HalfFmt;
template <> struct HalfFmt<__half> { static constexpr uint32_t kFmt = 0; };
template <> struct HalfFmt<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; };

// =================================================================================================
// k_gemm_tc
// =================================================================================================
template <int BN>
struct GemmTcCfg {
    static constexpr int kStages = (BN == 128) ? 3 : 4;   // BN<256: two CTAs per SM fit
    static constexpr int kWBytes = BN * 128;
    static constexpr int kStageBytes = kABytes + kWBytes;
    static constexpr int kBarBytes = 256;
    // NoiseBlock epilogue by TMA (BN = 256 only: one CTA per SM, 33 KB of shared memory to spare): per epilogue warp two
    // 2 KB tiles for y (32 rows x 32 columns, 64B-swizzled, loaded two pieces ahead) and two for the output (a piece is
    // written while the previous piece's store is still reading its tile)
    static constexpr int kEpiYBufs = 2;
    static constexpr int kEpiWarpBytes = (kEpiYBufs + 2) * 2048;
    static constexpr int kEpiBytes = (BN == 256) ? 4 * kEpiWarpBytes : 0;
    static constexpr int kOffEpi = kStages * kStageBytes + 1024;
    static constexpr int kSmem = kStages * kStageBytes + (kEpiBytes ? 1024 + kEpiBytes : kBarBytes) + 1024;
    static constexpr int kTmemCols = 2 * BN;   // two accumulator stages
};

template <int BN, int EPI, typename HT, typename OutT>
__global__ void __launch_bounds__(192, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
          const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmO, const GemmArgs a,
          const int num_m_tiles, const int num_n_tiles) {
    using Cfg = GemmTcCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::kStages;
    uint64_t* tfull = bars + 2 * Cfg::kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    uint64_t* ybar = bars + 16;                     // [4 epilogue warps][kEpiYBufs]: a y staging tile has landed (TMA epilogue)
    constexpr bool kTmaEpi = (BN == 256 && EPI == EPI_NOISE && sizeof(OutT) == 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmW);
        if (kTmaEpi && a.tma_epi) { prefetch_tmap(&tmY); prefetch_tmap(&tmO); }
        for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        for (int i = 0; i < 4 * Cfg::kEpiYBufs; ++i) mbar_init(&ybar[i], 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = num_m_tiles * num_n_tiles;
    const int t_lo = a.t_n > 0 ? a.t_lo : 0, t_n = a.t_n > 0 ? a.t_n : a.Tin;
    const int tiles_t = (t_n + a.Tbox - 1) / a.Tbox;
    const int kchunks = a.K / kChunkK;
    const int nk = a.ntaps * kchunks;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / num_n_tiles, nt = tile % num_n_tiles;
                const int sg = (mt / tiles_t) * a.Wbox;           // first stream of the tile
                const bool mapped = a.map.slot != nullptr || a.map.off != nullptr;
                const int s0 = sm_slot(a.map, sg);
                const int t0 = t_lo + sm_off(a.map, sg, a.rpf) + (mt % tiles_t) * a.Tbox;
                const int n0 = nt * BN;
                const int p = n0 / a.Cout;
                const int base_shift = (a.up > 1 && p >= a.up / 2) ? 1 : 0;
                for (int kk = 0; kk < nk; ++kk) {
                    const int tap = kk / kchunks, kc = kk % kchunks;
                    mbar_wait(&empty[stage], phase ^ 1u);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    const int ka = (a.a_wrap > 0 ? kc % a.a_wrap : kc) * kChunkK;
                    if (mapped && a.Wbox > 1) {
                        // session launch with several streams per tile: every stream has its own rows and slot, so its
                        // Tbox rows arrive by their own load (the A map's box is then (64, Tbox, 1))
                        const int nv = (a.S - sg) < a.Wbox ? (a.S - sg) : a.Wbox;
                        mbar_expect_tx(&full[stage], static_cast<uint32_t>(nv * a.Tbox * 128 + BN * 128));
                        for (int j = 0; j < nv; ++j)
                            tma_load_3d(sa + j * a.Tbox * 128, &tmA, ka,
                                        t_lo + sm_off(a.map, sg + j, a.rpf) + (mt % tiles_t) * a.Tbox + base_shift - tap,
                                        sm_slot(a.map, sg + j), &full[stage]);
                    } else {
                        mbar_expect_tx(&full[stage], Cfg::kStageBytes);
                        tma_load_3d(sa, &tmA, ka, t0 + base_shift - tap, s0, &full[stage]);
                    }
                    tma_load_2d_hint(sa + kABytes, &tmW, tap * a.K + kc * kChunkK, n0, &full[stage], kL2EvictLast);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_f16(kTileM, BN, a.mma_bf16 ? 1u : HalfFmt<HT>::kFmt);
        int stage = 0; uint32_t phase = 0;
        int as = 0; uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[as], aphase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
            for (int kk = 0; kk < nk; ++kk) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint32_t sw = sa + kABytes;
#pragma unroll
                    for (int k = 0; k < kChunkK / 16; ++k)
                        mma_f16_ss(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sw + k * 32), idesc,
                                    (kk > 0 || k > 0) ? 1u : 0u);
                    mma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
            }
            if (lane == 0) mma_commit(&tfull[as]);
            __syncwarp();
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (4 warps)
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const int r = q * 32 + lane;                  // tile row = TMEM lane
        int as = 0; uint32_t aphase = 0;
        unsigned long long key = 0;
        if (EPI == EPI_NOISE && a.noise == nullptr)
            key = splitmix64(a.seed * 0x100000001B3ull + static_cast<unsigned long long>(100 + a.noise_stage));
        OutT* out = static_cast<OutT*>(a.out);
        uint32_t yph = 0;                             // phase bits of this warp's two y barriers
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / num_n_tiles, nt = tile % num_n_tiles;
            const int s = (mt / tiles_t) * a.Wbox + r / a.Tbox;        // stream of the launch (noise key, row offset)
            const int so = (s < a.S) ? sm_off(a.map, s, a.rpf) : 0;
            const int sl = (s < a.S) ? sm_slot(a.map, s) : 0;          // its buffer slot
            const int m = t_lo + so + (mt % tiles_t) * a.Tbox + r % a.Tbox;
            const bool valid = (s < a.S) && (m < a.Tin) && (m < t_lo + so + t_n);
            const int n0 = nt * BN;
            float nz = 0.f;
            if (EPI == EPI_NOISE && valid)
                nz = a.noise ? a.noise[static_cast<size_t>(s) * a.Tin + m]
                             : counter_normal(key, noise_counter(a.stream_keys ? a.stream_keys[s] : a.stream_offset + s,
                                                                 m + a.t0 + sm_org(a.map, s, a.rpf)));
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
            if constexpr (kTmaEpi) {
                // NoiseBlock epilogue by TMA (a.tma_epi: Tbox = 128, one stream per tile, up = 1).  Row-per-lane global
                // accesses touch 32 cache lines per instruction and held this epilogue at ~1.9 k cycles per 32-column
                // piece against a 4.3 k-cycle MMA per tile; here y arrives 64B-swizzled in the warp's own 2 KB tiles (two
                // pieces ahead) and x leaves through a third by one TMA store per piece.  A 32-row group cut by the end
                // of the row range keeps the per-row path below.
                const int mw = t_lo + so + (mt % tiles_t) * a.Tbox + q * 32;
                const int m_end = (a.Tin < t_lo + so + t_n) ? a.Tin : t_lo + so + t_n;
                if (a.tma_epi && s < a.S && mw + 32 <= m_end) {
                    const int wq = warp - 2;
                    uint8_t* ystg = smem + Cfg::kOffEpi + wq * Cfg::kEpiWarpBytes;
                    constexpr int NY = Cfg::kEpiYBufs;
                    uint8_t* ostg = ystg + NY * 2048;
                    auto yload = [&](int c) {
                        if (lane == 0) {
                            uint64_t* bar = &ybar[wq * NY + (c % NY)];
                            mbar_expect_tx(bar, 2048);
                            tma_load_3d(ystg + (c % NY) * 2048, &tmY, n0 + c * 32, mw, sl, bar);
                        }
                    };
#pragma unroll
                    for (int c = 0; c < NY; ++c) yload(c);
                    mbar_wait(&tfull[as], aphase);
                    tc_fence_after();
                    const uint32_t sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll 1
                    for (int c = 0; c < BN / 32; ++c) {
                        const int buf = c % NY;
                        uint32_t raw[32];
                        tmem_ld32(taddr + c * 32, raw);
                        mbar_wait(&ybar[wq * NY + buf], (yph >> buf) & 1u);
                        yph ^= 1u << buf;
                        uint4 yv[4];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc)
                            yv[cc] = *reinterpret_cast<const uint4*>(ystg + buf * 2048 + lane * 64 + ((cc ^ sw) << 4));
                        tmem_ld_wait();
                        const uint32_t* yw = reinterpret_cast<const uint32_t*>(yv);
                        uint32_t ow[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 y = unpack2t(yw[j], static_cast<const HT*>(nullptr));
                            ow[j] = pack2(fmaf(nz, __uint_as_float(raw[2 * j]), y.x), fmaf(nz, __uint_as_float(raw[2 * j + 1]), y.y),
                                          static_cast<const OutT*>(nullptr));
                        }
                        if (lane == 0) bulk_wait_group_read<1>();        // the store before the previous one has read this tile
                        __syncwarp();
                        uint8_t* ob = ostg + (c & 1) * 2048;
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc)
                            *reinterpret_cast<uint4*>(ob + lane * 64 + ((cc ^ sw) << 4)) =
                                make_uint4(ow[4 * cc], ow[4 * cc + 1], ow[4 * cc + 2], ow[4 * cc + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_3d(&tmO, ob, n0 + c * 32, mw, sl);
                            bulk_commit_group();
                        }
                        if (c + NY < BN / 32) yload(c + NY);             // every lane has read y tile `buf` (fence above)
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                    if (++as == 2) { as = 0; aphase ^= 1u; }
                    continue;
                }
            }
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t raw[32];
                tmem_ld32(taddr + c * 32, raw);
                tmem_ld_wait();
                if (valid) {
                    const int n = n0 + c * 32;
                    const int pc = n / a.Cout, o = n - pc * a.Cout;
                    const size_t orow = (static_cast<size_t>(sl) * a.Tin + m) * a.up + pc;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                    if (EPI == EPI_BIAS || EPI == EPI_BIAS_SNAKE || EPI == EPI_RES || EPI == EPI_RES_SNAKE) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + o + j));
                            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                        }
                    }
                    if (EPI == EPI_RES || EPI == EPI_RES_SNAKE) {           // ResidualUnit: x + (W a + b)
                        float y[32];
                        load32(static_cast<const HT*>(a.resid) + orow * a.Cout + o, y);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += y[j];
                    }
                    if (EPI == EPI_NOISE) {
                        float y[32];
                        load32(static_cast<const HT*>(a.resid) + orow * a.Cout + o, y);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaf(nz, v[j], y[j]);
                    }
                    if (EPI == EPI_BIAS_SNAKE || EPI == EPI_RES_SNAKE) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 al = __ldg(reinterpret_cast<const float4*>(a.alpha + o + j));
                            const float4 ia = __ldg(reinterpret_cast<const float4*>(a.inv_alpha + o + j));
                            v[j] = snake_f<true>(v[j], al.x, ia.x);
                            v[j + 1] = snake_f<true>(v[j + 1], al.y, ia.y);
                            v[j + 2] = snake_f<true>(v[j + 2], al.z, ia.z);
                            v[j + 3] = snake_f<true>(v[j + 3], al.w, ia.w);
                        }
                    }
                    store32(out + orow * a.Cout + o, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
        if (kTmaEpi && lane == 0) bulk_wait_group<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

int gemm_tc_block_n(const GemmArgs& a) {
    const int span = (a.up > 1) ? (a.up / 2) * a.Cout : a.N;   // columns that share one tap shift
    if (span % 256 == 0) return 256;
    if (span % 128 == 0) return 128;
    return 64;
}

template <int BN, int EPI, typename HT, typename OutT>
static cudaError_t launch_gemm_tc_t(const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmY,
                                    const CUtensorMap& tmO, int sm_count, cudaStream_t st) {
    using Cfg = GemmTcCfg<BN>;
    static PerDeviceOnce once;
    int dev_;
    if (once.needed(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BN, EPI, HT, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmem);
        if (e != cudaSuccess) return e;
        once.done(dev_);
    }
    const int tiles_t = ((a.t_n > 0 ? a.t_n : a.Tin) + a.Tbox - 1) / a.Tbox;
    const int num_m = ((a.S + a.Wbox - 1) / a.Wbox) * tiles_t;
    const int num_n = a.N / BN;
    const int total = num_m * num_n;
    const int ctas_per_sm = (BN == 256) ? 1 : 2;
    int grid = total < sm_count * ctas_per_sm ? total : sm_count * ctas_per_sm;
    if (grid < 1) grid = 1;
    k_gemm_tc<BN, EPI, HT, OutT><<<grid, 192, Cfg::kSmem, st>>>(tmA, tmW, tmY, tmO, a, num_m, num_n);
    return cudaGetLastError();
}

template <int EPI, typename HT, typename OutT>
static cudaError_t launch_gemm_tc_e(const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmY,
                                    const CUtensorMap& tmO, int sm_count, cudaStream_t st) {
    switch (gemm_tc_block_n(a)) {
        case 256: return launch_gemm_tc_t<256, EPI, HT, OutT>(a, tmA, tmW, tmY, tmO, sm_count, st);
        case 128: return launch_gemm_tc_t<128, EPI, HT, OutT>(a, tmA, tmW, tmY, tmO, sm_count, st);
        default: return launch_gemm_tc_t<64, EPI, HT, OutT>(a, tmA, tmW, tmY, tmO, sm_count, st);
    }
}

template <typename HT>
static cudaError_t launch_gemm_tc_h(int epi, int out_f32, const GemmArgs& a, const CUtensorMap& tmA,
                                    const CUtensorMap& tmW, const CUtensorMap& tmY, const CUtensorMap& tmO, int sm_count,
                                    cudaStream_t st) {
    if (epi == EPI_BIAS) return launch_gemm_tc_e<EPI_BIAS, HT, HT>(a, tmA, tmW, tmY, tmO, sm_count, st);
    if (epi == EPI_BIAS_SNAKE) return launch_gemm_tc_e<EPI_BIAS_SNAKE, HT, HT>(a, tmA, tmW, tmY, tmO, sm_count, st);
    if (epi == EPI_NOISE) {
        if (out_f32) return launch_gemm_tc_e<EPI_NOISE, HT, float>(a, tmA, tmW, tmY, tmO, sm_count, st);
        return launch_gemm_tc_e<EPI_NOISE, HT, HT>(a, tmA, tmW, tmY, tmO, sm_count, st);
    }
    if (epi == EPI_RES) return launch_gemm_tc_e<EPI_RES, HT, HT>(a, tmA, tmW, tmY, tmO, sm_count, st);
    if (epi == EPI_RES_SNAKE) return launch_gemm_tc_e<EPI_RES_SNAKE, HT, HT>(a, tmA, tmW, tmY, tmO, sm_count, st);
    return cudaErrorInvalidValue;
}
// tmY / tmO (GemmArgs::tma_epi, NoiseBlock only; else any valid map): resid and out as [slot][Tin][Cout], box (32, 32, 1), 64B-swizzled
cudaError_t launch_gemm_tc(int epi, int half_fp16, int out_f32, const GemmArgs& a, const CUtensorMap& tmA,
                           const CUtensorMap& tmW, const CUtensorMap& tmY, const CUtensorMap& tmO, int sm_count, cudaStream_t st) {
    return half_fp16 ? launch_gemm_tc_h<__half>(epi, out_f32, a, tmA, tmW, tmY, tmO, sm_count, st)
                     : launch_gemm_tc_h<__nv_bfloat16>(epi, out_f32, a, tmA, tmW, tmY, tmO, sm_count, st);
}

// =================================================================================================
// k_resunit_tc
// =================================================================================================
template <int C>
struct ResTcCfg {
    static constexpr int kChunks = C / kChunkK;
    static constexpr int kNSA = kChunks < 2 ? 1 : 2;          // A-operand stages
    static constexpr int kNSW = kChunks < 2 ? 1 : 2;          // weight stages
    static constexpr int kWBytes = C * 128;                   // [C rows (N)] x 64 bf16
    static constexpr int kNHalf = C > 256 ? 2 : 1;            // N split into <=256-wide MMAs
    static constexpr int kMmaN = C > 256 ? 256 : C;
    static constexpr int kBarBytes = 256;
    static constexpr int smem_bytes(int dil) {
        return kNSA * kABytes + kNSW * kWBytes + (kTileM + 6 * dil) * 64 * 4 + kBarBytes + 1024;
    }
};

template <int C, int EPI, typename HT, typename XT, typename OutT>
__global__ void __launch_bounds__(320, (C <= 128) ? 2 : 1)
k_resunit_tc(const __grid_constant__ CUtensorMap tmW, const ResUnitArgs a) {
    using Cfg = ResTcCfg<C>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    uint8_t* sA = smem;
    uint8_t* sW = sA + Cfg::kNSA * kABytes;
    float* s1 = reinterpret_cast<float*>(sW + Cfg::kNSW * Cfg::kWBytes);
    const int R = kTileM + 6 * a.dil;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s1) + R * 64 * 4);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* w_full = bars + 4;        // [2]
    uint64_t* w_empty = bars + 6;       // [2]
    uint64_t* acc_full = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tiles_t = (a.T + kTileM - 1) / kTileM;
    const int s = blockIdx.x / tiles_t;
    const int t0 = (blockIdx.x % tiles_t) * kTileM;

    if (tid == 0) {
        prefetch_tmap(&tmW);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1);
            mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 9) { tmem_alloc(tmem_slot, C); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const XT* x = static_cast<const XT*>(a.x);
    const size_t row0 = static_cast<size_t>(s) * a.T;

    if (warp == 8) {
        // ------------------------------------------------------------------ weight producer (TMA)
        if (lane == 0) {
            for (int kc = 0; kc < Cfg::kChunks; ++kc) {
                const int sw = kc % Cfg::kNSW;
                if (kc >= Cfg::kNSW) mbar_wait(&w_empty[sw], ((kc / Cfg::kNSW) - 1) & 1);
                mbar_expect_tx(&w_full[sw], Cfg::kWBytes);
#pragma unroll
                for (int nh = 0; nh < Cfg::kNHalf; ++nh)
                    tma_load_2d(sW + sw * Cfg::kWBytes + nh * (256 * 128), &tmW, kc * kChunkK, nh * 256, &w_full[sw]);
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(kTileM, Cfg::kMmaN, HalfFmt<HT>::kFmt);
        for (int kc = 0; kc < Cfg::kChunks; ++kc) {
            const int sa = kc % Cfg::kNSA, sw = kc % Cfg::kNSW;
            mbar_wait(&w_full[sw], (kc / Cfg::kNSW) & 1);
            mbar_wait(&a_full[sa], (kc / Cfg::kNSA) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(sA + sa * kABytes);
                const uint32_t w_addr = smem_u32(sW + sw * Cfg::kWBytes);
#pragma unroll
                for (int k = 0; k < kChunkK / 16; ++k) {
#pragma unroll
                    for (int nh = 0; nh < Cfg::kNHalf; ++nh)
                        mma_f16_ss(tmem_base + nh * 256, umma_desc_sw128(a_addr + k * 32),
                                    umma_desc_sw128(w_addr + nh * (256 * 128) + k * 32), idesc,
                                    (kc > 0 || k > 0) ? 1u : 0u);
                }
                mma_commit(&a_empty[sa]);
                mma_commit(&w_empty[sw]);
                if (kc == Cfg::kChunks - 1) mma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ 8 compute warps
        // prologue: per 64-channel chunk, build the bf16 A operand  snake2(dw(snake1(x)))
        const int g = tid & 7, rr = tid >> 3;
        for (int kc = 0; kc < Cfg::kChunks; ++kc) {
            const int c0 = kc * kChunkK;
            {   // phase A: x (+halo) -> snake1 -> s1 (fp32)
                float al[8], ia[8];
                load8(a.alpha1 + c0 + 8 * g, al);
                load8(a.inv_alpha1 + c0 + 8 * g, ia);
#pragma unroll 3
                for (int r = rr; r < R; r += 32) {
                    const int t = t0 - 3 * a.dil + r;
                    float v[8];
                    if (t >= 0 && t < a.T) {
                        load8(x + (row0 + t) * C + c0 + 8 * g, v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = snake_f<true>(v[j], al[j], ia[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = 0.f;
                    }
                    store8(s1 + r * 64 + 8 * g, v);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int sa = kc % Cfg::kNSA;
            if (kc >= Cfg::kNSA) mbar_wait(&a_empty[sa], ((kc / Cfg::kNSA) - 1) & 1);
            {   // phase B: depthwise dilated k7 + bias -> snake2 -> swizzled bf16 A stage
                const int ch = c0 + 2 * lane;
                float w0[7], w1[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const float2 wv = *reinterpret_cast<const float2*>(a.dw_w + j * C + ch);
                    w0[j] = wv.x; w1[j] = wv.y;
                }
                const float2 bd = *reinterpret_cast<const float2*>(a.dw_b + ch);
                const float2 a2 = *reinterpret_cast<const float2*>(a.alpha2 + ch);
                const float2 i2 = *reinterpret_cast<const float2*>(a.inv_alpha2 + ch);
                uint8_t* dstA = sA + sa * kABytes;
                const int dil = a.dil;
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int r = warp * 16 + i;
                    float acc0 = bd.x, acc1 = bd.y;
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        const float2 sv = *reinterpret_cast<const float2*>(s1 + (r + j * dil) * 64 + 2 * lane);
                        acc0 = fmaf(w0[j], sv.x, acc0);
                        acc1 = fmaf(w1[j], sv.y, acc1);
                    }
                    acc0 = snake_f<true>(acc0, a2.x, i2.x);
                    acc1 = snake_f<true>(acc1, a2.y, i2.y);
                    *reinterpret_cast<uint32_t*>(dstA + sw128_offset(r, 2 * lane)) =
                        pack2(acc0, acc1, static_cast<const HT*>(nullptr));
                }
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) mbar_arrive(&a_full[sa]);
        }

        // epilogue: out = x + acc + bias (+ next Snake)
        const int q = warp & 3, h = warp >> 2;
        const int r = q * 32 + lane;
        const int t = t0 + r;
        const bool valid = t < a.T;
        OutT* out = static_cast<OutT*>(a.out);
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < C / 64; ++cc) {
            const int col = h * (C / 2) + cc * 32;
            uint32_t raw[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col, raw);
            tmem_ld_wait();
            if (valid) {
                float v[32], xr[32];
                load32(x + (row0 + t) * C + col, xr);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(a.pw_b + col + j));
                    v[j] = xr[j] + (__uint_as_float(raw[j]) + b.x);
                    v[j + 1] = xr[j + 1] + (__uint_as_float(raw[j + 1]) + b.y);
                    v[j + 2] = xr[j + 2] + (__uint_as_float(raw[j + 2]) + b.z);
                    v[j + 3] = xr[j + 3] + (__uint_as_float(raw[j + 3]) + b.w);
                }
                if (EPI == EPI_RES_SNAKE) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 al = __ldg(reinterpret_cast<const float4*>(a.alpha_next + col + j));
                        const float4 ia = __ldg(reinterpret_cast<const float4*>(a.inv_alpha_next + col + j));
                        v[j] = snake_f<true>(v[j], al.x, ia.x);
                        v[j + 1] = snake_f<true>(v[j + 1], al.y, ia.y);
                        v[j + 2] = snake_f<true>(v[j + 2], al.z, ia.z);
                        v[j + 3] = snake_f<true>(v[j + 3], al.w, ia.w);
                    }
                }
                store32(out + (row0 + t) * C + col, v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, C);
}

template <int C, int EPI, typename HT, typename XT, typename OutT>
static cudaError_t launch_resunit_tc_t(const ResUnitArgs& a, const CUtensorMap& tmW, cudaStream_t st) {
    using Cfg = ResTcCfg<C>;
    static PerDeviceOnce once;
    const int smem = Cfg::smem_bytes(a.dil);
    int dev_;
    if (once.needed(&dev_)) {
        cudaError_t e = cudaFuncSetAttribute(k_resunit_tc<C, EPI, HT, XT, OutT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes(9));
        if (e != cudaSuccess) return e;
        once.done(dev_);
    }
    const int tiles = a.S * ((a.T + kTileM - 1) / kTileM);
    if (tiles == 0) return cudaSuccess;
    k_resunit_tc<C, EPI, HT, XT, OutT><<<tiles, 320, smem, st>>>(tmW, a);
    return cudaGetLastError();
}

template <int EPI, typename HT, typename XT, typename OutT>
static cudaError_t launch_resunit_tc_c(const ResUnitArgs& a, const CUtensorMap& tmW, cudaStream_t st) {
    switch (a.C) {
        case 512: return launch_resunit_tc_t<512, EPI, HT, XT, OutT>(a, tmW, st);
        case 256: return launch_resunit_tc_t<256, EPI, HT, XT, OutT>(a, tmW, st);
        case 128: return launch_resunit_tc_t<128, EPI, HT, XT, OutT>(a, tmW, st);
        case 64: return launch_resunit_tc_t<64, EPI, HT, XT, OutT>(a, tmW, st);
        default: return cudaErrorInvalidValue;
    }
}

template <typename HT>
static cudaError_t launch_resunit_tc_h(int epi, int x_f32, const ResUnitArgs& a, const CUtensorMap& tmW,
                                       cudaStream_t st) {
    if (epi == EPI_RES) {
        if (x_f32) return launch_resunit_tc_c<EPI_RES, HT, float, float>(a, tmW, st);
        return launch_resunit_tc_c<EPI_RES, HT, HT, HT>(a, tmW, st);
    }
    if (epi == EPI_RES_SNAKE) {
        if (x_f32) return launch_resunit_tc_c<EPI_RES_SNAKE, HT, float, HT>(a, tmW, st);
        return launch_resunit_tc_c<EPI_RES_SNAKE, HT, HT, HT>(a, tmW, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_resunit_tc(int epi, int half_fp16, int x_f32, const ResUnitArgs& a, const CUtensorMap& tmW,
                              cudaStream_t st) {
    return half_fp16 ? launch_resunit_tc_h<__half>(epi, x_f32, a, tmW, st)
                     : launch_resunit_tc_h<__nv_bfloat16>(epi, x_f32, a, tmW, st);
}

cudaError_t init_tc_kernels() { return cudaSuccess; }

}  // namespace snacb
