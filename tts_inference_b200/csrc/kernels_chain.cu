// k_chain: one DecoderBlock's NoiseBlock + three ResidualUnits (+ the next layer's Snake) fused in ONE
// persistent kernel for sm_100a.  A tile of NB*128 time rows of one stream stays on the SM through all
// four layers:
//
//   x1 = y + n[t] * (Wn y)                                   NoiseBlock   (snac layers.py NoiseBlock)
//   x_{l+1} = x_l + W_l snake2(dw_d(snake1(x_l))) + b_l      d = 1, 3, 9  (ResidualUnit)
//   out = snake_next(x_4)
//
//   * the fp32 residual stream lives in TENSOR MEMORY: tcgen05.mma accumulates every layer's 1x1
//     conv directly on top of it (accumulate = 1); it is seeded with y by an MMA against a 64x64 identity;
//   * a single 16-bit copy of the tile lives in shared memory (128B-swizzled K-major, the UMMA A-operand
//     layout, filled by TMA).  Each layer rewrites it IN PLACE: the prologue turns x_l into the operand
//     snake2(dw(snake1(x_l))) with a 7-tap window sliding in registers along one dilation class (taps of row r
//     are r + j*d, so a class only ever reads its own rows; the 3 rows either side of a warp's span are
//     fetched before a CTA barrier), the MMAs consume it, and the epilogue writes x_{l+1} = TMEM + bias
//     back as 16-bit for the next prologue;
//   * a tile carries a 40-row halo either side (3*(1+3+9) = 39 rows of receptive field); halo results are
//     garbage by construction and never stored.  Rows outside [0, T) are forced to zero after every layer
//     (the convs' zero padding);
//   * HBM traffic: the ConvTranspose output is read once (+ halo), the block output written once.
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

template <typename HT> struct HalfFmtC;
template <> struct HalfFmtC<__half> { static constexpr uint32_t kFmt = 0; static constexpr uint16_t kOne = 0x3C00; };
template <> struct HalfFmtC<__nv_bfloat16> { static constexpr uint32_t kFmt = 1; static constexpr uint16_t kOne = 0x3F80; };

__device__ __forceinline__ float2 unpack2c(uint32_t v, const __half*) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ float2 unpack2c(uint32_t v, const __nv_bfloat16*) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

constexpr int kWarps = kChainWarps;          // 16 symmetric warps: prologue, epilogue; thread 0 also issues TMA / MMA
constexpr int kThreads = kWarps * 32;
constexpr int kHalo = kChainHalo;            // 40 >= 39, multiple of 8

template <int C, int NB>
struct ChainCfg {
    static constexpr int kCH = C / 64;                      // 64-channel K chunks
    static constexpr int kRows = NB * 128;                  // tile rows incl. halo
    static constexpr int kROut = kRows - 2 * kHalo;         // rows stored per tile
    static constexpr int kPlane = NB * 16384;               // one chunk plane of the tile [NB][128 rows][128 B]
    static constexpr int kXBytes = kCH * kPlane;
    static constexpr bool kWRes = (C == 64);                // all four 1x1 weights resident
    static constexpr int kWLayer = C * C * 2;               // one layer's weights [kCH][C rows][128 B]
    static constexpr int kWBytes = kWRes ? 4 * kWLayer : 2 * kWLayer;
    static constexpr int kIBytes = 8192;                    // 64 x 64 identity
    static constexpr int kPrmBytes = 3 * (C / 2) * 96;      // per layer and channel pair: 24 floats
    static constexpr int kEpiBytes = 5 * C * 4;             // bias_cum[3][C], alpha_next[C], inv_next[C]
    static constexpr int kOffX = 0;
    static constexpr int kOffW = kOffX + kXBytes;
    static constexpr int kOffI = kOffW + kWBytes;
    static constexpr int kOffPrm = kOffI + kIBytes;
    static constexpr int kOffEpi = kOffPrm + kPrmBytes;
    static constexpr int kOffBar = kOffEpi + kEpiBytes;
    static constexpr int kSmem = kOffBar + 128 + 1024;
    static constexpr int kTmemCols = NB * C;
    static_assert(kTmemCols == 512 || kTmemCols == 256 || kTmemCols == 128, "TMEM columns");
    static_assert(kSmem <= 232448, "shared memory budget");
    static_assert(kWBytes >= 16384, "tail pre-reads may run up to 90 rows past the tile");
    static_assert(kRows % kThreads == 0 || kThreads % kRows == 0, "noise row mapping");
};

}  // namespace

template <int C, int NB, typename HT>
__global__ void __launch_bounds__(kThreads, 1)
k_chain(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWn,
        const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
        const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ ChainArgs a, const int num_tiles) {
    using Cfg = ChainCfg<C, NB>;
    constexpr int CH = Cfg::kCH;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem + Cfg::kOffX;
    uint8_t* sW = smem + Cfg::kOffW;
    uint8_t* sI = smem + Cfg::kOffI;
    float* sPrm = reinterpret_cast<float*>(smem + Cfg::kOffPrm);
    float* sEpi = reinterpret_cast<float*>(smem + Cfg::kOffEpi);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
    uint64_t* ld_bar = bars;          // tile landed (TMA)
    uint64_t* mma_bar = bars + 1;     // all MMAs of a layer complete
    uint64_t* w_bar = bars + 2;       // [2] weight buffers landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tiles_t = (a.T + Cfg::kROut - 1) / Cfg::kROut;
    const int my_tiles = (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const CUtensorMap* wmaps[4] = {&tmWn, &tmW0, &tmW1, &tmW2};

    // ------------------------------------------------------------------ one-time setup
    if (tid == 0) {
        prefetch_tmap(&tmY); prefetch_tmap(&tmWn); prefetch_tmap(&tmW0); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
        mbar_init(ld_bar, 1); mbar_init(mma_bar, 1); mbar_init(&w_bar[0], 1); mbar_init(&w_bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
    for (int i = tid; i < 3 * (C / 2); i += kThreads) {            // per-layer prologue parameters, 24 floats per channel pair
        const int l = i / (C / 2), ch = 2 * (i % (C / 2));
        const ChainLayer& L = a.res[l];
        float* d = sPrm + i * 24;
        d[0] = L.alpha1[ch]; d[1] = L.alpha1[ch + 1];
        d[2] = L.inv1[ch]; d[3] = L.inv1[ch + 1];
#pragma unroll
        for (int j = 0; j < 7; ++j) { d[4 + 2 * j] = L.dw_w[j * C + ch]; d[5 + 2 * j] = L.dw_w[j * C + ch + 1]; }
        d[18] = L.dw_b[ch]; d[19] = L.dw_b[ch + 1];
        d[20] = L.alpha2[ch]; d[21] = L.alpha2[ch + 1];
        d[22] = L.inv2[ch]; d[23] = L.inv2[ch + 1];
    }
    for (int c = tid; c < 3 * C; c += kThreads) sEpi[c] = a.bias_cum[c];
    for (int c = tid; c < C; c += kThreads) { sEpi[3 * C + c] = a.alpha_next[c]; sEpi[4 * C + c] = a.inv_next[c]; }
    for (int i = tid; i < Cfg::kIBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(sI)[i] = 0u;
    __syncthreads();
    if (tid < 64) *reinterpret_cast<uint16_t*>(sI + sw128_offset(tid, tid)) = HalfFmtC<HT>::kOne;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // weight staging: resident (C = 64) or two rotating buffers, layer l of a tile uses buffer l & 1
    auto load_w = [&](int l, int buf) {       // thread 0
        uint8_t* dst = sW + (Cfg::kWRes ? l : buf) * Cfg::kWLayer;
        uint64_t* bar = &w_bar[Cfg::kWRes ? 0 : buf];
        if (!Cfg::kWRes) mbar_expect_tx(bar, Cfg::kWLayer);
#pragma unroll
        for (int kc = 0; kc < CH; ++kc) tma_load_2d(dst + kc * (C * 128), wmaps[l], kc * 64, 0, bar);
    };
    auto load_tile = [&](int n) {             // thread 0
        const int tile = blockIdx.x + n * gridDim.x;
        const int s = tile / tiles_t, t_start = (tile % tiles_t) * Cfg::kROut - kHalo;
        mbar_expect_tx(ld_bar, Cfg::kXBytes);
#pragma unroll
        for (int kc = 0; kc < CH; ++kc)
#pragma unroll
            for (int b = 0; b < NB; ++b)
                tma_load_3d(sX + kc * Cfg::kPlane + b * 16384, &tmY, kc * 64, t_start + b * 128, s, ld_bar);
    };
    if (tid == 0 && my_tiles > 0) {
        if (Cfg::kWRes) {
            mbar_expect_tx(&w_bar[0], 4 * Cfg::kWLayer);
            for (int l = 0; l < 4; ++l) load_w(l, 0);
        } else {
            load_w(0, 0);
            load_w(1, 1);
        }
        load_tile(0);
    }

    uint32_t mma_par = 0;
    constexpr uint32_t idescI = umma_idesc_f16(128, 64, HalfFmtC<HT>::kFmt);
    constexpr uint32_t idescW = umma_idesc_f16(128, C, HalfFmtC<HT>::kFmt);

    // swizzled byte offset of this lane's channel pair inside a 128-byte row, for each (row & 7)
    uint32_t swz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) swz[j] = ((((lane >> 2) ^ j) & 7) << 4) + ((lane & 3) << 2);
    const uint32_t sx_addr = smem_u32(sX);

    // issue one layer's 1x1 conv for the whole tile (thread 0): TMEM[blk] += A[blk] * W^T
    auto issue_layer = [&](int l, int n) {
        const int buf = l & 1;
        if (Cfg::kWRes) { if (n == 0 && l == 0) mbar_wait(&w_bar[0], 0); }
        else mbar_wait(&w_bar[buf], (2 * n + (l >> 1)) & 1);
        tc_fence_after();
        const uint32_t w_addr = smem_u32(sW + (Cfg::kWRes ? l : buf) * Cfg::kWLayer);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int kc = 0; kc < CH; ++kc)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_f16_ss(tmem_base + b * C, umma_desc_sw128(sx_addr + kc * Cfg::kPlane + b * 16384 + k * 32),
                               umma_desc_sw128(w_addr + kc * (C * 128) + k * 32), idescW, 1u);
        mma_commit(mma_bar);
    };
    // after layer l's MMAs completed: its weight buffer is free -> prefetch the layer two ahead (thread 0)
    auto prefetch_w = [&](int l, int n) {
        if (Cfg::kWRes) return;
        const int l2 = (l + 2) & 3;
        if (l + 2 < 4 || n + 1 < my_tiles) load_w(l2, l & 1);
    };

    // epilogue: TMEM (+ cumulative bias) -> 16-bit tile copy (kFinal = false) or Snake -> global (kFinal = true)
    auto epilogue = [&](const float* bias, bool has_bias, bool final_layer, int s, int t_start) {
        const int q = warp & 3, g = warp >> 2;
        constexpr int kPieces = NB * (C / 32);
        HT* out = static_cast<HT*>(a.out);
#pragma unroll 1
        for (int it = g; it < kPieces; it += 4) {
            const int blk = it / (C / 32), cg = it % (C / 32);
            uint32_t raw[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + blk * C + cg * 32, raw);
            tmem_ld_wait();
            const int i = blk * 128 + q * 32 + lane;
            const int t = t_start + i;
            const bool valid = static_cast<unsigned>(t) < static_cast<unsigned>(a.T);
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_bias) b = *reinterpret_cast<const float4*>(bias + cg * 32 + j);
                float v0 = __uint_as_float(raw[j]) + b.x, v1 = __uint_as_float(raw[j + 1]) + b.y;
                float v2 = __uint_as_float(raw[j + 2]) + b.z, v3 = __uint_as_float(raw[j + 3]) + b.w;
                if (final_layer) {
                    const float4 al = *reinterpret_cast<const float4*>(sEpi + 3 * C + cg * 32 + j);
                    const float4 ia = *reinterpret_cast<const float4*>(sEpi + 4 * C + cg * 32 + j);
                    v0 = snake_f<true>(v0, al.x, ia.x); v1 = snake_f<true>(v1, al.y, ia.y);
                    v2 = snake_f<true>(v2, al.z, ia.z); v3 = snake_f<true>(v3, al.w, ia.w);
                }
                o[j / 2] = pack2(v0, v1, static_cast<const HT*>(nullptr));
                o[j / 2 + 1] = pack2(v2, v3, static_cast<const HT*>(nullptr));
            }
            if (final_layer) {
                if (valid && i >= kHalo && i < Cfg::kRows - kHalo) {
                    uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(s) * a.T + t) * C + cg * 32);
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[c] = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                }
            } else {
                uint8_t* row = sX + (cg >> 1) * Cfg::kPlane + i * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int chunk = ((cg & 1) * 4 + c) ^ (i & 7);
                    *reinterpret_cast<uint4*>(row + chunk * 16) =
                        valid ? make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]) : make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
    };

    for (int n = 0; n < my_tiles; ++n) {
        const int tile = blockIdx.x + n * gridDim.x;
        const int s = tile / tiles_t, t_start = (tile % tiles_t) * Cfg::kROut - kHalo;

        // ---------------------------------------------------------------- tile landed -> TMEM = y (identity MMA)
        mbar_wait(ld_bar, n & 1);
        tc_fence_after();
        if (tid == 0) {
            const uint32_t i_addr = smem_u32(sI);
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int kc = 0; kc < CH; ++kc)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_f16_ss(tmem_base + b * C + kc * 64,
                                   umma_desc_sw128(sx_addr + kc * Cfg::kPlane + b * 16384 + k * 32),
                                   umma_desc_sw128(i_addr + k * 32), idescI, k > 0 ? 1u : 0u);
            mma_commit(mma_bar);
        }
        // noise values of this thread's rows (overlaps the identity MMA)
        constexpr int kRowsPerThread = (Cfg::kRows + kThreads - 1) / kThreads;
        float nzv[kRowsPerThread];
        {
            unsigned long long key = 0;
            if (a.noise == nullptr)
                key = splitmix64(a.seed * 0x100000001B3ull + static_cast<unsigned long long>(100 + a.noise_stage));
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) {
                const int i = tid + r * kThreads;
                const int t = t_start + i;
                float v = 0.f;
                if (i < Cfg::kRows && static_cast<unsigned>(t) < static_cast<unsigned>(a.T))
                    v = a.noise ? a.noise[static_cast<size_t>(s) * a.T + t]
                                : counter_normal(key, static_cast<unsigned long long>(a.stream_offset + s) * a.T + t);
                nzv[r] = v;
            }
        }
        mbar_wait(mma_bar, mma_par); mma_par ^= 1u;
        // ---------------------------------------------------------------- NoiseBlock: operand = n[t] * y, in place
#pragma unroll
        for (int r = 0; r < kRowsPerThread; ++r) {
            const int i = tid + r * kThreads;
            if (i < Cfg::kRows) {
                const float nz = nzv[r];
#pragma unroll
                for (int kc = 0; kc < CH; ++kc) {
                    uint8_t* row = sX + kc * Cfg::kPlane + i * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint4* p = reinterpret_cast<uint4*>(row + ((c ^ (i & 7)) << 4));
                        uint4 v = *p;
                        uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = unpack2c(w[e], static_cast<const HT*>(nullptr));
                            w[e] = pack2(f.x * nz, f.y * nz, static_cast<const HT*>(nullptr));
                        }
                        *p = v;
                    }
                }
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) issue_layer(0, n);
        mbar_wait(mma_bar, mma_par); mma_par ^= 1u;
        tc_fence_after();
        if (tid == 0) prefetch_w(0, n);
        epilogue(nullptr, false, false, s, t_start);
        tc_fence_before();
        __syncthreads();

        // ---------------------------------------------------------------- three ResidualUnits
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            const int d = (l == 0) ? 1 : (l == 1 ? 3 : 9);
            // ---- spans of this warp: pre-read the 3 rows before and after each span (owned by other warps)
            uint32_t hd[2][3], tl[2][3];
            int r_first[2], n_oct[2], kcs[2];
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
                const ChainSpan spn = a.spans[l][warp][sp];
                r_first[sp] = spn.r_first; n_oct[sp] = spn.n_oct; kcs[sp] = spn.kc;
                const uint8_t* plane = sX + spn.kc * Cfg::kPlane;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int rh = spn.r_first - (3 - j) * d;
                    const int rt = spn.r_first + (8 * spn.n_oct + j) * d;
                    uint32_t vh = 0u, vt = 0u;
                    if (spn.n_oct > 0) {
                        if (rh >= 0) vh = *reinterpret_cast<const uint32_t*>(plane + rh * 128 + ((((lane >> 2) ^ rh) & 7) << 4) + ((lane & 3) << 2));
                        vt = *reinterpret_cast<const uint32_t*>(plane + rt * 128 + ((((lane >> 2) ^ rt) & 7) << 4) + ((lane & 3) << 2));
                    }
                    hd[sp][j] = vh; tl[sp][j] = vt;
                }
            }
            __syncthreads();
#pragma unroll 1
            for (int sp = 0; sp < 2; ++sp) {
                const int nq = sp ? n_oct[1] : n_oct[0];
                if (nq == 0) continue;
                const int kc = sp ? kcs[1] : kcs[0];
                int r_oct = sp ? r_first[1] : r_first[0];
                const uint32_t h0 = sp ? hd[1][0] : hd[0][0], h1 = sp ? hd[1][1] : hd[0][1], h2 = sp ? hd[1][2] : hd[0][2];
                const uint32_t t0 = sp ? tl[1][0] : tl[0][0], t1 = sp ? tl[1][1] : tl[0][1], t2 = sp ? tl[1][2] : tl[0][2];
                // per-lane parameters of channels (kc*64 + 2*lane, +1)
                const float4* p4 = reinterpret_cast<const float4*>(sPrm + ((l * (C / 2)) + kc * 32 + lane) * 24);
                const float4 q0 = p4[0], q1 = p4[1], q2 = p4[2], q3 = p4[3], q4 = p4[4], q5 = p4[5];
                const float2 al1 = make_float2(q0.x, q0.y), ia1 = make_float2(q0.z, q0.w);
                float2 w[7];
                w[0] = make_float2(q1.x, q1.y); w[1] = make_float2(q1.z, q1.w);
                w[2] = make_float2(q2.x, q2.y); w[3] = make_float2(q2.z, q2.w);
                w[4] = make_float2(q3.x, q3.y); w[5] = make_float2(q3.z, q3.w);
                w[6] = make_float2(q4.x, q4.y);
                const float2 bd = make_float2(q4.z, q4.w);
                const float2 al2 = make_float2(q5.x, q5.y), ia2 = make_float2(q5.z, q5.w);

                uint8_t* plane = sX + kc * Cfg::kPlane;
                auto run = [&](auto dtag) {
                    constexpr int D = decltype(dtag)::value;
                    float2 win[7];
                    win[1] = snake_pair(unpack2c(h0, static_cast<const HT*>(nullptr)), al1, ia1);
                    win[2] = snake_pair(unpack2c(h1, static_cast<const HT*>(nullptr)), al1, ia1);
                    win[3] = snake_pair(unpack2c(h2, static_cast<const HT*>(nullptr)), al1, ia1);
                    uint8_t* ob = plane + r_oct * 128;             // r_oct = 0 (mod 8): (row & 7) of step k is (k*D) & 7
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        uint32_t raw = 0u;               // class starts may lie up to 24 rows above the tile
                        if (r_oct + j * D >= 0) raw = *reinterpret_cast<const uint32_t*>(ob + j * D * 128 + swz[(j * D) & 7]);
                        win[4 + j] = snake_pair(unpack2c(raw, static_cast<const HT*>(nullptr)), al1, ia1);
                    }
#pragma unroll 1
                    for (int qo = 0; qo < nq; ++qo) {
                        uint32_t raw[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            raw[k] = *reinterpret_cast<const uint32_t*>(ob + (k + 3) * D * 128 + swz[((k + 3) * D) & 7]);
                        if (qo == nq - 1) { raw[5] = t0; raw[6] = t1; raw[7] = t2; }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
#pragma unroll
                            for (int j = 0; j < 6; ++j) win[j] = win[j + 1];
                            win[6] = snake_pair(unpack2c(raw[k], static_cast<const HT*>(nullptr)), al1, ia1);
                            float2 acc = bd;
#pragma unroll
                            for (int j = 0; j < 7; ++j) acc = ffma2(w[j], win[j], acc);
                            acc = snake_pair(acc, al2, ia2);
                            const int r = r_oct + k * D;
                            if (static_cast<unsigned>(r) < static_cast<unsigned>(Cfg::kRows))
                                *reinterpret_cast<uint32_t*>(ob + k * D * 128 + swz[(k * D) & 7]) =
                                    pack2(acc.x, acc.y, static_cast<const HT*>(nullptr));
                        }
                        r_oct += 8 * D;
                        ob += 8 * D * 128;
                    }
                };
                if (d == 1) run(std::integral_constant<int, 1>{});
                else if (d == 3) run(std::integral_constant<int, 3>{});
                else run(std::integral_constant<int, 9>{});
            }
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) issue_layer(l + 1, n);
            mbar_wait(mma_bar, mma_par); mma_par ^= 1u;
            tc_fence_after();
            if (tid == 0) {
                prefetch_w(l + 1, n);
                if (l == 2 && n + 1 < my_tiles) load_tile(n + 1);    // the tile copy is dead: overlap the next load
            }
            epilogue(sEpi + l * C, true, l == 2, s, t_start);
            tc_fence_before();
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

namespace {

template <int C, int NB, typename HT>
cudaError_t launch_chain_t(const ChainArgs& a, const CUtensorMap& tmY, const CUtensorMap* tmW, int sm_count,
                           cudaStream_t st) {
    using Cfg = ChainCfg<C, NB>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k_chain<C, NB, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const int tiles = a.S * ((a.T + Cfg::kROut - 1) / Cfg::kROut);
    if (tiles == 0) return cudaSuccess;
    const int grid = tiles < sm_count ? tiles : sm_count;
    k_chain<C, NB, HT><<<grid, kThreads, Cfg::kSmem, st>>>(tmY, tmW[0], tmW[1], tmW[2], tmW[3], a, tiles);
    return cudaGetLastError();
}

}  // namespace

bool chain_supported(int C) { return C == 64 || C == 128; }
int chain_tile_rows(int C) { return C == 64 ? 1024 : 512; }

// Spans of the in-place prologue (see the header comment): for dilation d the rows of a tile split into d classes
// r = r0 + k d.  Class starts are multiples of 8 (so that the swizzle phase of step k is static) no larger than the
// first row whose result is needed at that layer; negative starts skip their first few steps.
void chain_build_spans(int C, ChainSpan (*spans)[kChainWarps][2]) {
    const int rows = chain_tile_rows(C), ch = C / 64;
    static const int dil[3] = {1, 3, 9};
    for (int l = 0; l < 3; ++l) {
        const int d = dil[l];
        struct Cls { int kc, r0, noct; };
        std::vector<Cls> cls;
        const int top = (d == 1) ? 0 : (d == 3 ? 8 : 40);
        for (int kc = 0; kc < ch; ++kc)
            for (int m = 0; m < d; ++m) {
                const int r0 = top - 8 * m;
                const int steps = (rows - r0 + d - 1) / d;
                cls.push_back({kc, r0, (steps + 7) / 8});
            }
        int total = 0;
        for (auto& c : cls) total += c.noct;
        for (int w = 0; w < kChainWarps; ++w) {
            const int lo = static_cast<int>(static_cast<long long>(w) * total / kChainWarps);
            const int hi = static_cast<int>(static_cast<long long>(w + 1) * total / kChainWarps);
            int nsp = 0, base = 0;
            spans[l][w][0] = ChainSpan{0, 0, 0, 0};
            spans[l][w][1] = ChainSpan{0, 0, 0, 0};
            for (auto& c : cls) {
                const int a0 = lo > base ? lo : base, a1 = hi < base + c.noct ? hi : base + c.noct;
                if (a1 > a0) {
                    if (nsp >= 2) { fprintf(stderr, "snacb: chain span table overflow\n"); abort(); }
                    spans[l][w][nsp++] = ChainSpan{static_cast<short>(c.r0 + 8 * d * (a0 - base)),
                                                   static_cast<short>(a1 - a0), static_cast<short>(c.kc), 0};
                }
                base += c.noct;
            }
        }
    }
}

cudaError_t launch_chain(int half_fp16, const ChainArgs& a, const CUtensorMap& tmY, const CUtensorMap* tmW,
                         int sm_count, cudaStream_t st) {
    if (a.C == 64)
        return half_fp16 ? launch_chain_t<64, 8, __half>(a, tmY, tmW, sm_count, st)
                         : launch_chain_t<64, 8, __nv_bfloat16>(a, tmY, tmW, sm_count, st);
    if (a.C == 128)
        return half_fp16 ? launch_chain_t<128, 4, __half>(a, tmY, tmW, sm_count, st)
                         : launch_chain_t<128, 4, __nv_bfloat16>(a, tmY, tmW, sm_count, st);
    return cudaErrorInvalidValue;
}

}  // namespace snacb
