// CUDA-core kernels of the SNAC decode path (sm_100a):
//   * k_unpack        token ids -> three code levels (integer, bit-exact)
//   * k_vq_stem       codebook gather + out_proj + repeat-interleave + level sum + stem depthwise k7
//   * k_gemm_f32      fp32 row-GEMM with taps (1x1 convs, ConvTranspose1d, NoiseBlock) -- fp32 precision path
//   * k_respre_f32    Snake -> depthwise dilated k7 -> Snake (fp32 precision path)
//   * k_tail          final conv 64->1 k7 + tanh + slice + int16 quantise
//   * k_to_f32        debug taps
// Activations are channel-last: [stream][time][channel].
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace snacb {

// ----------------------------------------------------------------------------------------------
// Token -> code unpack.  Reference: vllm_inference/modal_audio_stream.py:165-188 (offsets, clamp),
// tensorrt_tts/inference.py:54-93.  code = clamp(id - [128266] - 4096*(p mod 7), 0, 4095);
// p=0 -> level 0; p=1,4 -> level 1; p=2,3,5,6 -> level 2.
// ----------------------------------------------------------------------------------------------
__global__ void k_unpack(const int32_t* __restrict__ tok, int B, int ntok, int F, int raw_ids,
                         int32_t* __restrict__ c0, int32_t* __restrict__ c1, int32_t* __restrict__ c2) {
    const int per = F * kFrame;
    const long long total = static_cast<long long>(B) * per;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(idx / per), r = static_cast<int>(idx % per);
        const int f = r / kFrame, p = r % kFrame;
        long long v = static_cast<long long>(tok[static_cast<size_t>(b) * ntok + r]) -
                      (raw_ids ? kTokenAudioBase : 0) - 4096LL * p;
        const int c = static_cast<int>(v < 0 ? 0 : (v > 4095 ? 4095 : v));
        const size_t bf = static_cast<size_t>(b) * F + f;
        switch (p) {
            case 0: c0[bf] = c; break;
            case 1: c1[2 * bf] = c; break;
            case 4: c1[2 * bf + 1] = c; break;
            case 2: c2[4 * bf] = c; break;
            case 3: c2[4 * bf + 1] = c; break;
            case 5: c2[4 * bf + 2] = c; break;
            default: c2[4 * bf + 3] = c; break;
        }
    }
}

void launch_unpack(const int32_t* tok, int B, int ntok, int F, int raw_ids, int32_t* c0, int32_t* c1, int32_t* c2,
                   cudaStream_t st) {
    const long long total = static_cast<long long>(B) * F * kFrame;
    if (total == 0) return;
    const int threads = 256;
    const int blocks = static_cast<int>((total + threads - 1) / threads > 148 * 8 ? 148 * 8 : (total + threads - 1) / threads);
    k_unpack<<<blocks, threads, 0, st>>>(tok, B, ntok, F, raw_ids, c0, c1, c2);
}

// ----------------------------------------------------------------------------------------------
// VQ decode + stem depthwise conv.  Reference: snac quantizer.from_codes (oracle/snac_ref.py
// ResidualVectorQuantize.from_codes) followed by decoder.model.0 (depthwise k7, pad 3).
// One CTA = 16 latent steps of one stream; z_q never leaves the SM.
// ----------------------------------------------------------------------------------------------
constexpr int kVqTile = 16;

template <typename OutT>
__global__ void __launch_bounds__(256)
k_vq_stem(const int32_t* __restrict__ c0, const int32_t* __restrict__ c1, const int32_t* __restrict__ c2,
          const int32_t* __restrict__ tok, int tok_stride, int raw_ids, int F, int t_lo, VqStemWeights w,
          OutT* __restrict__ out, const StreamMap map) {
    const int T0 = 4 * F;
    // streaming session: the stream's tokens / output live in its slot, its rows start off[s] frames further on
    const int s = sm_slot(map, blockIdx.y), t0 = t_lo + sm_off(map, blockIdx.y, 4) + blockIdx.x * kVqTile;
    __shared__ float emb[3][kVqTile + 6][kCodeDim];
    __shared__ int ok[kVqTile + 6];
    for (int idx = threadIdx.x; idx < 3 * (kVqTile + 6); idx += blockDim.x) {
        const int lv = idx / (kVqTile + 6), u = idx % (kVqTile + 6);
        const int t = t0 - 3 + u;
        const bool valid = (t >= 0 && t < T0);
        if (lv == 0) ok[u] = valid;
        float e[kCodeDim];
#pragma unroll
        for (int j = 0; j < kCodeDim; ++j) e[j] = 0.f;
        if (valid) {
            int code;
            if (tok != nullptr) {
                // the token -> code unpack of k_unpack, fused into the gather: level 0 sits at position 0 of its frame,
                // level 1 at positions 1 and 4, level 2 at 2, 3, 5, 6 (modal_audio_stream.py:165-188)
                const int j = lv == 0 ? (t >> 2) : (lv == 1 ? (t >> 1) : t);
                const int f = lv == 0 ? j : (lv == 1 ? (j >> 1) : (j >> 2));
                const int p = lv == 0 ? 0 : (lv == 1 ? ((j & 1) ? 4 : 1) : ((j & 3) < 2 ? 2 + (j & 3) : 3 + (j & 3)));
                const long long v = static_cast<long long>(tok[static_cast<size_t>(s) * tok_stride + f * kFrame + p]) -
                                    (raw_ids ? kTokenAudioBase : 0) - 4096LL * p;
                code = static_cast<int>(v < 0 ? 0 : (v > 4095 ? 4095 : v));
            } else if (lv == 0) code = c0[static_cast<size_t>(s) * F + (t >> 2)];
            else if (lv == 1) code = c1[static_cast<size_t>(s) * 2 * F + (t >> 1)];
            else code = c2[static_cast<size_t>(s) * 4 * F + t];
            const float* cb = w.codebook[lv] + static_cast<size_t>(code) * kCodeDim;
#pragma unroll
            for (int j = 0; j < kCodeDim; ++j) e[j] = cb[j];
        }
#pragma unroll
        for (int j = 0; j < kCodeDim; ++j) emb[lv][u][j] = e[j];
    }
    __syncthreads();
    {
        const int c = blockIdx.z * blockDim.x + threadIdx.x;      // grid.z = 768 / 256 channel groups
        float wv[3][kCodeDim];
#pragma unroll
        for (int lv = 0; lv < 3; ++lv) {
#pragma unroll
            for (int j = 0; j < kCodeDim; ++j) wv[lv][j] = w.out_w[lv][c * kCodeDim + j];
        }
        float z[kVqTile + 6];
#pragma unroll
        for (int u = 0; u < kVqTile + 6; ++u) {
            float acc = 0.f;
#pragma unroll
            for (int lv = 0; lv < 3; ++lv) {
                float a = w.out_b[lv][c];
#pragma unroll
                for (int j = 0; j < kCodeDim; ++j) a = fmaf(wv[lv][j], emb[lv][u][j], a);
                acc += a;
            }
            z[u] = ok[u] ? acc : 0.f;
        }
        float dw[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) dw[j] = w.dw_w[j * kLatent + c];
        const float db = w.dw_b[c];
#pragma unroll
        for (int r = 0; r < kVqTile; ++r) {
            const int t = t0 + r;
            if (t < T0) {
                float acc = db;
#pragma unroll
                for (int j = 0; j < 7; ++j) acc = fmaf(dw[j], z[r + j], acc);
                from_f32(out[(static_cast<size_t>(s) * T0 + t) * kLatent + c], acc);
            }
        }
    }
}

template <typename OutT>
void launch_vq_stem(const int32_t* c0, const int32_t* c1, const int32_t* c2, const int32_t* tok, int tok_stride, int raw_ids,
                    int S, int F, int t_lo, int t_hi, const VqStemWeights& w, OutT* out, cudaStream_t st,
                    const StreamMap& map) {
    // latent steps [t_lo, t_hi) of every stream, in 16-step tiles from t_lo rounded down to a tile (the stateless decode) or
    // from the stream's own first step (session: t_lo + off[s] frames); steps past t_hi in the last tile are computed too
    if (t_hi <= t_lo) return;
    const int lo = map.off ? t_lo : (t_lo / kVqTile) * kVqTile;
    dim3 grid((t_hi - lo + kVqTile - 1) / kVqTile, S, kLatent / 256);
    k_vq_stem<OutT><<<grid, 256, 0, st>>>(c0, c1, c2, tok, tok_stride, raw_ids, F, lo, w, out, map);
}
template void launch_vq_stem<float>(const int32_t*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int, int, int,
                                    const VqStemWeights&, float*, cudaStream_t, const StreamMap&);
template void launch_vq_stem<__nv_bfloat16>(const int32_t*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int, int, int,
                                            const VqStemWeights&, __nv_bfloat16*, cudaStream_t, const StreamMap&);
template void launch_vq_stem<__half>(const int32_t*, const int32_t*, const int32_t*, const int32_t*, int, int, int, int, int, int,
                                     const VqStemWeights&, __half*, cudaStream_t, const StreamMap&);

// ----------------------------------------------------------------------------------------------
// fp32 row-GEMM with taps (CUDA cores).  64x64 tile, 16-deep K steps, 4x4 register micro-tile.
// ----------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256)
k_gemm_f32(GemmArgs a, const float* __restrict__ A, const float* __restrict__ W) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Ws[BK][BN];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * BN;
    const long long m0 = static_cast<long long>(blockIdx.y) * BM;
    const long long M = static_cast<long long>(a.S) * a.Tin;
    const int p = n0 / a.Cout;                         // output phase of this column tile
    const int base_shift = (a.up > 1 && p >= a.up / 2) ? 1 : 0;
    const int ldw = a.ntaps * a.K;

    // loader mapping: one float4 of A and one of W per thread per K step
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const long long ar = m0 + lrow;
    const int as_ = static_cast<int>(ar / a.Tin), am = static_cast<int>(ar % a.Tin);
    const float* wrow = W + static_cast<size_t>(n0 + lrow) * ldw;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < a.ntaps; ++tap) {
        const int sm = am + base_shift - tap;
        const bool avalid = (ar < M) && sm >= 0 && sm < a.Tin;
        const float* arow = A + (static_cast<size_t>(as_) * a.Tin + (avalid ? sm : 0)) * a.K;
        for (int k0 = 0; k0 < a.K; k0 += BK) {
            float4 av = avalid ? *reinterpret_cast<const float4*>(arow + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 wv = *reinterpret_cast<const float4*>(wrow + tap * a.K + k0 + lk);
            __syncthreads();
            As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
            Ws[lk + 0][lrow] = wv.x; Ws[lk + 1][lrow] = wv.y; Ws[lk + 2][lrow] = wv.z; Ws[lk + 3][lrow] = wv.w;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
                const float av4[4] = {a4.x, a4.y, a4.z, a4.w};
                const float bv4[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av4[i], bv4[j], acc[i][j]);
            }
        }
    }

    const int o0 = n0 - p * a.Cout + tx * 4;
    float* out = static_cast<float*>(a.out);
    const float* resid = static_cast<const float*>(a.resid);
    unsigned long long key = 0;
    if (EPI == EPI_NOISE && a.noise == nullptr)
        key = splitmix64(a.seed * 0x100000001B3ull + static_cast<unsigned long long>(100 + a.noise_stage));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = m0 + ty * 4 + i;
        if (r >= M) continue;
        const int s = static_cast<int>(r / a.Tin), m = static_cast<int>(r % a.Tin);
        const size_t orow = (static_cast<size_t>(s) * a.Tin + m) * a.up + p;
        float v[4];
        float nz = 0.f;
        if (EPI == EPI_NOISE)
            nz = a.noise ? a.noise[static_cast<size_t>(s) * a.Tin + m]
                         : counter_normal(key, noise_counter(a.stream_keys ? a.stream_keys[s] : a.stream_offset + s, m + a.t0));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = o0 + j;
            float t = acc[i][j];
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_SNAKE) t += a.bias[o];
            if (EPI == EPI_NOISE) t = resid[orow * a.Cout + o] + nz * t;
            if (EPI == EPI_RES || EPI == EPI_RES_SNAKE) t = resid[orow * a.Cout + o] + (t + a.bias[o]);
            if (EPI == EPI_BIAS_SNAKE || EPI == EPI_RES_SNAKE) t = snake_f<false>(t, a.alpha[o], a.inv_alpha[o]);
            v[j] = t;
        }
        *reinterpret_cast<float4*>(out + orow * a.Cout + o0) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

void launch_gemm_f32(int epi, const GemmArgs& a, const float* A, const float* W, cudaStream_t st) {
    const long long M = static_cast<long long>(a.S) * a.Tin;
    dim3 grid(a.N / 64, static_cast<unsigned>((M + 63) / 64));
    switch (epi) {
        case EPI_BIAS: k_gemm_f32<EPI_BIAS><<<grid, 256, 0, st>>>(a, A, W); break;
        case EPI_BIAS_SNAKE: k_gemm_f32<EPI_BIAS_SNAKE><<<grid, 256, 0, st>>>(a, A, W); break;
        case EPI_NOISE: k_gemm_f32<EPI_NOISE><<<grid, 256, 0, st>>>(a, A, W); break;
        case EPI_RES: k_gemm_f32<EPI_RES><<<grid, 256, 0, st>>>(a, A, W); break;
        default: k_gemm_f32<EPI_RES_SNAKE><<<grid, 256, 0, st>>>(a, A, W); break;
    }
}

// ----------------------------------------------------------------------------------------------
// ResidualUnit front half, fp32 path:  P = snake2( dw_b + sum_j dw_w[j] * snake1(x[t + (j-3) d]) ).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_respre_f32(ResUnitArgs a, float* __restrict__ P) {
    const int c4n = a.C / 4;
    const long long total = static_cast<long long>(a.S) * a.T * c4n;
    const float* x = static_cast<const float*>(a.x);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c4n) * 4;
        const long long row = idx / c4n;
        const int s = static_cast<int>(row / a.T), t = static_cast<int>(row % a.T);
        const float4 al = *reinterpret_cast<const float4*>(a.alpha1 + c);
        const float4 ia = *reinterpret_cast<const float4*>(a.inv_alpha1 + c);
        float4 acc = *reinterpret_cast<const float4*>(a.dw_b + c);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int tt = t + (j - 3) * a.dil;
            if (tt < 0 || tt >= a.T) continue;
            const float4 xv = *reinterpret_cast<const float4*>(x + (static_cast<size_t>(s) * a.T + tt) * a.C + c);
            const float4 wv = *reinterpret_cast<const float4*>(a.dw_w + j * a.C + c);
            acc.x = fmaf(wv.x, snake_f<false>(xv.x, al.x, ia.x), acc.x);
            acc.y = fmaf(wv.y, snake_f<false>(xv.y, al.y, ia.y), acc.y);
            acc.z = fmaf(wv.z, snake_f<false>(xv.z, al.z, ia.z), acc.z);
            acc.w = fmaf(wv.w, snake_f<false>(xv.w, al.w, ia.w), acc.w);
        }
        const float4 a2 = *reinterpret_cast<const float4*>(a.alpha2 + c);
        const float4 i2 = *reinterpret_cast<const float4*>(a.inv_alpha2 + c);
        float4 o;
        o.x = snake_f<false>(acc.x, a2.x, i2.x);
        o.y = snake_f<false>(acc.y, a2.y, i2.y);
        o.z = snake_f<false>(acc.z, a2.z, i2.z);
        o.w = snake_f<false>(acc.w, a2.w, i2.w);
        *reinterpret_cast<float4*>(P + row * a.C + c) = o;
    }
}

void launch_respre_f32(const ResUnitArgs& a, float* P, cudaStream_t st) {
    const long long total = static_cast<long long>(a.S) * a.T * (a.C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    k_respre_f32<<<static_cast<unsigned>(blocks), 256, 0, st>>>(a, P);
}

// ----------------------------------------------------------------------------------------------
// Streaming session helpers (snacb_session_step_multi): new tokens of n streams go to each stream's own position of its
// slot's token row; the launch's noise keys are gathered; a full window slides by moving its last frames to the front.
// ----------------------------------------------------------------------------------------------
__global__ void k_session_scatter(const int32_t* __restrict__ new_tok, int tok_stride, int n_int, const int* __restrict__ slot,
                                  const int* __restrict__ pos, int32_t* __restrict__ tok_buf, int slot_ints,
                                  const int32_t* __restrict__ keys_in, int32_t* __restrict__ keys_out, int n) {
    const int s = blockIdx.y;
    if (s >= n) return;
    int32_t* dst = tok_buf + static_cast<size_t>(slot[s]) * slot_ints + static_cast<size_t>(7) * pos[s];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_int; j += gridDim.x * blockDim.x)
        dst[j] = new_tok[static_cast<size_t>(s) * tok_stride + j];
    if (blockIdx.x == 0 && threadIdx.x == 0) keys_out[s] = keys_in ? keys_in[s] : slot[s];
}
void launch_session_scatter(const int32_t* new_tok, int tok_stride, int n_int, const int* slot, const int* pos, int32_t* tok_buf,
                            int slot_ints, const int32_t* keys_in, int32_t* keys_out, int n, cudaStream_t st) {
    if (n <= 0) return;
    dim3 grid(n_int > 0 ? (n_int + 127) / 128 : 1, n);
    k_session_scatter<<<grid, 128, 0, st>>>(new_tok, tok_stride, n_int, slot, pos, tok_buf, slot_ints, keys_in, keys_out, n);
}

// rows [slide[i], slide[i] + keep) frames of slot list[i] -> rows [0, keep); source and destination never overlap (slide >= keep)
__global__ void k_session_slide(uint8_t* __restrict__ buf, size_t slot_bytes, size_t frame_bytes, int keep, const int* __restrict__ list,
                                const int* __restrict__ slide, int n) {
    const int i = blockIdx.y;
    if (i >= n) return;
    uint4* dst = reinterpret_cast<uint4*>(buf + static_cast<size_t>(list[i]) * slot_bytes);
    const uint4* src = reinterpret_cast<const uint4*>(buf + static_cast<size_t>(list[i]) * slot_bytes + static_cast<size_t>(slide[i]) * frame_bytes);
    const size_t n16 = static_cast<size_t>(keep) * frame_bytes / 16;
    for (size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; j < n16; j += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[j] = src[j];
}
void launch_session_slide(void* buf, size_t slot_bytes, size_t frame_bytes, int keep, const int* list, const int* slide, int n,
                          cudaStream_t st) {
    if (n <= 0 || !buf) return;
    const size_t n16 = static_cast<size_t>(keep) * frame_bytes / 16;
    unsigned gx = static_cast<unsigned>((n16 + 255) / 256);
    if (gx > 256) gx = 256;
    if (gx < 1) gx = 1;
    dim3 grid(gx, n);
    k_session_slide<<<grid, 256, 0, st>>>(static_cast<uint8_t*>(buf), slot_bytes, frame_bytes, keep, list, slide, n);
}

// ----------------------------------------------------------------------------------------------
// bf16x3 path (precision = bf16 at the 40 dB bar): activations are STORED in fp16, the tensor
// cores multiply bf16.  An fp16 value splits EXACTLY into two bf16 values (11 significand bits = 8 + 3), a weight into
// hi + lo to 16 bits, so  A W^T = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T  with every product exact in fp32: the GEMM sees
// A'' = [A_hi | A_lo | A_hi] (this kernel) against W'' = [W_hi | W_hi | W_lo] (packed at load) and K'' = 3K.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_split3(const __half* __restrict__ x, int S, int T, int K, int r_lo, int r_n, __nv_bfloat16* __restrict__ out,
         const StreamMap map, int rpf) {
    const int k8n = K / 8;
    const long long total = static_cast<long long>(S) * r_n * k8n;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(idx % k8n) * 8;
        const long long rr = idx / k8n;
        const int s = static_cast<int>(rr / r_n), t = r_lo + sm_off(map, s, rpf) + static_cast<int>(rr % r_n);
        if (t >= T) continue;
        const size_t row = static_cast<size_t>(sm_slot(map, s)) * T + t;
        const uint4 raw = *reinterpret_cast<const uint4*>(x + row * K + k);
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
        uint4 hi, lo;
        uint32_t* ph = reinterpret_cast<uint32_t*>(&hi);
        uint32_t* pl = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 v = __half22float2(h2[j]);
            const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
            const float2 hf = __bfloat1622float2(h);
            const __nv_bfloat162 l = __floats2bfloat162_rn(v.x - hf.x, v.y - hf.y);
            ph[j] = *reinterpret_cast<const uint32_t*>(&h);
            pl[j] = *reinterpret_cast<const uint32_t*>(&l);
        }
        // [hi | lo]; the GEMM reads the third block of A'' = [hi | lo | hi] by wrapping its K chunk index (GemmArgs::a_wrap)
        __nv_bfloat16* o = out + row * (2 * static_cast<size_t>(K)) + k;
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + K) = lo;
    }
}

void launch_split3(const __half* x, int S, int T, int K, int r_lo, int r_n, __nv_bfloat16* out, cudaStream_t st,
                   const StreamMap& map, int rpf) {
    const long long total = static_cast<long long>(S) * r_n * (K / 8);
    if (total <= 0) return;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    k_split3<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, S, T, K, r_lo, r_n, out, map, rpf);
}

// ResidualUnit front half of the bf16x3 path: a = snake2(dw_b + sum_j dw_w[j] * snake1(x[t + (j - 3) d])) in fp32 from the
// fp16 residual stream, written as bf16 [rows][C] (the 1x1 conv runs against [W_hi | W_lo], reading a twice through
// GemmArgs::a_wrap).  A thread owns 4 channels and walks kRespreRun steps of one dilation class (rows r, r + d, ...): the
// 7-tap window of snake1 values slides in registers, so each Snake is evaluated once per element.
constexpr int kRespreRun = 64;
__global__ void __launch_bounds__(256)
k_respre16(ResUnitArgs a, int r_lo_, int r_n, __nv_bfloat16* __restrict__ out) {
    const int c4n = a.C / 4, d = a.dil;
    const int per_class = (r_n + d - 1) / d;                       // steps of the longest class inside [r_lo, r_lo + r_n)
    const int segs = (per_class + kRespreRun - 1) / kRespreRun;
    const long long total = static_cast<long long>(a.S) * d * segs * c4n;
    const __half* x = static_cast<const __half*>(a.x);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c4n) * 4;
        long long rest = idx / c4n;
        const int seg = static_cast<int>(rest % segs); rest /= segs;
        const int cls = static_cast<int>(rest % d);
        const int s_ = static_cast<int>(rest / d), s = sm_slot(a.map, s_);   // s: buffer slot (addressing)
        const int r_lo = r_lo_ + sm_off(a.map, s_, a.rpf);
        const float4 al = *reinterpret_cast<const float4*>(a.alpha1 + c);
        const float4 ia = *reinterpret_cast<const float4*>(a.inv_alpha1 + c);
        const float4 a2 = *reinterpret_cast<const float4*>(a.alpha2 + c);
        const float4 i2 = *reinterpret_cast<const float4*>(a.inv_alpha2 + c);
        const float4 db = *reinterpret_cast<const float4*>(a.dw_b + c);
        float4 w[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) w[j] = *reinterpret_cast<const float4*>(a.dw_w + j * a.C + c);
        const __half* xs = x + static_cast<size_t>(s) * a.T * a.C + c;
        auto s1 = [&](int t) -> float4 {                          // snake1(x[t]), zero outside the stream (conv padding)
            if (t < 0 || t >= a.T) return make_float4(0.f, 0.f, 0.f, 0.f);
            const uint2 raw = *reinterpret_cast<const uint2*>(xs + static_cast<size_t>(t) * a.C);
            const float2 x01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            const float2 x23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            return make_float4(snake_f<true>(x01.x, al.x, ia.x), snake_f<true>(x01.y, al.y, ia.y),
                               snake_f<true>(x23.x, al.z, ia.z), snake_f<true>(x23.y, al.w, ia.w));
        };
        const int t_first = r_lo + cls + seg * kRespreRun * d;     // first output row of this run
        float4 win[7];
#pragma unroll
        for (int j = 0; j < 6; ++j) win[j + 1] = s1(t_first + (j - 3) * d);
        __nv_bfloat16* os = out + static_cast<size_t>(s) * a.T * a.C + c;
        for (int k = 0; k < kRespreRun; ++k) {
            const int t = t_first + k * d;
            if (t >= r_lo + r_n || t >= a.T) break;
#pragma unroll
            for (int j = 0; j < 6; ++j) win[j] = win[j + 1];
            win[6] = s1(t + 3 * d);
            float4 acc = db;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                acc.x = fmaf(w[j].x, win[j].x, acc.x); acc.y = fmaf(w[j].y, win[j].y, acc.y);
                acc.z = fmaf(w[j].z, win[j].z, acc.z); acc.w = fmaf(w[j].w, win[j].w, acc.w);
            }
            const __nv_bfloat162 o01 = __floats2bfloat162_rn(snake_f<true>(acc.x, a2.x, i2.x), snake_f<true>(acc.y, a2.y, i2.y));
            const __nv_bfloat162 o23 = __floats2bfloat162_rn(snake_f<true>(acc.z, a2.z, i2.z), snake_f<true>(acc.w, a2.w, i2.w));
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t*>(&o01);
            o.y = *reinterpret_cast<const uint32_t*>(&o23);
            *reinterpret_cast<uint2*>(os + static_cast<size_t>(t) * a.C) = o;
        }
    }
}

void launch_respre16(const ResUnitArgs& a, int r_lo, int r_n, __nv_bfloat16* out, cudaStream_t st) {
    if (r_n <= 0) return;
    const int per_class = (r_n + a.dil - 1) / a.dil;
    const long long total = static_cast<long long>(a.S) * a.dil * ((per_class + kRespreRun - 1) / kRespreRun) * (a.C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    k_respre16<<<static_cast<unsigned>(blocks), 256, 0, st>>>(a, r_lo, r_n, out);
}

// ----------------------------------------------------------------------------------------------
// Tail: conv 64->1 k7 pad 3 over the (already Snake'd) block-3 output, tanh, optional slice
// [2048:4096] (vllm_inference/modal_audio_stream.py:94-95,195-198), int16 quantise (:201).
// One warp = 32 consecutive samples; lane = channel pair; 7-row sliding window in registers.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 load_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load_pair(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }
__device__ __forceinline__ float2 load_pair(const __nv_bfloat16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

// One warp = 32 consecutive output samples of one stream; lane = channel pair while the 7-tap window slides down
// the rows (each input row is loaded once, 128 B per warp), every lane keeping its 2-channel partial sum of all 32
// samples in registers; one recursive-halving exchange (31 shuffles) then leaves sample `lane` summed in lane `lane`.
// rows(i), i < 38: channel pair of this lane in input row base - 3 + i (zero outside the stream).  Returns sample
// base + lane in lane `lane` (before bias / tanh).  The arithmetic (tap order, lane tree) is the same wherever the
// rows come from, so every tail kernel below yields the same bits.
template <typename RowFn>
__device__ __forceinline__ float tail_chunk(const float2 (&wj)[7], int lane, RowFn rows) {
    float2 r[38];
#pragma unroll
    for (int i = 0; i < 38; ++i) r[i] = rows(i);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 7; ++j) acc = ffma2(wj[j], r[i + j], acc);
        v[i] = acc.x + acc.y;
    }
    // transpose-reduce: after the step with offset o, a lane keeps the half of its values whose sample index has
    // bit o equal to the lane's bit o, summed with the partner's copy
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = up ? v[i] : v[i + o];
            const float keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// MAPPED: a streaming-session launch (per-stream slot / first sample); the stateless decode compiles without the map
template <typename InT, bool MAPPED>
__global__ void __launch_bounds__(256)
k_tail(const InT* __restrict__ a, int T, int t_begin_, int n_out, const float* __restrict__ w /*[7][64]*/, float bias,
       int16_t* __restrict__ pcm, float* __restrict__ wave, const StreamMap map) {
    const int s = blockIdx.y;
    const int t_begin = MAPPED ? t_begin_ + sm_off(map, s, 2048) : t_begin_;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = t_begin + blockIdx.x * 256 + warp * 32;
    if (base >= t_begin + n_out) return;
    float2 wj[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) wj[j] = make_float2(w[j * 64 + 2 * lane], w[j * 64 + 2 * lane + 1]);
    const InT* src = a + static_cast<size_t>(MAPPED ? sm_slot(map, s) : s) * T * 64 + 2 * lane;
    const float v = tail_chunk(wj, lane, [&](int i) -> float2 {     // one 4-byte (16-bit types) or 8-byte load per lane and row
        const int t = base - 3 + i;
        if (t < 0 || t >= T) return make_float2(0.f, 0.f);
        return load_pair(src + static_cast<size_t>(t) * 64);
    });
    const int t = base + lane;
    if (t < t_begin + n_out && t < T) {
        const float r = tanhf(v + bias);
        const size_t o = static_cast<size_t>(s) * n_out + (t - t_begin);
        pcm[o] = pcm16(r);
        if (wave) wave[o] = r;
    }
}

// The same for large batches of 16-bit activations, shaped for HBM bandwidth: persistent CTAs, the 262 input rows
// (33.5 KB, contiguous in the channel-last layout) of a 256-sample tile arrive by ONE bulk async copy (cp.async.bulk,
// mbarrier completion) into a 3-deep shared-memory ring, so ~100 KB per CTA are in flight while the 8 warps compute
// from shared memory (lane = channel pair: 128 contiguous bytes per row, conflict-free).
constexpr int kTailTile = 256;
constexpr int kTailRows = kTailTile + 6;
constexpr int kTailStages = 3;
constexpr int kTailStageBytes = kTailRows * 128;
constexpr int kTailSmem = kTailStages * kTailStageBytes + 64 + 128;

template <typename InT, bool MAPPED>
__global__ void __launch_bounds__(256, 2)
k_tail_bulk(const InT* __restrict__ a, int T, int t_begin_, int n_out, const float* __restrict__ w, float bias,
            int16_t* __restrict__ pcm, float* __restrict__ wave, int tiles_per_stream, int num_tiles, const StreamMap map) {
    static_assert(sizeof(InT) == 2, "16-bit activations");
    extern __shared__ __align__(128) uint8_t tail_smem[];
    uint8_t* ring = tail_smem + ((128u - (ptx::smem_u32(tail_smem) & 127u)) & 127u);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kTailStages * kTailStageBytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTailStages; ++i) ptx::mbar_init(&full[i], 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    float2 wj[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) wj[j] = make_float2(w[j * 64 + 2 * lane], w[j * 64 + 2 * lane + 1]);

    auto issue = [&](int tile, int stage) {               // thread 0: rows [base - 3, base + 259) clipped to the stream
        const int s = tile / tiles_per_stream;
        const int base = t_begin_ + (MAPPED ? sm_off(map, s, 2048) : 0) + (tile % tiles_per_stream) * kTailTile;
        const int lo = max(base - 3, 0), hi = min(base + kTailTile + 3, T);
        const uint32_t bytes = static_cast<uint32_t>(hi - lo) * 128u;
        const InT* src = a + (static_cast<size_t>(MAPPED ? sm_slot(map, s) : s) * T + lo) * 64;
        uint8_t* dst = ring + stage * kTailStageBytes + (lo - (base - 3)) * 128;
        ptx::mbar_expect_tx(&full[stage], bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(ptx::smem_u32(dst)), "l"(src), "r"(bytes), "r"(ptx::smem_u32(&full[stage])) : "memory");
    };
    if (threadIdx.x == 0)
        for (int k = 0; k < kTailStages; ++k) {
            const int tile = blockIdx.x + k * gridDim.x;
            if (tile < num_tiles) issue(tile, k);
        }
    int k = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
        const int stage = k % kTailStages;
        const int s = tile / tiles_per_stream;
        const int t_begin = t_begin_ + (MAPPED ? sm_off(map, s, 2048) : 0);
        const int tile_base = t_begin + (tile % tiles_per_stream) * kTailTile;
        ptx::mbar_wait(&full[stage], (k / kTailStages) & 1);
        const int base = tile_base + warp * 32;
        if (base < t_begin + n_out) {
            const uint8_t* rows = ring + stage * kTailStageBytes + (warp * 32) * 128 + lane * 4;
            const float v = tail_chunk(wj, lane, [&](int i) -> float2 {
                const int t = base - 3 + i;
                if (t < 0 || t >= T) return make_float2(0.f, 0.f);
                return load_pair(reinterpret_cast<const InT*>(rows + i * 128));
            });
            const int t = base + lane;
            if (t < t_begin + n_out && t < T) {
                const float r = tanhf(v + bias);
                const size_t o = static_cast<size_t>(s) * n_out + (t - t_begin);
                pcm[o] = pcm16(r);
                if (wave) wave[o] = r;
            }
        }
        __syncthreads();                                   // every warp is done with the stage: refill it
        if (threadIdx.x == 0) {
            const int nxt = tile + kTailStages * gridDim.x;
            if (nxt < num_tiles) issue(nxt, stage);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// k_tail_tc: the same conv 64 -> 1 k7 + tanh + int16 for 16-bit activations, shaped so that nothing but HBM is left:
// the channel contraction runs on the tensor core.  P[t][k] = sum_c a[t][c] * w[k][c] is a GEMM with M = time rows,
// N = the 7 taps, K = 64 channels (tcgen05.mma M = 128, N = 16: columns 0..6 carry the 16-bit "hi" part of the fp32
// weights, columns 8..14 the "lo" part -- w = hi + lo to 2^-22 -- accumulated in fp32 in TMEM), and the output is the
// diagonal sum  y[t] = b + sum_k P[t + k - 3][k].  Per 256-sample tile: the 264 input rows [base - 3, base + 261) arrive
// 128B-swizzled by three TMA loads (boxes of 128, 128 and 8 rows; rows outside the stream are the zero fill = the conv's
// padding) into a 3-stage ring, one thread issues 12 MMAs (three 128-row blocks at stage rows 0, 128 and 136), the eight
// warps drain the 16 accumulator columns of their rows (tcgen05.ld), fold hi + lo and park P tap-major in shared memory,
// and every thread sums one output's diagonal.  Each P[t][k] depends on row t alone and the diagonal is summed in a fixed
// order, so the result does not depend on where a tile starts: a session step, a ranged decode and the batch decode give
// the same bits (the property the FFMA kernels above had by sharing tail_chunk).  Before: 38 LDS + 224 FFMA2 + a
// 31-step transposing lane reduction per 32 samples and warp -- issue-bound at 3.4 TB/s of input.
// ----------------------------------------------------------------------------------------------
constexpr int kTtTile = 256;
constexpr int kTtRows = 264;                       // 3 + 256 + 3, rounded up to the 8-row box of the last load
constexpr int kTtStageBytes = kTtRows * 128;       // 33 KB, a multiple of 1024
constexpr int kTtStages = 3;
constexpr int kTtOffW = kTtStages * kTtStageBytes; // [16 n][64 k] weights, 128B-swizzled K-major
constexpr int kTtOffP = kTtOffW + 2048;            // P tap-major: [7][kTtRows] fp32
constexpr int kTtOffBar = kTtOffP + 7 * kTtRows * 4;
constexpr int kTtSmem = kTtOffBar + 64 + 1024;
constexpr int kTtTmemCols = 128;                   // two accumulator stages of three blocks x 16 columns (64 columns each)

template <typename InT, bool MAPPED>
__global__ void __launch_bounds__(256, 2)
k_tail_tc(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm8, int T, int t_begin_, int n_out,
          const float* __restrict__ w, float bias, int16_t* __restrict__ pcm, float* __restrict__ wave, int tiles_per_stream,
          int num_tiles, const StreamMap map) {
    using namespace ptx;
    static_assert(sizeof(InT) == 2, "16-bit activations");
    extern __shared__ __align__(1024) uint8_t tt_smem[];
    uint8_t* smem = tt_smem + ((1024u - (smem_u32(tt_smem) & 1023u)) & 1023u);
    uint8_t* sWt = smem + kTtOffW;
    float* sP = reinterpret_cast<float*>(smem + kTtOffP);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kTtOffBar);      // [kTtStages]
    uint64_t* mma_bar = full + kTtStages;                                // [2] by tile parity
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        prefetch_tmap(&tm128); prefetch_tmap(&tm8);
        for (int i = 0; i < kTtStages; ++i) mbar_init(&full[i], 1);
        mbar_init(&mma_bar[0], 1); mbar_init(&mma_bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, kTtTmemCols); tmem_relinquish(); }
    for (int i = tid; i < 16 * 64; i += 256) {       // weights: row n = tap (hi) / 8 + tap (lo), column k = channel
        const int n = i >> 6, k = i & 63, tap = n & 7;
        float v = 0.f;
        if (tap < 7) {
            const float wf = w[tap * 64 + k];
            const InT hi = static_cast<InT>(wf);
            v = (n < 8) ? static_cast<float>(hi) : wf - static_cast<float>(hi);
        }
        *reinterpret_cast<InT*>(sWt + sw128_offset(n, k)) = static_cast<InT>(v);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto issue = [&](int tile, int stage) {          // thread 0
        const int s = tile / tiles_per_stream;
        const int base = t_begin_ + (MAPPED ? sm_off(map, s, 2048) : 0) + (tile % tiles_per_stream) * kTtTile;
        const int sl = MAPPED ? sm_slot(map, s) : s;
        uint8_t* dst = smem + stage * kTtStageBytes;
        mbar_expect_tx(&full[stage], kTtStageBytes);
        tma_load_3d(dst, &tm128, 0, base - 3, sl, &full[stage]);
        tma_load_3d(dst + 128 * 128, &tm128, 0, base + 125, sl, &full[stage]);
        tma_load_3d(dst + 256 * 128, &tm8, 0, base + 253, sl, &full[stage]);
    };
    if (tid == 0)
        for (int k = 0; k < kTtStages; ++k) {
            const int tile = blockIdx.x + k * gridDim.x;
            if (tile < num_tiles) issue(tile, k);
        }
    constexpr uint32_t idesc = umma_idesc_f16(128, 16, std::is_same<InT, __half>::value ? 0u : 1u);
    // the MMAs of tile k (thread 0): into accumulator stage k & 1, completion on mma_bar[k & 1].  They are issued one tile
    // ahead -- tile k + 1's while tile k is drained and summed -- so that their latency is off the critical path
    auto issue_mma = [&](int kq) {
        const int stage = kq % kTtStages;
        mbar_wait(&full[stage], (kq / kTtStages) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + stage * kTtStageBytes), w0 = smem_u32(sWt);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const uint32_t ab = a0 + (b == 0 ? 0 : (b == 1 ? 128 * 128 : 136 * 128));     // block 2 = stage rows 136..263
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                mma_f16_ss(tmem_base + (kq & 1) * 64 + b * 16, umma_desc_sw128(ab + kk * 32), umma_desc_sw128(w0 + kk * 32), idesc,
                           kk > 0 ? 1u : 0u);
        }
        mma_commit(&mma_bar[kq & 1]);
    };
    if (tid == 0 && static_cast<int>(blockIdx.x) < num_tiles) issue_mma(0);
    int k = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
        const int stage = k % kTtStages;
        const int s = tile / tiles_per_stream;
        const int t_begin = t_begin_ + (MAPPED ? sm_off(map, s, 2048) : 0);
        const int tile_base = t_begin + (tile % tiles_per_stream) * kTtTile;
        mbar_wait(&mma_bar[k & 1], (k >> 1) & 1);
        tc_fence_after();
        if (tid == 0) {
            const int nxt = tile + kTtStages * gridDim.x;     // the MMAs have read the stage: refill it
            if (nxt < num_tiles) issue(nxt, stage);
            if (tile + static_cast<int>(gridDim.x) < num_tiles) issue_mma(k + 1);   // stage (k + 1) & 1 was drained in iteration k - 1
        }
        const uint32_t tacc = tmem_base + (k & 1) * 64;
        {
            // warps 0..3: block 0 (stage rows 0..127); warps 4..7: block 1 (128..255); warp 3 also the last 8 rows
            // (256..263 = TMEM lanes 120..127 of block 2)
            const int q = warp & 3, blk = warp >> 2;
            uint32_t r[16];
            tmem_ld16(tacc + (static_cast<uint32_t>(q * 32) << 16) + blk * 16, r);
            tmem_ld_wait();
            const int row = blk * 128 + q * 32 + lane;
#pragma unroll
            for (int j = 0; j < 7; ++j) sP[j * kTtRows + row] = __uint_as_float(r[j]) + __uint_as_float(r[8 + j]);
            if (warp == 3) {
                tmem_ld16(tacc + (static_cast<uint32_t>(96) << 16) + 32, r);
                tmem_ld_wait();
                if (lane >= 24) {
#pragma unroll
                    for (int j = 0; j < 7; ++j) sP[j * kTtRows + 232 + lane] = __uint_as_float(r[j]) + __uint_as_float(r[8 + j]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        {
            const int t = tile_base + tid;             // stage row tid + j holds input row t - 3 + j
            if (t < t_begin + n_out && t < T) {
                float v = sP[tid];
#pragma unroll
                for (int j = 1; j < 7; ++j) v += sP[j * kTtRows + tid + j];
                const float r = tanhf(v + bias);
                const size_t o = static_cast<size_t>(s) * n_out + (t - t_begin);
                pcm[o] = pcm16(r);
                if (wave) wave[o] = r;
            }
        }
        __syncthreads();                               // P and the accumulators are free for the next tile
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTtTmemCols);
}

template <typename InT, bool MAPPED>
static void launch_tail_m(const InT* a, int S, int T, int t_begin, int n_out, const float* w, float bias, int16_t* pcm,
                          float* wave, cudaStream_t st, const StreamMap& map, const CUtensorMap* tm128, const CUtensorMap* tm8) {
    if constexpr (sizeof(InT) == 2) {
        const int tps = (n_out + kTailTile - 1) / kTailTile;
        const long long tiles = static_cast<long long>(S) * tps;
        const char* v1 = getenv("SNACB_TAIL_V1");          // A/B switch (tests): 1 = the per-warp-load FFMA kernel for every
                                                           // size, 2 = the FFMA kernels (bulk ring for large batches)
        if (tm128 != nullptr && tm8 != nullptr && tiles > 0 && tiles < (1LL << 31) && !(v1 && (v1[0] == '1' || v1[0] == '2'))) {
            static PerDeviceOnce once_tc;
            int dev;
            bool ok = true;
            if (once_tc.needed(&dev)) {
                ok = cudaFuncSetAttribute(k_tail_tc<InT, MAPPED>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTtSmem) == cudaSuccess;
                if (ok) once_tc.done(dev);
            }
            if (ok) {
                const int grid = tiles < 2 * 148 ? static_cast<int>(tiles) : 2 * 148;
                k_tail_tc<InT, MAPPED><<<grid, 256, kTtSmem, st>>>(*tm128, *tm8, T, t_begin, n_out, w, bias, pcm, wave, tps,
                                                                    static_cast<int>(tiles), map);
                return;
            }
            (void)cudaGetLastError();
        }
        if (tiles >= 2 * 148 && tiles < (1LL << 31) && !(v1 && v1[0] == '1')) {
            static PerDeviceOnce once;
            int dev;
            bool ok = true;
            if (once.needed(&dev)) {
                ok = cudaFuncSetAttribute(k_tail_bulk<InT, MAPPED>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmem) == cudaSuccess;
                if (ok) once.done(dev);
            }
            if (ok) {
                k_tail_bulk<InT, MAPPED><<<2 * 148, 256, kTailSmem, st>>>(a, T, t_begin, n_out, w, bias, pcm, wave, tps,
                                                                           static_cast<int>(tiles), map);
                return;
            }
            (void)cudaGetLastError();       // the opt-in failed: fall through to the kernel that needs no dynamic shared memory
        }
    }
    dim3 grid((n_out + 255) / 256, S);
    k_tail<InT, MAPPED><<<grid, 256, 0, st>>>(a, T, t_begin, n_out, w, bias, pcm, wave, map);
}
// tm128 / tm8 (16-bit activations; null = the FFMA kernels): the input as [slot][T][64], 128B-swizzled, boxes (64, 128, 1) / (64, 8, 1)
template <typename InT>
void launch_tail(const InT* a, int S, int T, int t_begin, int n_out, const float* w, float bias, int16_t* pcm,
                 float* wave, cudaStream_t st, const StreamMap& map, const CUtensorMap* tm128, const CUtensorMap* tm8) {
    if (map.slot || map.off) launch_tail_m<InT, true>(a, S, T, t_begin, n_out, w, bias, pcm, wave, st, map, tm128, tm8);
    else launch_tail_m<InT, false>(a, S, T, t_begin, n_out, w, bias, pcm, wave, st, map, tm128, tm8);
}
template void launch_tail<float>(const float*, int, int, int, int, const float*, float, int16_t*, float*, cudaStream_t,
                                 const StreamMap&, const CUtensorMap*, const CUtensorMap*);
template void launch_tail<__nv_bfloat16>(const __nv_bfloat16*, int, int, int, int, const float*, float, int16_t*,
                                         float*, cudaStream_t, const StreamMap&, const CUtensorMap*, const CUtensorMap*);
template void launch_tail<__half>(const __half*, int, int, int, int, const float*, float, int16_t*, float*, cudaStream_t,
                                  const StreamMap&, const CUtensorMap*, const CUtensorMap*);

// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_to_f32(const T* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = to_f32(in[i]);
}
template <typename T>
void launch_to_f32(const T* in, float* out, size_t n, cudaStream_t st) {
    if (n == 0) return;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_to_f32<T><<<static_cast<unsigned>(blocks), 256, 0, st>>>(in, out, n);
}
template void launch_to_f32<float>(const float*, float*, size_t, cudaStream_t);
template void launch_to_f32<__nv_bfloat16>(const __nv_bfloat16*, float*, size_t, cudaStream_t);
template void launch_to_f32<__half>(const __half*, float*, size_t, cudaStream_t);

}  // namespace snacb
