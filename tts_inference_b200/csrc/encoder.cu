// SNAC 24 kHz ENCODE path: audio -> Encoder -> residual vector quantisation -> codes
// (-> token ids).  Upstream snac: SNAC.preprocess / encode (snac/snac.py), Encoder / EncoderBlock / ResidualUnit
// (snac/layers.py), VectorQuantize.forward / decode_latents (snac/vq.py); restated on the CPU in oracle/snac_enc_ref.py.
// The reference never calls encode at inference; this completes the codec for dataset tokenisation and round-trip tests.
//
// fp32 CUDA-core arithmetic end to end: the code indices are an argmax over 4096 distances per latent step, so the
// latent has to match the fp32 reference closely for the codes to be the same; 16-bit tensor-core operands would flip
// near-ties.  Activations are channel-last [stream][time][channel] as on the decode path.
//   k_enc_conv0     1 -> 48, k7
//   k_respre_f32    Snake -> depthwise dilated k7 -> Snake          (kernels_simt.cu, shared with the fp32 decode path)
//   k_enc_gemm      1x1 conv (+ residual, + next Snake) and the strided down-sampling conv (k = 2s): the 2s input rows
//                   of an output step are CONTIGUOUS in the channel-last layout, so the strided conv is a GEMM whose A
//                   rows overlap (row pitch s * Cin, K = 2 s Cin), zero padding resolved in the loader
//   k_enc_dw7       final depthwise k7 (768)
//   k_enc_vq        one RVQ level: avg-pool, in_proj 768 -> 8, L2-normalise, nearest of 4096 normalised codes,
//                   z_q = out_proj(codebook[idx]), residual -= repeat_interleave(z_q)
//   k_pack_tokens   codes -> 7 token ids per frame (inverse of k_unpack)
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <new>
#include <string>
#include <vector>

#include "../../include/snacb.h"
#include "common.cuh"
#include "kernels.h"

namespace snacb {
namespace {

constexpr int kEncDim = 48;
constexpr int kEncRates[4] = {2, 4, 8, 8};

// ---------------------------------------------------------------------------------------------- conv0
__global__ void __launch_bounds__(256)
k_enc_conv0(const float* __restrict__ x, int S, int T, int n_valid, int x_stride, const float* __restrict__ w /*[7][48]*/,
            const float* __restrict__ b, float* __restrict__ out) {
    const long long total = static_cast<long long>(S) * T * (kEncDim / 4);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % (kEncDim / 4)) * 4;
        const long long row = idx / (kEncDim / 4);
        const int s = static_cast<int>(row / T), t = static_cast<int>(row % T);
        float4 acc = *reinterpret_cast<const float4*>(b + c);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int tt = t + j - 3;
            if (tt < 0 || tt >= n_valid) continue;            // conv zero padding; samples past n_valid are preprocess()'s right pad
            const float xv = x[static_cast<size_t>(s) * x_stride + tt];
            const float4 wv = *reinterpret_cast<const float4*>(w + j * kEncDim + c);
            acc.x = fmaf(wv.x, xv, acc.x); acc.y = fmaf(wv.y, xv, acc.y);
            acc.z = fmaf(wv.z, xv, acc.z); acc.w = fmaf(wv.w, xv, acc.w);
        }
        *reinterpret_cast<float4*>(out + row * kEncDim + c) = acc;
    }
}

// ---------------------------------------------------------------------------------------------- GEMM
struct EncGemm {
    int S, Tin, Tout, Cin, N, stride, pad, K;   // K = taps * Cin, taps = 1 (1x1) or 2 * stride
    const float* A;        // [S][Tin][Cin]
    const float* W;        // [N][K], k = tap * Cin + ci
    const float* bias;     // [N]
    const float* resid;    // [S][Tout][N] or null
    const float *alpha, *inv_alpha;   // [N] Snake applied to the result, or null
    float* out;            // [S][Tout][N]
};

template <bool RES, bool SNAKE>
__global__ void __launch_bounds__(256)
k_enc_gemm(const EncGemm a) {
    constexpr int BM = 64, BN = 48, BK = 16;
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Ws[BK][BN];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * BN;
    const long long m0 = static_cast<long long>(blockIdx.y) * BM;
    const long long M = static_cast<long long>(a.S) * a.Tout;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const long long ar = m0 + lrow;
    const int as_ = static_cast<int>(ar / a.Tout), at = static_cast<int>(ar % a.Tout);
    const int r0 = at * a.stride - a.pad;                       // input row of tap 0
    const float* abase = a.A + static_cast<size_t>(as_) * a.Tin * a.Cin;
    const float* wrow = a.W + static_cast<size_t>(n0 + (lrow < BN ? lrow : 0)) * a.K;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < a.K; k0 += BK) {
        const int k = k0 + lk;
        const int tap = k / a.Cin, ci = k - tap * a.Cin;        // Cin % 4 == 0: a float4 never straddles two input rows
        const int r = r0 + tap;
        const bool ok = (ar < M) && r >= 0 && r < a.Tin;
        const float4 av = ok ? *reinterpret_cast<const float4*>(abase + static_cast<size_t>(r) * a.Cin + ci)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lrow < BN) wv = *reinterpret_cast<const float4*>(wrow + k);
        __syncthreads();
        As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
        if (lrow < BN) { Ws[lk + 0][lrow] = wv.x; Ws[lk + 1][lrow] = wv.y; Ws[lk + 2][lrow] = wv.z; Ws[lk + 3][lrow] = wv.w; }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float b0 = Ws[kk][tx * 3], b1 = Ws[kk][tx * 3 + 1], b2 = Ws[kk][tx * 3 + 2];
            const float av4[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(av4[i], b0, acc[i][0]);
                acc[i][1] = fmaf(av4[i], b1, acc[i][1]);
                acc[i][2] = fmaf(av4[i], b2, acc[i][2]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long row = m0 + ty * 4 + i;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int n = n0 + tx * 3 + j;
            float v = acc[i][j] + a.bias[n];
            if (RES) v += a.resid[row * a.N + n];
            if (SNAKE) v = snake_f<false>(v, a.alpha[n], a.inv_alpha[n]);
            a.out[row * a.N + n] = v;
        }
    }
}

void launch_enc_gemm(const EncGemm& a, cudaStream_t st) {
    const long long M = static_cast<long long>(a.S) * a.Tout;
    dim3 grid(a.N / 48, static_cast<unsigned>((M + 63) / 64));
    const bool res = a.resid != nullptr, sn = a.alpha != nullptr;
    if (res && sn) k_enc_gemm<true, true><<<grid, 256, 0, st>>>(a);
    else if (res) k_enc_gemm<true, false><<<grid, 256, 0, st>>>(a);
    else if (sn) k_enc_gemm<false, true><<<grid, 256, 0, st>>>(a);
    else k_enc_gemm<false, false><<<grid, 256, 0, st>>>(a);
}

// ---------------------------------------------------------------------------------------------- final depthwise k7
__global__ void __launch_bounds__(256)
k_enc_dw7(const float* __restrict__ x, int S, int T, int C, const float* __restrict__ w /*[7][C]*/, const float* __restrict__ b,
          float* __restrict__ out) {
    const int c4n = C / 4;
    const long long total = static_cast<long long>(S) * T * c4n;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % c4n) * 4;
        const long long row = idx / c4n;
        const int s = static_cast<int>(row / T), t = static_cast<int>(row % T);
        float4 acc = *reinterpret_cast<const float4*>(b + c);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int tt = t + j - 3;
            if (tt < 0 || tt >= T) continue;
            const float4 xv = *reinterpret_cast<const float4*>(x + (static_cast<size_t>(s) * T + tt) * C + c);
            const float4 wv = *reinterpret_cast<const float4*>(w + j * C + c);
            acc.x = fmaf(wv.x, xv.x, acc.x); acc.y = fmaf(wv.y, xv.y, acc.y);
            acc.z = fmaf(wv.z, xv.z, acc.z); acc.w = fmaf(wv.w, xv.w, acc.w);
        }
        *reinterpret_cast<float4*>(out + row * C + c) = acc;
    }
}

// ---------------------------------------------------------------------------------------------- one RVQ level
// One CTA per pooled step.  dist = |e|^2 - 2 e.c + |c|^2 on the L2-normalised encoding e and codebook c, evaluated in the
// same order as the reference expression (snac/vq.py decode_latents); the first index of the minimum wins, as torch.max
// returns it.  cb_n: normalised codebook [4096][8]; cb2: its row norms^2.
struct EncVq {
    int S, T0, stride;
    float* residual;            // [S][T0][768], updated in place
    const float* in_w;          // [8][768]
    const float* in_b;          // [8]
    const float* cb;            // [4096][8] codebook
    const float* cb_n;          // [4096][8] normalised
    const float* cb2;           // [4096]
    const float* out_w;         // [768][8]
    const float* out_b;         // [768]
    int32_t* codes;             // [S][T0 / stride]
    float* best_dist;           // optional [S][T0 / stride]: the winning distance (tests)
};

__global__ void __launch_bounds__(256)
k_enc_vq(const EncVq a) {
    __shared__ float zp[kLatent];
    __shared__ float e[kCodeDim];
    __shared__ float rd[8];
    __shared__ int ri[8];
    __shared__ int win;
    const int Tp = a.T0 / a.stride;
    const int s = blockIdx.x / Tp, u = blockIdx.x % Tp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* base = a.residual + (static_cast<size_t>(s) * a.T0 + static_cast<size_t>(u) * a.stride) * kLatent;
    // avg_pool1d over the level's stride
    for (int c = tid; c < kLatent; c += 256) {
        float v = 0.f;
        for (int j = 0; j < a.stride; ++j) v += base[j * kLatent + c];
        zp[c] = a.stride > 1 ? v / static_cast<float>(a.stride) : v;
    }
    __syncthreads();
    // in_proj: warp w computes output channel w
    {
        float v = 0.f;
        for (int c = lane; c < kLatent; c += 32) v = fmaf(a.in_w[warp * kLatent + c], zp[c], v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) e[warp] = v + a.in_b[warp];
    }
    __syncthreads();
    // F.normalize: e / max(||e||, 1e-12)
    float en[kCodeDim], e2 = 0.f;
    {
        float n2 = 0.f;
#pragma unroll
        for (int d = 0; d < kCodeDim; ++d) n2 = fmaf(e[d], e[d], n2);
        const float inv = 1.f / fmaxf(sqrtf(n2), 1e-12f);
#pragma unroll
        for (int d = 0; d < kCodeDim; ++d) { en[d] = e[d] * inv; e2 = fmaf(en[d], en[d], e2); }
    }
    float best = INFINITY;
    int bi = 0;
    for (int k = tid; k < kCodebook; k += 256) {                 // ascending k per thread: the first minimum is kept
        const float4 c0 = *reinterpret_cast<const float4*>(a.cb_n + k * kCodeDim);
        const float4 c1 = *reinterpret_cast<const float4*>(a.cb_n + k * kCodeDim + 4);
        float dot = en[0] * c0.x;
        dot = fmaf(en[1], c0.y, dot); dot = fmaf(en[2], c0.z, dot); dot = fmaf(en[3], c0.w, dot);
        dot = fmaf(en[4], c1.x, dot); dot = fmaf(en[5], c1.y, dot); dot = fmaf(en[6], c1.z, dot); dot = fmaf(en[7], c1.w, dot);
        const float dist = (e2 - 2.f * dot) + a.cb2[k];
        if (dist < best) { best = dist; bi = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { rd[warp] = best; ri[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        float b = rd[0]; int i = ri[0];
        for (int w = 1; w < 8; ++w) if (rd[w] < b || (rd[w] == b && ri[w] < i)) { b = rd[w]; i = ri[w]; }
        win = i;
        a.codes[static_cast<size_t>(s) * Tp + u] = i;
        if (a.best_dist) a.best_dist[static_cast<size_t>(s) * Tp + u] = b;
    }
    __syncthreads();
    // z_q = out_proj(codebook[win]); residual -= repeat_interleave(z_q, stride)
    float cv[kCodeDim];
#pragma unroll
    for (int d = 0; d < kCodeDim; ++d) cv[d] = a.cb[win * kCodeDim + d];
    for (int c = tid; c < kLatent; c += 256) {
        float v = 0.f;
#pragma unroll
        for (int d = 0; d < kCodeDim; ++d) v = fmaf(a.out_w[c * kCodeDim + d], cv[d], v);
        v += a.out_b[c];
        for (int j = 0; j < a.stride; ++j) base[j * kLatent + c] -= v;
    }
}

// ---------------------------------------------------------------------------------------------- codes -> token ids
// Inverse of k_unpack (modal_audio_stream.py:156-188): frame f = [c0[f], c1[2f], c2[4f], c2[4f+1], c1[2f+1], c2[4f+2],
// c2[4f+3]], token = code + 4096 * position (+ 128266 for raw LLM ids).
__global__ void k_pack_tokens(const int32_t* __restrict__ c0, const int32_t* __restrict__ c1, const int32_t* __restrict__ c2,
                              int B, int F, int raw_ids, int32_t* __restrict__ tok) {
    const long long total = static_cast<long long>(B) * F * kFrame;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(idx % kFrame);
        const size_t bf = static_cast<size_t>(idx / kFrame);
        int c;
        switch (p) {
            case 0: c = c0[bf]; break;
            case 1: c = c1[2 * bf]; break;
            case 4: c = c1[2 * bf + 1]; break;
            case 2: c = c2[4 * bf]; break;
            case 3: c = c2[4 * bf + 1]; break;
            case 5: c = c2[4 * bf + 2]; break;
            default: c = c2[4 * bf + 3]; break;
        }
        tok[idx] = c + 4096 * p + (raw_ids ? kTokenAudioBase : 0);
    }
}

}  // namespace
}  // namespace snacb

// =================================================================================================
// C ABI
// =================================================================================================
using namespace snacb;

struct snacb_encoder_s {
    int device = 0;
    std::string err;
    std::vector<void*> allocs;
    float *conv0_w = nullptr, *conv0_b = nullptr;
    struct Res { float *a1, *i1, *dw_w, *dw_b, *a2, *i2, *pw_w, *pw_b; };
    struct Blk { int Cin, s; Res r[3]; float *alpha, *inv, *cw, *cb; } blk[4]{};
    float *fin_w = nullptr, *fin_b = nullptr;
    struct Lv { float *in_w, *in_b, *cb, *cb_n, *cb2, *out_w, *out_b; } lv[3]{};
    float* ws[3] = {nullptr, nullptr, nullptr};
    size_t ws_bytes = 0;
    uint64_t launches = 0;
};

namespace {

thread_local std::string g_enc_create_error;

int efail(snacb_encoder e, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e) e->err = buf; else g_enc_create_error = buf;
    return code;
}

#define ECK(e, call)                                                                                          \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return efail(e, SNACB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int up(snacb_encoder e, float** out, const std::vector<float>& v) {
    ECK(e, cudaMalloc(reinterpret_cast<void**>(out), v.size() * sizeof(float)));
    e->allocs.push_back(*out);
    ECK(e, cudaMemcpy(*out, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}
std::vector<float> hv(const float* p, size_t n) { return std::vector<float>(p, p + n); }
std::vector<float> inv_a(const float* a, size_t n) {
    std::vector<float> r(n);
    for (size_t i = 0; i < n; ++i) r[i] = 1.0f / (a[i] + 1e-9f);
    return r;
}
// depthwise [C][1][7] -> tap-major [7][C]
std::vector<float> dw_t(const float* w, int C) {
    std::vector<float> r(static_cast<size_t>(7) * C);
    for (int c = 0; c < C; ++c) for (int j = 0; j < 7; ++j) r[static_cast<size_t>(j) * C + c] = w[c * 7 + j];
    return r;
}
// strided conv [Cout][Cin][k] -> [Cout][k * Cin + ci]
std::vector<float> conv_pack(const float* w, int Cout, int Cin, int k) {
    std::vector<float> r(static_cast<size_t>(Cout) * Cin * k);
    for (int o = 0; o < Cout; ++o)
        for (int ci = 0; ci < Cin; ++ci)
            for (int j = 0; j < k; ++j)
                r[(static_cast<size_t>(o) * k + j) * Cin + ci] = w[(static_cast<size_t>(o) * Cin + ci) * k + j];
    return r;
}

}  // namespace

extern "C" {

const char* snacb_encoder_last_error(snacb_encoder e) { return e ? e->err.c_str() : g_enc_create_error.c_str(); }

int snacb_encoder_create(snacb_encoder* out, const snacb_encoder_weights* w, int device) {
    if (!out || !w) return SNACB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return efail(nullptr, SNACB_ERR_NO_GPU, "snacb_encoder_create: no CUDA device (there is no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return efail(nullptr, SNACB_ERR_ARG, "snacb_encoder_create: bad device %d", device);
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
        return efail(nullptr, SNACB_ERR_NO_GPU, "snacb_encoder_create: device %d is not sm_100 (this library is built for sm_100a only)", device);
    snacb_encoder e = new (std::nothrow) snacb_encoder_s();
    if (!e) return SNACB_ERR_NOMEM;
    e->device = device;
    int rc = 0;
    auto fin = [&](int code) { g_enc_create_error = e->err; snacb_encoder_destroy(e); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return fin(efail(e, SNACB_ERR_CUDA, "cudaSetDevice failed"));
#define UP(dst, vec) do { rc = up(e, &(dst), (vec)); if (rc) return fin(rc); } while (0)
    UP(e->conv0_w, dw_t(w->conv0_w, kEncDim));
    UP(e->conv0_b, hv(w->conv0_b, kEncDim));
    int C = kEncDim;
    for (int bi = 0; bi < 4; ++bi) {
        auto& b = e->blk[bi];
        const auto& s = w->block[bi];
        b.Cin = C; b.s = kEncRates[bi];
        for (int ri = 0; ri < 3; ++ri) {
            const auto& r = s.res[ri];
            UP(b.r[ri].a1, hv(r.alpha1, C)); UP(b.r[ri].i1, inv_a(r.alpha1, C));
            UP(b.r[ri].dw_w, dw_t(r.dw_w, C)); UP(b.r[ri].dw_b, hv(r.dw_b, C));
            UP(b.r[ri].a2, hv(r.alpha2, C)); UP(b.r[ri].i2, inv_a(r.alpha2, C));
            UP(b.r[ri].pw_w, hv(r.pw_w, static_cast<size_t>(C) * C)); UP(b.r[ri].pw_b, hv(r.pw_b, C));
        }
        UP(b.alpha, hv(s.alpha, C)); UP(b.inv, inv_a(s.alpha, C));
        UP(b.cw, conv_pack(s.conv_w, 2 * C, C, 2 * b.s)); UP(b.cb, hv(s.conv_b, 2 * C));
        C *= 2;
    }
    UP(e->fin_w, dw_t(w->final_w, kLatent));
    UP(e->fin_b, hv(w->final_b, kLatent));
    for (int l = 0; l < 3; ++l) {
        auto& v = e->lv[l];
        UP(v.in_w, hv(w->in_proj_w[l], static_cast<size_t>(kCodeDim) * kLatent));
        UP(v.in_b, hv(w->in_proj_b[l], kCodeDim));
        UP(v.cb, hv(w->codebook[l], static_cast<size_t>(kCodebook) * kCodeDim));
        std::vector<float> cn(static_cast<size_t>(kCodebook) * kCodeDim), c2(kCodebook);
        for (int k = 0; k < kCodebook; ++k) {                      // F.normalize(codebook) and its row norms^2, fp32
            float n2 = 0.f;
            for (int d = 0; d < kCodeDim; ++d) n2 = fmaf(w->codebook[l][k * kCodeDim + d], w->codebook[l][k * kCodeDim + d], n2);
            const float inv = 1.f / fmaxf(sqrtf(n2), 1e-12f);
            float s2 = 0.f;
            for (int d = 0; d < kCodeDim; ++d) {
                const float v1 = w->codebook[l][k * kCodeDim + d] * inv;
                cn[static_cast<size_t>(k) * kCodeDim + d] = v1;
                s2 = fmaf(v1, v1, s2);
            }
            c2[k] = s2;
        }
        UP(v.cb_n, cn); UP(v.cb2, c2);
        UP(v.out_w, hv(w->out_proj_w[l], static_cast<size_t>(kLatent) * kCodeDim));
        UP(v.out_b, hv(w->out_proj_b[l], kLatent));
    }
#undef UP
    *out = e;
    return SNACB_OK;
}

void snacb_encoder_destroy(snacb_encoder e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (void* p : e->allocs) cudaFree(p);
    for (auto& p : e->ws) if (p) cudaFree(p);
    delete e;
}

int snacb_encode_frames(int n_samples) { return n_samples <= 0 ? 0 : (n_samples + 2047) / 2048; }

int snacb_encode(snacb_encoder e, const float* audio, int B, int n_samples, int audio_stride, int32_t* c0, int32_t* c1,
                 int32_t* c2, float* latent, float* best_dist, void* stream) {
    if (!e) return SNACB_ERR_ARG;
    if (B < 0 || n_samples < 0 || audio_stride < n_samples) return efail(e, SNACB_ERR_ARG, "snacb_encode: bad sizes B=%d n=%d stride=%d", B, n_samples, audio_stride);
    if (B == 0 || n_samples == 0) return SNACB_OK;
    if (!audio || !c0 || !c1 || !c2) return efail(e, SNACB_ERR_ARG, "snacb_encode: null pointer");
    const int F = snacb_encode_frames(n_samples);
    if (F > 16384) return efail(e, SNACB_ERR_ARG, "snacb_encode: %d frames is too long", F);
    ECK(e, cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int T = 2048 * F;                                            // preprocess(): right-padded length
    const size_t need = static_cast<size_t>(B) * T * kEncDim * sizeof(float);   // T*C is the same (or smaller) at every stage
    if (need > e->ws_bytes) {
        for (auto& p : e->ws) if (p) { ECK(e, cudaFree(p)); p = nullptr; }
        e->ws_bytes = 0;
        for (auto& p : e->ws) ECK(e, cudaMalloc(reinterpret_cast<void**>(&p), need));
        e->ws_bytes = need;
    }
    float *h = e->ws[0], *P = e->ws[1], *o = e->ws[2];
    auto grid1 = [](long long total) { long long b = (total + 255) / 256; return static_cast<unsigned>(b > 148LL * 32 ? 148LL * 32 : b); };
    k_enc_conv0<<<grid1(static_cast<long long>(B) * T * (kEncDim / 4)), 256, 0, st>>>(audio, B, T, n_samples, audio_stride,
                                                                                     e->conv0_w, e->conv0_b, h);
    e->launches++;
    int C = kEncDim, Tc = T;
    static const int dils[3] = {1, 3, 9};
    for (int bi = 0; bi < 4; ++bi) {
        auto& b = e->blk[bi];
        for (int ri = 0; ri < 3; ++ri) {
            auto& r = b.r[ri];
            ResUnitArgs ra{};
            ra.S = B; ra.T = Tc; ra.C = C; ra.dil = dils[ri]; ra.x = h;
            ra.alpha1 = r.a1; ra.inv_alpha1 = r.i1; ra.dw_w = r.dw_w; ra.dw_b = r.dw_b; ra.alpha2 = r.a2; ra.inv_alpha2 = r.i2;
            launch_respre_f32(ra, P, st);
            EncGemm g{};
            g.S = B; g.Tin = Tc; g.Tout = Tc; g.Cin = C; g.N = C; g.stride = 1; g.pad = 0; g.K = C;
            g.A = P; g.W = r.pw_w; g.bias = r.pw_b; g.resid = h; g.out = o;
            if (ri == 2) { g.alpha = b.alpha; g.inv_alpha = b.inv; }     // the block's Snake in front of the strided conv
            launch_enc_gemm(g, st);
            e->launches += 2;
            float* t = h; h = o; o = t;
        }
        EncGemm g{};
        g.S = B; g.Tin = Tc; g.Tout = Tc / b.s; g.Cin = C; g.N = 2 * C; g.stride = b.s; g.pad = (b.s + 1) / 2; g.K = 2 * b.s * C;
        g.A = h; g.W = b.cw; g.bias = b.cb; g.out = o;
        launch_enc_gemm(g, st);
        e->launches++;
        float* t = h; h = o; o = t;
        C *= 2; Tc /= b.s;
    }
    // final depthwise k7 -> z (the latent), then the three RVQ levels on the residual
    k_enc_dw7<<<grid1(static_cast<long long>(B) * Tc * (kLatent / 4)), 256, 0, st>>>(h, B, Tc, kLatent, e->fin_w, e->fin_b, o);
    e->launches++;
    ECK(e, cudaGetLastError());
    if (latent) ECK(e, cudaMemcpyAsync(latent, o, static_cast<size_t>(B) * Tc * kLatent * sizeof(float), cudaMemcpyDeviceToDevice, st));
    int32_t* codes[3] = {c0, c1, c2};
    static const int strides[3] = {4, 2, 1};
    size_t doff = 0;
    for (int l = 0; l < 3; ++l) {
        EncVq v{};
        v.S = B; v.T0 = Tc; v.stride = strides[l]; v.residual = o;
        v.in_w = e->lv[l].in_w; v.in_b = e->lv[l].in_b; v.cb = e->lv[l].cb; v.cb_n = e->lv[l].cb_n; v.cb2 = e->lv[l].cb2;
        v.out_w = e->lv[l].out_w; v.out_b = e->lv[l].out_b; v.codes = codes[l];
        v.best_dist = best_dist ? best_dist + doff : nullptr;
        doff += static_cast<size_t>(B) * (Tc / strides[l]);
        k_enc_vq<<<B * (Tc / strides[l]), 256, 0, st>>>(v);
        e->launches++;
    }
    ECK(e, cudaGetLastError());
    return SNACB_OK;
}

int snacb_pack_tokens(const int32_t* c0, const int32_t* c1, const int32_t* c2, int B, int frames, int flags, int32_t* tok,
                      void* stream) {
    if (B < 0 || frames < 0) return SNACB_ERR_ARG;
    if (B == 0 || frames == 0) return SNACB_OK;
    if (!c0 || !c1 || !c2 || !tok) return SNACB_ERR_ARG;
    const long long total = static_cast<long long>(B) * frames * kFrame;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    k_pack_tokens<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        c0, c1, c2, B, frames, (flags & SNACB_RAW_IDS) ? 1 : 0, tok);
    return cudaGetLastError() == cudaSuccess ? SNACB_OK : SNACB_ERR_CUDA;
}

uint64_t snacb_encoder_launches(snacb_encoder e) { return e ? e->launches : 0; }

}  // extern "C"
