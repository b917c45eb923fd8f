// Host side of libsnacb: handle, weight packing, workspace, the per-group kernel pipeline and the
// C ABI declared in include/snacb.h.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/snacb.h"
#include "kernels.h"

using namespace snacb;

namespace {

constexpr size_t kMaxActMaps = 4096;
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

thread_local std::string g_create_error;

struct ResW {
    float *alpha1, *inv1, *dw_w, *dw_b, *alpha2, *inv2, *pw_f32, *pw_b;
    void* pw_h[2];          // 16-bit copies: [0] bf16, [1] fp16
    CUtensorMap tm_pw[2];   // box (64, min(C,256)) for k_resunit_tc
    void* pw_x3 = nullptr;  // bf16x3 path: [C][2C] = [W_hi | W_lo]
    void* pwc_h = nullptr;  // fp16 1x1 weights with 1/alpha2 folded into the K columns (fp16 chain kernel)
    CUtensorMap tm_pwc;
};
struct BlockW {
    int Cin, Cout, s;
    float *alpha, *inv_alpha;           // Snake in front of the ConvTranspose (applied by the producer epilogue)
    float *ct_f32, *ct_b;               // packed [s*Cout][2*Cin]
    void* ct_h[2];
    float* nz_f32;                      // [Cout][Cout]
    void* nz_h[2];
    void *ct_x3 = nullptr, *nz_x3 = nullptr;   // bf16x3 path: per tap [W_hi | W_hi | W_lo] (K'' = 3K)
    ResW res[3];
    float* bias_cum;                    // [3][Cout] running sums of the ResidualUnit 1x1 biases (fused chain)
    bool chain[2];                      // the fused NoiseBlock + ResidualUnit chain covers this block ([0] bf16, [1] fp16)
    ChainSpan spans[3][kChainWarps][kChainSpans];
    ChainSpan spans_ws[3][kChainWarps][kChainSpans];   // warp-specialised chain kernel (kernels_chain_ws.cu): quads inside a block
    ChainSpan spans_carry[3][kChainWarps][kChainSpans]; // schedule of a carry-top tile (kernels_chain.cu)
    bool spans_carry_ok = false;
    ChainSpan spans_last[3][kChainWarps][kChainSpans]; // schedule of a short last tile of type / size spans_last_key (cached)
    int spans_last_key = 0;
    bool spans_last_ok = false;
    bool fold = false;                  // fp16 chain: the alpha-folded formulation is numerically safe for this block's
                                        // Snake alphas (chain_fold_safe); otherwise the general fp32-Snake variant runs
};
struct Tap {
    std::string name;
    int64_t rows, cols;
    float* dev;
};
struct TmapKey {
    const void* p; int c, t, s, tb, wb, dt, sw;
    bool operator<(const TmapKey& o) const {
        return std::tie(p, c, t, s, tb, wb, dt, sw) < std::tie(o.p, o.c, o.t, o.s, o.tb, o.wb, o.dt, o.sw);
    }
};

}  // namespace

struct snacb_handle_s {
    int device = 0;
    int sm_count = 148;
    std::string err;
    cudaStream_t own_stream = nullptr;
    PFN_tmapEncodeTiled encode = nullptr;

    std::vector<void*> allocs;          // weight allocations
    VqStemWeights vq{};
    float *stem_pw_f32 = nullptr, *stem_pw_b = nullptr;
    void* stem_pw_h[2] = {nullptr, nullptr};
    void* stem_pw_x3 = nullptr;         // bf16x3 path: [1024][3 * 768]
    BlockW blk[4]{};
    float *tail_alpha = nullptr, *tail_inv = nullptr, *tail_w = nullptr;
    float tail_b = 0.f;
    std::map<std::pair<const void*, int>, CUtensorMap> wmaps;   // (weight ptr, box rows) -> map
    std::map<TmapKey, CUtensorMap> amaps;
    size_t max_act_maps = kMaxActMaps;  // SNACB_TMAP_CACHE=n shrinks the cache (tests: eviction inside a launch sequence)

    // workspace
    size_t group_bytes = static_cast<size_t>(1) << 30;   // per activation buffer; launches are per group
    void* ws_buf[3] = {nullptr, nullptr, nullptr};
    size_t ws_buf_bytes = 0;
    void* ws_a0 = nullptr; size_t ws_a0_bytes = 0;
    void* ws_split = nullptr; size_t ws_split_bytes = 0;      // bf16x3 path: the split A operand of the running GEMM
    void* chain_carry = nullptr; size_t chain_carry_bytes = 0; // chain kernel: per-CTA rows handed from tile to tile of a strip
    int32_t* ws_codes = nullptr; size_t ws_codes_elems = 0;
    int* tile_counter = nullptr;                              // dynamic tile scheduler of the chain kernel
    int32_t* st_tok = nullptr; size_t st_tok_elems = 0;       // decode_host staging
    int16_t* st_pcm = nullptr; size_t st_pcm_elems = 0;
    // pipelined host boundary (snacb_decode_host_submit / _wait): two staging slots, a copy stream, per-slot events
    int32_t* pl_tok[2] = {nullptr, nullptr}; size_t pl_tok_bytes[2] = {0, 0};
    int16_t* pl_pcm[2] = {nullptr, nullptr}; size_t pl_pcm_bytes[2] = {0, 0};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t pl_done[2] = {nullptr, nullptr}, pl_copied[2] = {nullptr, nullptr};
    uint64_t pl_submitted = 0, pl_waited = 0;

    std::vector<Tap> taps;
    uint64_t launches = 0, streams = 0;
    bool res_v1 = false;                // SNACB_RES_V1=1: use the non-persistent ResidualUnit kernel
    bool chain_ws = false;              // SNACB_CHAIN_WS=1: warp-specialised, block-pipelined chain kernel where it applies
    bool no_fold = false;               // SNACB_NO_FOLD=1: general (fp32 Snake) chain variant even where the folded one is safe (A/B)
    int chain_prof = 0;                 // SNACB_CHAIN_PROF=1|2: in-kernel clock64 phase timing of k_chain, printed per launch (debug)
    unsigned chain_jitter = 0;          // SNACB_CHAIN_JITTER=seed: pseudo-random per-warp delays inside k_chain's in-place prologue (race detector)
    bool bf16_plain = false;            // SNACB_BF16_PLAIN=1: precision bf16 = bf16 STORAGE and single bf16 operands (34 dB; the
                                        // pre-bf16x3 behaviour, kept for A/B) instead of the bf16x3 path
    bool no_chain = false;              // SNACB_NO_CHAIN=1: per-layer kernels instead of the fused chain
    bool no_convt_res = false;          // SNACB_NO_CONVT_RES=1: generic k_gemm_tc for every ConvTranspose
    bool no_trim = false;               // SNACB_NO_TRIM=1: sliced output still decodes every sample of the window

    // optional per-launch CUDA-event timing (snacb_profile / snacb_profile_report)
    struct ProfRec { int name_id; cudaEvent_t a, b; };
    bool prof_on = false;
    std::vector<std::string> prof_names;
    std::vector<ProfRec> prof_recs;
};

namespace {

int fail(snacb_handle h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(h, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(h, SNACB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

void prof_begin(snacb_handle h, const char* name, cudaStream_t st) {
    if (!h->prof_on) return;
    int id = -1;
    for (size_t i = 0; i < h->prof_names.size(); ++i) if (h->prof_names[i] == name) { id = static_cast<int>(i); break; }
    if (id < 0) { id = static_cast<int>(h->prof_names.size()); h->prof_names.push_back(name); }
    snacb_handle_s::ProfRec r{id, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    h->prof_recs.push_back(r);
}
void prof_end(snacb_handle h, cudaStream_t st) {
    if (!h->prof_on || h->prof_recs.empty()) return;
    cudaEventRecord(h->prof_recs.back().b, st);
}

template <typename T>
int dev_alloc(snacb_handle h, T** out, size_t n) {
    void* p = nullptr;
    CK(h, cudaMalloc(&p, n * sizeof(T)));
    h->allocs.push_back(p);
    *out = static_cast<T*>(p);
    return 0;
}
int upload_f32(snacb_handle h, float** out, const std::vector<float>& v) {
    int rc = dev_alloc(h, out, v.size());
    if (rc) return rc;
    CK(h, cudaMemcpy(*out, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}
// both 16-bit copies of a weight matrix: out[0] bf16, out[1] fp16 (round to nearest even)
int upload_h16(snacb_handle h, void** out, const std::vector<float>& v) {
    std::vector<__nv_bfloat16> b(v.size());
    std::vector<__half> f(v.size());
    for (size_t i = 0; i < v.size(); ++i) { b[i] = __float2bfloat16_rn(v[i]); f[i] = __float2half_rn(v[i]); }
    __nv_bfloat16* db; __half* df;
    int rc = dev_alloc(h, &db, v.size());
    if (rc) return rc;
    rc = dev_alloc(h, &df, v.size());
    if (rc) return rc;
    CK(h, cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(df, f.data(), f.size() * 2, cudaMemcpyHostToDevice));
    out[0] = db; out[1] = df;
    return 0;
}
// bf16x3 weights: w [rows][ntaps * K] -> bf16 [rows][ntaps * parts * K]; per tap [hi | hi | lo] (parts = 3, against an
// A operand [hi | lo | hi]) or [hi | lo] (parts = 2, against [a | a]); hi = bf16(w), lo = bf16(w - hi)
int upload_x3(snacb_handle h, void** out, const std::vector<float>& v, size_t rows, int ntaps, int K, int parts) {
    std::vector<__nv_bfloat16> b(rows * ntaps * parts * K);
    for (size_t r = 0; r < rows; ++r)
        for (int t = 0; t < ntaps; ++t)
            for (int k = 0; k < K; ++k) {
                const float w = v[(r * ntaps + t) * K + k];
                const __nv_bfloat16 hi = __float2bfloat16_rn(w);
                const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
                __nv_bfloat16* d = b.data() + (r * ntaps + t) * static_cast<size_t>(parts) * K + k;
                d[0] = hi;
                if (parts == 3) { d[K] = hi; d[2 * K] = lo; } else d[K] = lo;
            }
    __nv_bfloat16* db;
    int rc = dev_alloc(h, &db, b.size());
    if (rc) return rc;
    CK(h, cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    *out = db;
    return 0;
}
std::vector<float> vec(const float* p, size_t n) { return std::vector<float>(p, p + n); }
std::vector<float> inv_alpha(const float* a, size_t n) {
    std::vector<float> r(n);
    for (size_t i = 0; i < n; ++i) r[i] = 1.0f / (a[i] + 1e-9f);     // (alpha + 1e-9).reciprocal(), fp32
    return r;
}
// [C][7] -> [7][C]
std::vector<float> dw_transpose(const float* w, int C) {
    std::vector<float> r(static_cast<size_t>(7) * C);
    for (int c = 0; c < C; ++c)
        for (int j = 0; j < 7; ++j) r[static_cast<size_t>(j) * C + c] = w[static_cast<size_t>(c) * 7 + j];
    return r;
}
// ConvTranspose1d weight [Cin][Cout][2s] -> per-output-phase 2-tap GEMM weight [s*Cout][2*Cin]:
//   Wp[p*Cout + o][tap*Cin + c] = W[c][o][((p + s/2) mod s) + tap*s]
// (y[o, m*s+p] = sum_c x[c, m+sh-0] W[c,o,r] + x[c, m+sh-1] W[c,o,r+s], r = (p+s/2) mod s, sh = [p >= s/2])
std::vector<float> pack_convt(const float* w, int Cin, int Cout, int s) {
    std::vector<float> r(static_cast<size_t>(s) * Cout * 2 * Cin);
    const int k = 2 * s;
    for (int p = 0; p < s; ++p)
        for (int o = 0; o < Cout; ++o)
            for (int tap = 0; tap < 2; ++tap)
                for (int c = 0; c < Cin; ++c)
                    r[(static_cast<size_t>(p) * Cout + o) * (2 * Cin) + static_cast<size_t>(tap) * Cin + c] =
                        w[(static_cast<size_t>(c) * Cout + o) * k + ((p + s / 2) % s) + tap * s];
    return r;
}

int make_tmap_2d(snacb_handle h, CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows,
                 int fp16) {
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = h->encode(m, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, SNACB_ERR_CUDA, "cuTensorMapEncodeTiled(2d %llu x %llu) failed: %d",
                                       (unsigned long long)rows, (unsigned long long)cols, (int)r);
    return 0;
}
int make_tmap_3d(snacb_handle h, CUtensorMap* m, const void* base, uint64_t C, uint64_t T, uint64_t S, uint32_t tbox,
                 uint32_t wbox, int fp16, int swizzle = 1) {
    cuuint64_t gdim[3] = {C, T, S};
    cuuint64_t gstr[2] = {C * 2, T * C * 2};
    cuuint32_t box[3] = {swizzle == 2 ? 32u : 64u, tbox, wbox};         // swizzle: 0 none, 1 128B (64-column box), 2 64B (32-column box)
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = h->encode(m, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, SNACB_ERR_CUDA, "cuTensorMapEncodeTiled(3d C=%llu T=%llu S=%llu) failed: %d",
                                       (unsigned long long)C, (unsigned long long)T, (unsigned long long)S, (int)r);
    return 0;
}
int weight_map(snacb_handle h, const CUtensorMap** out, const void* w, int rows, int cols, int box_rows, int fp16) {
    auto key = std::make_pair(w, box_rows);
    auto it = h->wmaps.find(key);
    if (it == h->wmaps.end()) {
        CUtensorMap m;
        int rc = make_tmap_2d(h, &m, w, cols, rows, box_rows, fp16);
        if (rc) return rc;
        it = h->wmaps.emplace(key, m).first;
    }
    *out = &it->second;
    return 0;
}
// Activation tensor maps are cached per (buffer, shape, box).  The map is returned BY VALUE: a launch sequence holds
// several maps at once, and a cache that handed out pointers would leave them dangling whenever an insertion evicts
// (the cache is bounded; a serving process with variable batch sizes / ranges sees thousands of distinct shapes).
int act_map(snacb_handle h, CUtensorMap* out, const void* base, int C, int T, int S, int tbox, int wbox,
            int fp16, int swizzle = 1) {
    TmapKey key{base, C, T, S, tbox, wbox, fp16, swizzle};
    auto it = h->amaps.find(key);
    if (it == h->amaps.end()) {
        if (h->amaps.size() >= h->max_act_maps) h->amaps.clear();
        CUtensorMap m;
        int rc = make_tmap_3d(h, &m, base, C, T, S, tbox, wbox, fp16, swizzle);
        if (rc) return rc;
        it = h->amaps.emplace(key, m).first;
    }
    *out = it->second;
    return 0;
}

void tile_boxes(int Tin, int* tbox, int* wbox) {
    if (Tin < 128 && (Tin & (Tin - 1)) == 0) { *tbox = Tin; *wbox = 128 / Tin; }
    else { *tbox = 128; *wbox = 1; }
}

int grow(snacb_handle h, void** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) { CK(h, cudaFree(*p)); *p = nullptr; *cap = 0; h->amaps.clear(); }
    CK(h, cudaMalloc(p, need));
    *cap = need;
    return 0;
}

template <typename T>
int add_tap(snacb_handle h, const char* name, const T* src, int64_t rows, int64_t cols, cudaStream_t st) {
    Tap t{name, rows, cols, nullptr};
    CK(h, cudaMalloc(reinterpret_cast<void**>(&t.dev), static_cast<size_t>(rows) * cols * sizeof(float)));
    launch_to_f32<T>(src, t.dev, static_cast<size_t>(rows) * cols, st);
    CK(h, cudaGetLastError());
    h->taps.push_back(t);
    return 0;
}
void clear_taps(snacb_handle h) {
    for (auto& t : h->taps) cudaFree(t.dev);
    h->taps.clear();
}

// The fp16 chain kernel's alpha-folded formulation forms dw_w * alpha2 / alpha1, dw_b * alpha2 and pw_w / alpha2 in
// half precision (kernels_chain.cu).  Trained Snake alphas may be ~0, negative or large; the fold is used only when every
// alpha of the block's three ResidualUnits lies in [2^-8, 2^6] in magnitude and every folded parameter stays well inside
// the fp16 range.  Otherwise the block runs the general variant, x + (alpha + 1e-9)^-1 sin^2(alpha x) in fp32, which for
// alpha = 0 returns x exactly as the reference does (snac layers.py snake()).
bool chain_fold_safe(const snacb_block_weights& s, int C) {
    constexpr float kLo = 1.0f / 256.0f, kHi = 64.0f, kMaxH = 3.0e4f;
    for (int ri = 0; ri < 3; ++ri) {
        const snacb_resunit_weights& r = s.res[ri];
        for (int c = 0; c < C; ++c) {
            const float a1 = r.alpha1[c], a2 = r.alpha2[c];
            if (!(fabsf(a1) >= kLo && fabsf(a1) <= kHi && fabsf(a2) >= kLo && fabsf(a2) <= kHi)) return false;   // also NaN
            if (!(fabsf(r.dw_b[c] * a2) <= kMaxH)) return false;
            for (int j = 0; j < 7; ++j)
                if (!(fabsf(r.dw_w[static_cast<size_t>(c) * 7 + j] * (a2 / a1)) <= kMaxH)) return false;
            for (int n = 0; n < C; ++n)
                if (!(fabsf(r.pw_w[static_cast<size_t>(n) * C + c] / a2) <= kMaxH)) return false;
        }
    }
    return true;
}

// SNACB_CHAIN_PROF (debug): per-phase clock64 sums of CTA 0 / thread 0 and per-CTA lifetimes of one k_chain launch
int chain_prof_begin(snacb_handle h, ChainArgs* ca, cudaStream_t st) {
    CK(h, cudaMalloc(reinterpret_cast<void**>(&ca->prof), 1024 * sizeof(unsigned long long)));
    CK(h, cudaMemsetAsync(ca->prof, 0, 1024 * sizeof(unsigned long long), st));
    return 0;
}
int chain_prof_report(snacb_handle h, const ChainArgs& ca, int bi, cudaStream_t st) {
    unsigned long long pv[1024];
    CK(h, cudaStreamSynchronize(st));
    CK(h, cudaMemcpy(pv, ca.prof, sizeof pv, cudaMemcpyDeviceToHost));
    CK(h, cudaFree(ca.prof));
    const int rows = chain_tile_rows(ca.C) - 2 * kChainHalo;
    const int tiles = ca.S * (((ca.t_n > 0 ? ca.t_n : ca.T) + rows - 1) / rows);
    const int slots = h->sm_count * (chain_warps(ca.C) == 8 ? 2 : 1);
    const int mine = (tiles + slots - 1) / slots;
    fprintf(stderr, "chain b%d C=%d: CTA0 cycles per tile (~%d tiles per CTA): nz+ld %llu noise %llu |", bi, ca.C, mine,
            pv[0] / mine, pv[1] / mine);
    for (int l = 0; l < 3; ++l)
        fprintf(stderr, " L%d: pre %llu pro %llu sync %llu mma+epi %llu |", l, pv[2 + 4 * l] / mine, pv[3 + 4 * l] / mine,
                pv[4 + 4 * l] / mine, pv[5 + 4 * l] / mine);
    fprintf(stderr, " store/load issue %llu\n", pv[14] / mine);
    const int grid = tiles < slots ? tiles : slots;
    unsigned long long mn = ~0ull, mx = 0, sum = 0;
    for (int c = 0; c < grid && c < 500; ++c) {
        const unsigned long long v = pv[20 + 2 * c];
        mn = v < mn ? v : mn; mx = v > mx ? v : mx; sum += v;
    }
    fprintf(stderr, "   CTA lifetimes (cycles): min %llu avg %llu max %llu\n", mn, sum / (grid < 500 ? grid : 500), mx);
    return 0;
}

int chain_ws_prof_report(snacb_handle h, const ChainArgs& ca, int bi, cudaStream_t st) {
    unsigned long long pv[32];
    CK(h, cudaStreamSynchronize(st));
    CK(h, cudaMemcpy(pv, ca.prof, sizeof pv, cudaMemcpyDeviceToHost));
    CK(h, cudaFree(ca.prof));
    const int rows = chain_ws_tile_rows(ca.C) - 2 * kChainHalo;
    const int tiles = ca.S * (((ca.t_n > 0 ? ca.t_n : ca.T) + rows - 1) / rows);
    const int mine = (tiles + h->sm_count - 1) / h->sm_count;
    fprintf(stderr, "chain_ws b%d C=%d: cycles per tile (~%d tiles per CTA) | P warp 0:", bi, ca.C, mine);
    for (int l = 0; l < 3; ++l)
        fprintf(stderr, " L%d wait-S1 %llu pre+bar %llu spans %llu |", l, pv[4 * l] / mine, pv[1 + 4 * l] / mine, pv[2 + 4 * l] / mine);
    fprintf(stderr, " next-tile wait %llu || E warp 0:", pv[12] / mine);
    for (int i = 0; i < 4; ++i) fprintf(stderr, " B%d wait-MMA %llu drain %llu |", i, pv[17 + 2 * i] / mine, pv[18 + 2 * i] / mine);
    fprintf(stderr, " || IO idle polls %llu\n", pv[16] / mine);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// One group of S streams through the whole path.
// ------------------------------------------------------------------------------------------------
struct Rng { int lo, hi; };

// A stateful streaming step (snacb_session_step): every stage writes to its OWN persistent buffer in absolute row
// coordinates and processes only the rows that became computable since the previous step, [lo, hi) per stage; the rows
// below lo are read back from the same buffers (they were final when written).  nullptr = the stateless decode.
struct SessionPlan {
    int32_t *c0, *c1, *c2;   // unpacked codes [S][Fmax], [S][2 Fmax], [S][4 Fmax]
    void* a0;                // stem depthwise output [S][4 Fmax][768]
    void* stem;              // stem 1x1 (+ Snake) output
    void* ct[4];             // ConvTranspose outputs
    void* nz[4];             // NoiseBlock outputs (blocks without the fused chain)
    void* res[4][3];         // ResidualUnit outputs (blocks without the fused chain); [2] is the block output
    void* out[4];            // block outputs (fused chain)
    Rng stem_r;              // latent steps of the stem
    Rng ct_in[4];            // ConvTranspose input rows
    Rng post[4];             // fused chain: output rows
    Rng nz_r[4], res_r[4][3];// per-layer rows of the blocks without the fused chain
    int n_slots;             // streams the buffers hold (tensor-map extent: a mapped launch addresses slots, not 0..S-1); 0 = S
    StreamMap map;           // per-stream slot / frame offset / window origin (device arrays; null: stream s = slot s at the
                             // reference position) -- streams at different positions advancing by the same number of frames
};

int run_group(snacb_handle h, const int32_t* tok, int S, int tok_stride, int F, int flags, const float* const* noise,
              uint64_t seed, int stream_offset, const int32_t* stream_keys, int out_lo, int out_hi, int16_t* pcm,
              float* wave, cudaStream_t st, const SessionPlan* plan = nullptr, int origin_frames = 0) {
    const bool f32 = (flags & SNACB_FP32) != 0;
    // precision bf16 = the bf16x3 path: activations STORED in fp16, every contraction a bf16 tcgen05.mma on operands split
    // so that the products are exact (k_split3 / k_respre16 in kernels_simt.cu): A'' = [A_hi | A_lo | A_hi] against
    // W'' = [W_hi | W_hi | W_lo] for the GEMMs fed from storage, [a | a] against [W_hi | W_lo] for the ResidualUnit 1x1s
    // (only [A_hi | A_lo] and [a] are materialised: the GEMM wraps its K chunk index, GemmArgs::a_wrap).
    // One kernel per layer (the fused chain keeps a single 16-bit tile copy and has no room for a second operand).
    const bool x3 = !f32 && (flags & SNACB_BF16) && !h->bf16_plain;
    const bool xf32 = f32 || (!x3 && (flags & SNACB_STREAM_FP32));      // residual stream dtype
    const int hk = ((flags & SNACB_BF16) && !x3) ? 0 : 1;      // 16-bit storage type: 0 bf16, 1 fp16
    const bool taps = (flags & SNACB_KEEP_TAPS) != 0;
    const int T0 = 4 * F;
    if (taps) clear_taps(h);
    enum { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
    const int dt_h = f32 ? DT_F32 : (hk ? DT_F16 : DT_BF16);   // dtype of operand tensors
    const int dt_x = xf32 ? DT_F32 : dt_h;                     // dtype of the residual stream
    auto tap_any = [&](const char* name, const void* p, int dt, int64_t rows, int64_t cols) -> int {
        if (!taps) return 0;
        if (dt == DT_F32) return add_tap(h, name, static_cast<const float*>(p), rows, cols, st);
        if (dt == DT_BF16) return add_tap(h, name, static_cast<const __nv_bfloat16*>(p), rows, cols, st);
        return add_tap(h, name, static_cast<const __half*>(p), rows, cols, st);
    };

    int32_t* c0 = plan ? plan->c0 : h->ws_codes;
    int32_t* c1 = plan ? plan->c1 : c0 + static_cast<size_t>(S) * F;
    int32_t* c2 = plan ? plan->c2 : c1 + static_cast<size_t>(S) * 2 * F;
    // the token -> code unpack is fused into the VQ gather (k_vq_stem reads the token rows); the separate k_unpack pass
    // runs only for SNACB_KEEP_TAPS decodes, whose callers may read the codes (snacb_unpack keeps the stand-alone call)
    const bool fused_unpack = !taps;
    if (!fused_unpack) {
        prof_begin(h, "unpack", st);
        launch_unpack(tok, S, tok_stride, F, (flags & SNACB_RAW_IDS) ? 1 : 0, c0, c1, c2, st);
        prof_end(h, st);
        h->launches++;
    }
    // ---- dead-sample trimming (sliced output): only the rows of each stage inside the receptive field of samples
    //      [out_lo, out_hi) are computed.  Backward range propagation; the stem is always computed in full, block 0 unless
    //      less than half of it is live.
    Rng ct_in[4], post[4];                 // ConvTranspose input rows; rows the post-ConvTranspose layers process
    bool trimmed[4] = {false, false, false, false};
    if (plan) {
        for (int bi = 0; bi < 4; ++bi) { trimmed[bi] = true; ct_in[bi] = plan->ct_in[bi]; post[bi] = plan->post[bi]; }
    } else {
        const int Tfin = 2048 * F;
        // output samples [out_lo, out_hi): the caller's range (snacb_decode_range), the reference's slice, or everything
        if (out_hi <= out_lo) {
            const bool sl = (flags & SNACB_EXTRACT_SLICE) && Tfin > 4096;
            out_lo = sl ? 2048 : 0; out_hi = sl ? 4096 : Tfin;
        }
        const bool want = (out_hi - out_lo < Tfin) && !f32 && !xf32 && !taps && !h->no_trim;
        int Tb[4], tt = T0;
        for (int bi = 0; bi < 4; ++bi) { tt *= h->blk[bi].s; Tb[bi] = tt; }
        auto clip = [](Rng r, int T) { return Rng{r.lo < 0 ? 0 : r.lo, r.hi > T ? T : r.hi}; };
        Rng need = clip(Rng{out_lo - 3, out_hi + 3}, Tb[3]);        // tail conv k7
        for (int bi = 3; bi >= 0; --bi) {
            const BlockW& b = h->blk[bi];
            const int Tinb = bi ? Tb[bi - 1] : T0;
            // block 0 (per-layer kernels, 128-row tiles) is trimmed only when less than half of it is needed: a short
            // range of a long prefix (snacb_decode_range); for the 4-frame window's slice 111 of its 128 rows are live
            const bool skip0 = bi == 0 && 2 * (need.hi - need.lo + 78) >= Tb[0];
            if (!want || skip0) { post[bi] = Rng{0, Tb[bi]}; ct_in[bi] = Rng{0, Tinb}; continue; }
            trimmed[bi] = true;
            Rng y;                                                  // ConvTranspose output rows that must be valid
            const bool unfused = (flags & SNACB_UNFUSED) != 0 || h->no_chain || x3;
            if (b.chain[hk] && !unfused) {
                const bool ws = h->chain_ws && chain_ws_supported(b.Cout, hk) && !plan;
                const int rows = (ws ? chain_ws_tile_rows(b.Cout) : chain_tile_rows(b.Cout)) - 2 * kChainHalo;
                const int n = (need.hi - need.lo + rows - 1) / rows;
                post[bi] = Rng{need.lo, need.lo + n * rows};
                y = clip(Rng{post[bi].lo - kChainHalo, post[bi].hi + kChainHalo}, Tb[bi]);
            } else {
                post[bi] = clip(Rng{need.lo - 39, need.hi + 39}, Tb[bi]);   // 3 * (1 + 3 + 9) rows of receptive field
                y = post[bi];
            }
            ct_in[bi] = clip(Rng{y.lo / b.s - 1, (y.hi - 1) / b.s + 2}, Tinb);
            need = ct_in[bi];
        }
    }

    prof_begin(h, "vq_stem", st);
    // latent steps the stem has to produce: the input rows of block 0's ConvTranspose (all of them unless block 0 is trimmed)
    const int stem_lo = plan ? plan->stem_r.lo : (trimmed[0] ? ct_in[0].lo : 0);
    const int stem_hi = plan ? plan->stem_r.hi : (trimmed[0] ? ct_in[0].hi : T0);
    void* const a0 = plan ? plan->a0 : h->ws_a0;
    const bool stem_live = stem_hi > stem_lo;
    const StreamMap smap = plan ? plan->map : StreamMap{nullptr, nullptr, nullptr};
    const bool mapped = smap.slot != nullptr || smap.off != nullptr || smap.org != nullptr;
    const int S_buf = (plan && plan->n_slots > 0) ? plan->n_slots : S;      // stream extent of the buffers (tensor maps)
    if (!stem_live) {}
    else if (f32) launch_vq_stem<float>(c0, c1, c2, fused_unpack ? tok : nullptr, tok_stride, (flags & SNACB_RAW_IDS) ? 1 : 0, S, F, stem_lo, stem_hi, h->vq, static_cast<float*>(a0), st, smap);
    else if (hk) launch_vq_stem<__half>(c0, c1, c2, fused_unpack ? tok : nullptr, tok_stride, (flags & SNACB_RAW_IDS) ? 1 : 0, S, F, stem_lo, stem_hi, h->vq, static_cast<__half*>(a0), st, smap);
    else launch_vq_stem<__nv_bfloat16>(c0, c1, c2, fused_unpack ? tok : nullptr, tok_stride, (flags & SNACB_RAW_IDS) ? 1 : 0, S, F, stem_lo, stem_hi, h->vq, static_cast<__nv_bfloat16*>(a0), st, smap);
    prof_end(h, st);
    if (stem_live) h->launches++;
    CK(h, cudaGetLastError());
    {
        int rc = tap_any("stem_dw", a0, dt_h, (int64_t)S * T0, kLatent);
        if (rc) return rc;
    }

    // stateless decode: two ping-pong activation buffers; session step: one persistent buffer per stage
    auto other = [&](void* p) -> void* { return p == h->ws_buf[0] ? h->ws_buf[1] : h->ws_buf[0]; };
    void* cur = plan ? plan->stem : h->ws_buf[0];
    float* P = static_cast<float*>(h->ws_buf[2]);
    auto live = [](Rng r) { return r.hi > r.lo; };

    // Wx3: the bf16x3 pack of the weights; a_split: A is already the split operand (ResidualUnit: k_respre16 wrote [a | a]
    // into ws_split and a.K is its width), otherwise an fp16 tensor that k_split3 expands to 3K columns first
    auto gemm = [&](const char* pname, int epi, bool out_f32, GemmArgs& a, const void* A, const float* Wf,
                    void* const* Wh, int Wrows, int Wcols, const void* Wx3 = nullptr, bool a_split = false) -> int {
        tile_boxes(a.Tin, &a.Tbox, &a.Wbox);
        a.map = smap; a.rpf = a.Tin / F;                  // session: per-stream slot / offset / origin (A rows per frame)
        if (a.Wbox != 1) a.t_n = 0;
        else if (!f32 && a.t_n > 0 && a.t_n <= 64) {
            // the trimmed rows of a stream fill at most half a 128-row tile (block-1 ConvTranspose of the sliced call:
            // 57 of 128 rows): pack the same row range of several streams into one tile
            int tb = 8;
            while (tb < a.t_n) tb <<= 1;
            a.Tbox = tb; a.Wbox = 128 / tb;
        }
        a.seed = seed;
        a.stream_offset = stream_offset; a.stream_keys = stream_keys;
        h->launches++;
        if (f32) {
            prof_begin(h, pname, st);
            launch_gemm_f32(epi, a, static_cast<const float*>(A), Wf, st);
            prof_end(h, st);
            CK(h, cudaGetLastError());
            return 0;
        }
        CUtensorMap ma;
        const CUtensorMap* mw;
        if (x3) {
            prof_begin(h, pname, st);
            if (!a_split) {
                // rows the taps read: [t_lo - 1, t_lo + t_n + 1) (all rows of an untrimmed launch)
                const int r_lo = a.t_n > 0 ? (a.t_lo > 0 ? a.t_lo - 1 : 0) : 0;
                const int r_hi = a.t_n > 0 ? (a.t_lo + a.t_n + 1 < a.Tin ? a.t_lo + a.t_n + 1 : a.Tin) : a.Tin;
                launch_split3(static_cast<const __half*>(A), a.S, a.Tin, a.K, r_lo, r_hi - r_lo,
                              static_cast<__nv_bfloat16*>(h->ws_split), st, smap, a.rpf);
                h->launches++;
                a.a_wrap = 2 * a.K / 64;              // ws_split holds [hi | lo]; K chunks beyond it wrap around to hi
                a.K *= 3;
                Wcols *= 3;
            }
            a.mma_bf16 = 1;
            int rc = act_map(h, &ma, h->ws_split, a.a_wrap * 64, a.Tin, S_buf, a.Tbox, mapped ? 1 : a.Wbox, 0);
            if (rc) return rc;
            rc = weight_map(h, &mw, Wx3, Wrows, Wcols, gemm_tc_block_n(a), 0);
            if (rc) return rc;
            cudaError_t le = launch_gemm_tc(epi, 1, 0, a, ma, *mw, ma, ma, h->sm_count, st);     // fp16 resid / out
            prof_end(h, st);
            CK(h, le);
            return 0;
        }
        int rc = act_map(h, &ma, A, a.K, a.Tin, S_buf, a.Tbox, mapped ? 1 : a.Wbox, hk);
        if (rc) return rc;
        rc = weight_map(h, &mw, Wh[hk], Wrows, Wcols, gemm_tc_block_n(a), hk);
        if (rc) return rc;
        // NoiseBlock of block 0 (the blocks without the fused chain): y in and x out by TMA through per-warp staging tiles
        CUtensorMap my = ma, mo = ma;
        if (epi == EPI_NOISE && !out_f32 && a.Tbox == 128 && a.Wbox == 1 && a.up == 1 && gemm_tc_block_n(a) == 256 &&
            !getenv("SNACB_NO_TMA_EPI")) {
            rc = act_map(h, &my, a.resid, a.Cout, a.Tin, S_buf, 32, 1, hk, 2);
            if (rc) return rc;
            rc = act_map(h, &mo, a.out, a.Cout, a.Tin, S_buf, 32, 1, hk, 2);
            if (rc) return rc;
            a.tma_epi = 1;
        }
        prof_begin(h, pname, st);
        cudaError_t le = launch_gemm_tc(epi, hk, out_f32 ? 1 : 0, a, ma, *mw, my, mo, h->sm_count, st);
        prof_end(h, st);
        CK(h, le);
        return 0;
    };

    // ---- stem 1x1 768 -> 1024, Snake of block 0 applied in the epilogue
    {
        GemmArgs a{};
        a.S = S; a.Tin = T0; a.K = kLatent; a.N = kDecDim; a.Cout = kDecDim; a.ntaps = 1; a.up = 1;
        a.bias = h->stem_pw_b; a.alpha = h->blk[0].alpha; a.inv_alpha = h->blk[0].inv_alpha;
        if (trimmed[0]) { a.t_lo = stem_lo; a.t_n = stem_hi - stem_lo; }
        a.out = cur;
        int rc = stem_live ? gemm("stem_pw", EPI_BIAS_SNAKE, false, a, a0, h->stem_pw_f32, h->stem_pw_h, kDecDim, kLatent, h->stem_pw_x3) : 0;
        if (rc) return rc;
        rc = tap_any("stem", cur, dt_h, (int64_t)S * T0, kDecDim);
        if (rc) return rc;
    }

    int Tin = T0;
    char nm[32];
    for (int bi = 0; bi < 4; ++bi) {
        BlockW& b = h->blk[bi];
        const int T = Tin * b.s;
        void* const oth = plan ? plan->ct[bi] : other(cur);
        // ---- ConvTranspose1d as a 2-tap GEMM per output phase: cur [S*Tin][Cin] -> oth [S*T][Cout]
        if (!plan || live(ct_in[bi])) {
            GemmArgs a{};
            a.S = S; a.Tin = Tin; a.K = b.Cin; a.N = b.s * b.Cout; a.Cout = b.Cout; a.ntaps = 2; a.up = b.s;
            if (trimmed[bi]) { a.t_lo = ct_in[bi].lo; a.t_n = ct_in[bi].hi - ct_in[bi].lo; }
            a.bias = b.ct_b; a.out = oth;
            snprintf(nm, sizeof nm, "b%d.convt", bi);
            int rc = 0;
            if (!f32 && !x3 && !h->no_convt_res && convt_res_supported(b.Cin, b.Cout, b.s) && Tin >= 128) {
                // weights resident in smem, one activation load per tile, row-shifted descriptors per tap
                CUtensorMap ma, mo;
                const CUtensorMap* mw;
                rc = act_map(h, &ma, cur, b.Cin, Tin, S_buf, convt_res_box_rows(), 1, hk);
                if (rc) return rc;
                rc = weight_map(h, &mw, b.ct_h[hk], b.s * b.Cout, 2 * b.Cin, b.s * b.Cout, hk);
                if (rc) return rc;
                rc = act_map(h, &mo, oth, b.Cout, T, S_buf, 128 * b.s, 1, hk);
                if (rc) return rc;
                a.seed = seed; a.stream_offset = stream_offset; a.stream_keys = stream_keys; a.Tbox = 128; a.Wbox = 1;
                a.map = smap; a.rpf = Tin / F;
                prof_begin(h, nm, st);
                cudaError_t le = launch_convt_res(hk, a, ma, *mw, mo, h->sm_count, st);
                prof_end(h, st);
                CK(h, le);
                h->launches++;
            } else if (!f32 && !x3 && !h->no_convt_res && convt_ph_supported(b.Cin, b.Cout, b.s) && Tin >= 128) {
                // one output phase's weights resident per CTA group, one activation load per tile
                CUtensorMap ma, mo;
                const CUtensorMap* mw;
                rc = act_map(h, &ma, cur, b.Cin, Tin, S_buf, convt_res_box_rows(), 1, hk);
                if (rc) return rc;
                rc = weight_map(h, &mw, b.ct_h[hk], b.s * b.Cout, 2 * b.Cin, b.Cout, hk);
                if (rc) return rc;
                rc = act_map(h, &mo, oth, b.s * b.Cout, Tin, S_buf, 32, 1, hk);     // a row = the s phases of an input row
                if (rc) return rc;
                a.seed = seed; a.stream_offset = stream_offset; a.stream_keys = stream_keys; a.Tbox = 128; a.Wbox = 1;
                a.map = smap; a.rpf = Tin / F;
                prof_begin(h, nm, st);
                cudaError_t le;
#ifdef SNACB_EXPERIMENTS
                if (getenv("SNACB_CONVT_2CTA")) {                 // experiment (measured slower): CTA pairs, half of the weights per CTA
                    const CUtensorMap* mw2;
                    rc = weight_map(h, &mw2, b.ct_h[hk], b.s * b.Cout, 2 * b.Cin, b.Cout / 2, hk);
                    if (rc) return rc;
                    le = launch_convt_ph2(hk, a, ma, *mw2, mo, h->sm_count, st);
                } else
#endif
                    le = launch_convt_ph(hk, a, ma, *mw, mo, h->sm_count, st);
                prof_end(h, st);
                CK(h, le);
                h->launches++;
            } else {
                rc = gemm(nm, EPI_BIAS, false, a, cur, b.ct_f32, b.ct_h, b.s * b.Cout, 2 * b.Cin, b.ct_x3);
            }
            if (rc) return rc;
            rc = tap_any(nm, oth, dt_h, (int64_t)S * T, b.Cout);
            if (rc) return rc;
        }
        // ---- fused NoiseBlock + 3 ResidualUnits + next Snake: oth -> cur, one kernel
        const bool unfused = (flags & SNACB_UNFUSED) != 0 || h->no_chain || x3;
        if (!f32 && !xf32 && !unfused && b.chain[hk]) {
            void* const ob = plan ? plan->out[bi] : cur;      // ping-pong: back into the ConvTranspose's input buffer
            if (plan && !live(post[bi])) { cur = ob; Tin = T; continue; }
            ChainArgs ca{};
            ca.S = S; ca.T = T; ca.C = b.Cout; ca.out = ob;
            if (trimmed[bi]) { ca.t_lo = post[bi].lo; ca.t_n = post[bi].hi - post[bi].lo; }
            for (int ri = 0; ri < 3; ++ri) {
                const ResW& r = b.res[ri];
                ca.res[ri] = ChainLayer{r.alpha1, r.inv1, r.dw_w, r.dw_b, r.alpha2, r.inv2};
            }
            ca.bias_cum = b.bias_cum;
            ca.alpha_next = bi < 3 ? h->blk[bi + 1].alpha : h->tail_alpha;
            ca.inv_next = bi < 3 ? h->blk[bi + 1].inv_alpha : h->tail_inv;
            ca.noise = noise ? noise[bi] : nullptr; ca.seed = seed; ca.noise_stage = bi; ca.stream_offset = stream_offset; ca.stream_keys = stream_keys;
            ca.t0 = origin_frames * 4 * (T / T0);      // the buffers hold the stream from frame origin_frames on
            ca.map = smap; ca.rpf = T / F;
            const bool ws = h->chain_ws && chain_ws_supported(b.Cout, hk) && !plan;     // the experiment knows no sessions
            memcpy(ca.spans, ws ? b.spans_ws : b.spans, sizeof ca.spans);
            ca.strip_tiles = 1; ca.sps = 1; ca.last_strip_tiles = 1;
            if (!ws) {
                // strips of tiles walked in order by one CTA (carry-top tiles, kernels_chain.cu); A/B switches for the tests:
                // SNACB_NO_CARRY=1: one halo-top tile per strip; SNACB_NO_SHORT_TILE=1: the last tile runs as a full tile
                const char* nst = getenv("SNACB_NO_SHORT_TILE");
                const char* ncy = getenv("SNACB_NO_CARRY");
                const int slots = h->sm_count * (chain_warps(b.Cout) == 8 ? 2 : 1);
                const int t_n = ca.t_n > 0 ? ca.t_n : T;
                bool no_carry = (ncy && ncy[0] == '1');
                if (!no_carry && !b.spans_carry_ok) no_carry = true;
                if (!no_carry) {
                    const size_t need = chain_carry_bytes_max(h->sm_count);     // the largest over the blocks, once
                    if (grow(h, &h->chain_carry, &h->chain_carry_bytes, need)) no_carry = true;
                }
                int lrows = 0;
                chain_plan_strips(b.Cout, t_n, S, slots, no_carry, &ca.strip_tiles, &ca.sps, &ca.last_strip_tiles, &lrows);
                ca.carry = h->chain_carry;
                memcpy(ca.spans_carry, b.spans_carry, sizeof ca.spans_carry);
                if (lrows > 0 && !(nst && nst[0] == '1')) {
                    const bool last_carry = ca.last_strip_tiles > 1;
                    const int own_end = (last_carry ? 0 : kChainHalo) + lrows;
                    const int key = own_end | (last_carry ? 1 << 16 : 0);
                    if (b.spans_last_key != key) {
                        b.spans_last_ok = chain_build_spans(b.Cout, b.spans_last, own_end, last_carry);
                        b.spans_last_key = key;
                    }
                    if (b.spans_last_ok) { memcpy(ca.spans_last, b.spans_last, sizeof ca.spans_last); ca.last_rows = lrows; }
                }
            }
            ca.tile_counter = h->tile_counter;
            ca.jitter = h->chain_jitter;
            CK(h, cudaMemsetAsync(h->tile_counter, 0, sizeof(int), st));
            const bool fold = hk && b.fold && !h->no_fold;
            CUtensorMap my, moe, mom;
            const CUtensorMap* mn;
            int rc = act_map(h, &my, oth, b.Cout, T, S_buf, 128, 1, hk, 1);
            if (rc) return rc;
            rc = act_map(h, &moe, ob, b.Cout, T, S_buf, 128 - kChainHalo, 1, hk, 1);
            if (rc) return rc;
            rc = act_map(h, &mom, ob, b.Cout, T, S_buf, 128, 1, hk, 1);
            if (rc) return rc;
            rc = weight_map(h, &mn, b.nz_h[hk], b.Cout, b.Cout, b.Cout, hk);
            if (rc) return rc;
            const CUtensorMap tm[7] = {my, moe, mom, *mn, fold ? b.res[0].tm_pwc : b.res[0].tm_pw[hk],
                                       fold ? b.res[1].tm_pwc : b.res[1].tm_pw[hk], fold ? b.res[2].tm_pwc : b.res[2].tm_pw[hk]};
            snprintf(nm, sizeof nm, "b%d.chain", bi);
            prof_begin(h, nm, st);
            if (h->chain_prof) { rc = chain_prof_begin(h, &ca, st); if (rc) return rc; }
            cudaError_t le = ws ? launch_chain_ws(hk, fold ? 1 : 0, ca, tm, h->sm_count, st)
                                : launch_chain(hk, fold ? 1 : 0, ca, tm, h->sm_count, st);
            prof_end(h, st);
            CK(h, le);
            if (h->chain_prof) { rc = ws ? chain_ws_prof_report(h, ca, bi, st) : chain_prof_report(h, ca, bi, st); if (rc) return rc; }
            h->launches++;
            cur = ob;
            snprintf(nm, sizeof nm, "b%d.res2", bi);
            rc = tap_any(nm, cur, dt_h, (int64_t)S * T, b.Cout);
            if (rc) return rc;
            Tin = T;
            continue;
        }
        // ---- NoiseBlock: oth -> cur   x = y + n * (Wn y)   (ping-pong: cur is the ConvTranspose's input buffer)
        if (plan) cur = plan->nz[bi];
        if (!plan || live(plan->nz_r[bi])) {
            GemmArgs a{};
            a.S = S; a.Tin = T; a.K = b.Cout; a.N = b.Cout; a.Cout = b.Cout; a.ntaps = 1; a.up = 1;
            if (plan) { a.t_lo = plan->nz_r[bi].lo; a.t_n = plan->nz_r[bi].hi - plan->nz_r[bi].lo; }
            else if (trimmed[bi]) { a.t_lo = post[bi].lo; a.t_n = post[bi].hi - post[bi].lo; }
            a.noise = noise ? noise[bi] : nullptr; a.noise_stage = bi;
            a.t0 = origin_frames * 4 * (T / T0);
            a.resid = oth; a.out = cur;
            snprintf(nm, sizeof nm, "b%d.noise", bi);
            int rc = gemm(nm, EPI_NOISE, xf32, a, oth, b.nz_f32, b.nz_h, b.Cout, b.Cout, b.nz_x3);
            if (rc) return rc;
            rc = tap_any(nm, cur, dt_x, (int64_t)S * T, b.Cout);
            if (rc) return rc;
        }
        // ---- three ResidualUnits, dilations 1, 3, 9: cur -> oth -> cur -> oth
        static const int dils[3] = {1, 3, 9};
        for (int ri = 0; ri < 3; ++ri) {
            ResW& r = b.res[ri];
            const bool last = (ri == 2);
            const float* an = last ? (bi < 3 ? h->blk[bi + 1].alpha : h->tail_alpha) : nullptr;
            const float* ian = last ? (bi < 3 ? h->blk[bi + 1].inv_alpha : h->tail_inv) : nullptr;
            void* const ro = plan ? plan->res[bi][ri] : other(cur);
            if (plan && !live(plan->res_r[bi][ri])) { cur = ro; continue; }
            ResUnitArgs ra{};
            ra.S = S; ra.T = T; ra.C = b.Cout; ra.dil = dils[ri];
            if (plan) { ra.t_lo = plan->res_r[bi][ri].lo; ra.t_n = plan->res_r[bi][ri].hi - plan->res_r[bi][ri].lo; }
            else if (trimmed[bi]) { ra.t_lo = post[bi].lo; ra.t_n = post[bi].hi - post[bi].lo; }
            ra.x = cur; ra.out = ro;
            ra.map = smap; ra.rpf = T / F;
            ra.alpha1 = r.alpha1; ra.inv_alpha1 = r.inv1; ra.dw_w = r.dw_w; ra.dw_b = r.dw_b;
            ra.alpha2 = r.alpha2; ra.inv_alpha2 = r.inv2; ra.pw_b = r.pw_b;
            ra.alpha_next = an; ra.inv_alpha_next = ian;
            snprintf(nm, sizeof nm, "b%d.res%d", bi, ri);
            prof_begin(h, nm, st);
            if (f32) {
                launch_respre_f32(ra, P, st);
                GemmArgs a{};
                a.S = S; a.Tin = T; a.K = b.Cout; a.N = b.Cout; a.Cout = b.Cout; a.ntaps = 1; a.up = 1;
                a.bias = r.pw_b; a.alpha = an; a.inv_alpha = ian; a.resid = cur; a.out = ro;
                tile_boxes(a.Tin, &a.Tbox, &a.Wbox);
                launch_gemm_f32(last ? EPI_RES_SNAKE : EPI_RES, a, P, r.pw_f32, st);
                prof_end(h, st);
                h->launches += 2;
                CK(h, cudaGetLastError());
            } else if (x3) {
                // front half on the CUDA cores (fp32 math on the fp16 stream) -> [a | a] in bf16, then the 1x1 conv as a
                // bf16 GEMM against [W_hi | W_lo] with the residual add (+ the next Snake) in its epilogue
                launch_respre16(ra, ra.t_n > 0 ? ra.t_lo : 0, ra.t_n > 0 ? ra.t_n : T, static_cast<__nv_bfloat16*>(h->ws_split), st);
                prof_end(h, st);
                h->launches++;
                CK(h, cudaGetLastError());
                GemmArgs a{};
                a.S = S; a.Tin = T; a.K = 2 * b.Cout; a.N = b.Cout; a.Cout = b.Cout; a.ntaps = 1; a.up = 1;
                a.a_wrap = b.Cout / 64;               // ws_split holds [a]; the second K half reads it again (against W_lo)
                a.t_lo = ra.t_lo; a.t_n = ra.t_n;
                a.bias = r.pw_b; a.alpha = an; a.inv_alpha = ian; a.resid = cur; a.out = ro;
                snprintf(nm, sizeof nm, "b%d.res%d", bi, ri);
                int rc2 = gemm(nm, last ? EPI_RES_SNAKE : EPI_RES, false, a, nullptr, nullptr, nullptr, b.Cout, 2 * b.Cout, r.pw_x3, true);
                if (rc2) return rc2;
            } else if (!xf32 && !h->res_v1) {
                int tile_m, box_rows;
                resunit2_geometry(ra.C, ra.dil, &tile_m, &box_rows);
                CUtensorMap mx;
                int rc2 = act_map(h, &mx, cur, ra.C, T, S_buf, box_rows, 1, hk, resunit2_swizzled_x(ra.C) ? 1 : 0);
                if (rc2) return rc2;
                const bool rprof = getenv("SNACB_RES_PROF") != nullptr;     // debug: in-kernel wait / phase cycles of CTA 0
                if (rprof) {
                    CK(h, cudaMalloc(reinterpret_cast<void**>(&ra.prof), 8 * sizeof(long long)));
                    CK(h, cudaMemsetAsync(ra.prof, 0, 8 * sizeof(long long), st));
                }
                CUtensorMap mo = mx;                                          // C = 512: staged TMA-store epilogue
                if (resunit2_swizzled_x(ra.C)) {
                    rc2 = act_map(h, &mo, ro, ra.C, T, S_buf, 32, 1, hk, 2);
                    if (rc2) return rc2;
                }
                cudaError_t le = launch_resunit2(hk, ra, mx, r.tm_pw[hk], mo, h->sm_count, st);
                prof_end(h, st);
                CK(h, le);
                h->launches++;
                if (rprof) {
                    long long pv[8];
                    CK(h, cudaStreamSynchronize(st));
                    CK(h, cudaMemcpy(pv, ra.prof, sizeof pv, cudaMemcpyDeviceToHost));
                    CK(h, cudaFree(ra.prof));
                    fprintf(stderr, "%s C=%d: CTA0 cycles: mma-warp waits a_full %lld w_full %lld acc_empty %lld | warp0 waits x_full %lld "
                            "a_empty %lld, prologue %lld, epilogue %lld over %lld tiles\n", nm, ra.C, pv[0], pv[1], pv[2], pv[3], pv[4],
                            pv[5], pv[6], pv[7]);
                }
            } else {
                cudaError_t le = launch_resunit_tc(last ? EPI_RES_SNAKE : EPI_RES, hk, xf32 ? 1 : 0, ra, r.tm_pw[hk], st);
                prof_end(h, st);
                CK(h, le);
                h->launches++;
            }
            int rc = tap_any(nm, ro, last ? dt_h : dt_x, (int64_t)S * T, b.Cout);
            if (rc) return rc;
            cur = ro;
        }
        Tin = T;
    }

    // ---- tail
    const int T = Tin;                                   // 2048 * F samples
    const int t_begin = out_lo, n_out = out_hi - out_lo;
    if (plan && n_out <= 0) { h->streams += S; return 0; }
    CUtensorMap mt128, mt8;                              // 16-bit activations: the tensor-core tail reads its rows by TMA
    if (!f32) {
        int rc = act_map(h, &mt128, cur, 64, T, S_buf, 128, 1, hk, 1);
        if (rc) return rc;
        rc = act_map(h, &mt8, cur, 64, T, S_buf, 8, 1, hk, 1);
        if (rc) return rc;
    }
    prof_begin(h, "tail", st);
    if (f32) launch_tail<float>(static_cast<const float*>(cur), S, T, t_begin, n_out, h->tail_w, h->tail_b, pcm, wave, st, smap);
    else if (hk) launch_tail<__half>(static_cast<const __half*>(cur), S, T, t_begin, n_out, h->tail_w, h->tail_b, pcm, wave, st, smap, &mt128, &mt8);
    else launch_tail<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(cur), S, T, t_begin, n_out, h->tail_w, h->tail_b, pcm, wave, st, smap, &mt128, &mt8);
    prof_end(h, st);
    h->launches++;
    CK(h, cudaGetLastError());
    h->streams += S;
    return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int snacb_version(void) { return SNACB_VERSION; }

int snacb_samples_out(int frames, int flags) {
    const int T = 2048 * frames;
    return ((flags & SNACB_EXTRACT_SLICE) && T > 4096) ? 2048 : T;
}

const char* snacb_last_error(snacb_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int snacb_create(snacb_handle* out, const snacb_weights* w, int device) {
    if (!out || !w) return fail(nullptr, SNACB_ERR_ARG, "snacb_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, SNACB_ERR_NO_GPU, "snacb_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, SNACB_ERR_ARG, "snacb_create: bad device %d", device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, SNACB_ERR_NO_GPU, "snacb_create: device %d is sm_%d%d, this build is sm_100a only", device,
                    prop.major, prop.minor);
    snacb_handle h = new (std::nothrow) snacb_handle_s();
    if (!h) return fail(nullptr, SNACB_ERR_NOMEM, "snacb_create: out of host memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    auto bail = [&](int rc) { g_create_error = h->err; snacb_destroy(h); return rc; };
#define CKH(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        fail(h, SNACB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(SNACB_ERR_CUDA); } } while (0)
#define RC(call) do { int rc_ = (call); if (rc_) return bail(rc_); } while (0)
    CKH(cudaSetDevice(device));
    CKH(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CKH(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) {
            fail(h, SNACB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
            return bail(SNACB_ERR_CUDA);
        }
        h->encode = reinterpret_cast<PFN_tmapEncodeTiled>(fn);
    }
    // ---- VQ + stem
    for (int i = 0; i < 3; ++i) {
        float* p;
        RC(upload_f32(h, &p, vec(w->codebook[i], (size_t)kCodebook * kCodeDim))); h->vq.codebook[i] = p;
        RC(upload_f32(h, &p, vec(w->out_proj_w[i], (size_t)kLatent * kCodeDim))); h->vq.out_w[i] = p;
        RC(upload_f32(h, &p, vec(w->out_proj_b[i], kLatent))); h->vq.out_b[i] = p;
    }
    {
        float* p;
        RC(upload_f32(h, &p, dw_transpose(w->stem_dw_w, kLatent))); h->vq.dw_w = p;
        RC(upload_f32(h, &p, vec(w->stem_dw_b, kLatent))); h->vq.dw_b = p;
        auto pw = vec(w->stem_pw_w, (size_t)kDecDim * kLatent);
        RC(upload_f32(h, &h->stem_pw_f32, pw));
        RC(upload_h16(h, h->stem_pw_h, pw));
        RC(upload_x3(h, &h->stem_pw_x3, pw, kDecDim, 1, kLatent, 3));
        RC(upload_f32(h, &h->stem_pw_b, vec(w->stem_pw_b, kDecDim)));
    }
    // ---- decoder blocks
    static const int strides[4] = {8, 8, 4, 2};
    int cin = kDecDim;
    for (int bi = 0; bi < 4; ++bi) {
        const snacb_block_weights& s = w->block[bi];
        BlockW& b = h->blk[bi];
        b.Cin = cin; b.Cout = cin / 2; b.s = strides[bi];
        RC(upload_f32(h, &b.alpha, vec(s.alpha, cin)));
        RC(upload_f32(h, &b.inv_alpha, inv_alpha(s.alpha, cin)));
        auto ct = pack_convt(s.convt_w, b.Cin, b.Cout, b.s);
        RC(upload_f32(h, &b.ct_f32, ct));
        RC(upload_h16(h, b.ct_h, ct));
        RC(upload_x3(h, &b.ct_x3, ct, static_cast<size_t>(b.s) * b.Cout, 2, b.Cin, 3));
        RC(upload_f32(h, &b.ct_b, vec(s.convt_b, b.Cout)));
        auto nz = vec(s.noise_w, (size_t)b.Cout * b.Cout);
        RC(upload_f32(h, &b.nz_f32, nz));
        RC(upload_h16(h, b.nz_h, nz));
        RC(upload_x3(h, &b.nz_x3, nz, b.Cout, 1, b.Cout, 3));
        for (int ri = 0; ri < 3; ++ri) {
            const snacb_resunit_weights& rs = s.res[ri];
            ResW& r = b.res[ri];
            const int C = b.Cout;
            RC(upload_f32(h, &r.alpha1, vec(rs.alpha1, C)));
            RC(upload_f32(h, &r.inv1, inv_alpha(rs.alpha1, C)));
            RC(upload_f32(h, &r.dw_w, dw_transpose(rs.dw_w, C)));
            RC(upload_f32(h, &r.dw_b, vec(rs.dw_b, C)));
            RC(upload_f32(h, &r.alpha2, vec(rs.alpha2, C)));
            RC(upload_f32(h, &r.inv2, inv_alpha(rs.alpha2, C)));
            auto pw = vec(rs.pw_w, (size_t)C * C);
            RC(upload_f32(h, &r.pw_f32, pw));
            RC(upload_h16(h, r.pw_h, pw));
            RC(upload_x3(h, &r.pw_x3, pw, C, 1, C, 2));
            RC(upload_f32(h, &r.pw_b, vec(rs.pw_b, C)));
            for (int k = 0; k < 2; ++k) RC(make_tmap_2d(h, &r.tm_pw[k], r.pw_h[k], C, C, C > 256 ? 256 : C, k));
        }
        {
            std::vector<float> bc(static_cast<size_t>(3) * b.Cout);
            for (int c = 0; c < b.Cout; ++c) {
                float acc = 0.f;
                for (int ri = 0; ri < 3; ++ri) { acc += s.res[ri].pw_b[c]; bc[static_cast<size_t>(ri) * b.Cout + c] = acc; }
            }
            RC(upload_f32(h, &b.bias_cum, bc));
            b.chain[0] = chain_supported(b.Cout, 0);
            b.chain[1] = chain_supported(b.Cout, 1);
            if (b.chain[0] || b.chain[1]) {
                chain_build_spans(b.Cout, b.spans);
                b.spans_carry_ok = chain_build_spans(b.Cout, b.spans_carry, 0, true);
            }
            if (chain_ws_supported(b.Cout, 1)) chain_ws_build_spans(b.Cout, b.spans_ws);
            // fp16 chain: the alpha-folded formulation (kernels_chain.cu) divides by the Snake alphas when it packs its
            // parameters; it is used only where that is numerically safe for THIS checkpoint, else the general variant runs
            b.fold = b.chain[1] && chain_fold_safe(s, b.Cout);
            if (b.fold) {
                const int C = b.Cout;
                for (int ri = 0; ri < 3; ++ri) {
                    // snake2(a) = (a'' + sin^2 a'') / alpha2 with a'' = alpha2 a: the 1 / alpha2 goes into W's K columns
                    std::vector<__half> wc(static_cast<size_t>(C) * C);
                    for (int n = 0; n < C; ++n)
                        for (int k = 0; k < C; ++k)
                            wc[static_cast<size_t>(n) * C + k] =
                                __float2half_rn(s.res[ri].pw_w[static_cast<size_t>(n) * C + k] / s.res[ri].alpha2[k]);
                    __half* d;
                    RC(dev_alloc(h, &d, wc.size()));
                    CKH(cudaMemcpy(d, wc.data(), wc.size() * 2, cudaMemcpyHostToDevice));
                    b.res[ri].pwc_h = d;
                    RC(make_tmap_2d(h, &b.res[ri].tm_pwc, d, C, C, C, 1));
                }
            }
        }
        cin = b.Cout;
    }
    RC(upload_f32(h, &h->tail_alpha, vec(w->tail_alpha, 64)));
    RC(upload_f32(h, &h->tail_inv, inv_alpha(w->tail_alpha, 64)));
    {
        std::vector<float> tw(7 * 64);
        for (int c = 0; c < 64; ++c)
            for (int j = 0; j < 7; ++j) tw[j * 64 + c] = w->tail_w[c * 7 + j];
        RC(upload_f32(h, &h->tail_w, tw));
        h->tail_b = w->tail_b[0];
    }
    if (const char* e = getenv("SNACB_RES_V1")) h->res_v1 = atoi(e) != 0;
    if (const char* e = getenv("SNACB_NO_CHAIN")) h->no_chain = atoi(e) != 0;
    if (const char* e = getenv("SNACB_BF16_PLAIN")) h->bf16_plain = atoi(e) != 0;
    if (const char* e = getenv("SNACB_NO_FOLD")) h->no_fold = atoi(e) != 0;
    if (const char* e = getenv("SNACB_CHAIN_WS")) h->chain_ws = atoi(e) != 0 && chain_ws_built();
    if (const char* e = getenv("SNACB_CHAIN_JITTER")) h->chain_jitter = static_cast<unsigned>(strtoul(e, nullptr, 10));
    if (const char* e = getenv("SNACB_TMAP_CACHE")) { const long n = atol(e); if (n >= 1) h->max_act_maps = static_cast<size_t>(n); }
    if (const char* e = getenv("SNACB_CHAIN_PROF")) h->chain_prof = atoi(e);
    if (const char* e = getenv("SNACB_NO_TRIM")) h->no_trim = atoi(e) != 0;
    if (const char* e = getenv("SNACB_NO_CONVT_RES")) h->no_convt_res = atoi(e) != 0;
    if (const char* e = getenv("SNACB_GROUP_MB")) {
        long mb = atol(e);
        if (mb > 0) h->group_bytes = static_cast<size_t>(mb) << 20;
    }
    RC(dev_alloc(h, &h->tile_counter, 1));
    CKH(cudaDeviceSynchronize());
#undef CKH
#undef RC
    *out = h;
    return SNACB_OK;
}

void snacb_destroy(snacb_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    clear_taps(h);
    for (auto& r : h->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (void* p : h->allocs) cudaFree(p);
    for (int i = 0; i < 3; ++i) if (h->ws_buf[i]) cudaFree(h->ws_buf[i]);
    if (h->ws_a0) cudaFree(h->ws_a0);
    if (h->ws_split) cudaFree(h->ws_split);
    if (h->chain_carry) cudaFree(h->chain_carry);
    if (h->ws_codes) cudaFree(h->ws_codes);
    if (h->st_tok) cudaFree(h->st_tok);
    if (h->st_pcm) cudaFree(h->st_pcm);
    for (int i = 0; i < 2; ++i) {
        if (h->pl_tok[i]) cudaFree(h->pl_tok[i]);
        if (h->pl_pcm[i]) cudaFree(h->pl_pcm[i]);
        if (h->pl_done[i]) cudaEventDestroy(h->pl_done[i]);
        if (h->pl_copied[i]) cudaEventDestroy(h->pl_copied[i]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int snacb_set_group_bytes(snacb_handle h, size_t bytes) {
    if (!h) return SNACB_ERR_ARG;
    if (bytes) h->group_bytes = bytes;
    return SNACB_OK;
}

int snacb_stats(snacb_handle h, uint64_t* kernel_launches, uint64_t* streams_decoded) {
    if (!h) return SNACB_ERR_ARG;
    if (kernel_launches) *kernel_launches = h->launches;
    if (streams_decoded) *streams_decoded = h->streams;
    return SNACB_OK;
}

int snacb_chain_modes(snacb_handle h, int32_t* modes) {
    if (!h || !modes) return SNACB_ERR_ARG;
    for (int bi = 0; bi < 4; ++bi) {
        const BlockW& b = h->blk[bi];
        modes[bi] = !b.chain[1] ? 0 : ((b.fold && !h->no_fold) ? 2 : 1);
    }
    return SNACB_OK;
}

int snacb_unpack(snacb_handle h, const int32_t* tok, int B, int ntok, int flags, int32_t* c0, int32_t* c1, int32_t* c2,
                 void* stream) {
    if (!h) return SNACB_ERR_ARG;
    if (B < 0 || ntok < 0) return fail(h, SNACB_ERR_ARG, "snacb_unpack: negative size");
    const int F = ntok / kFrame;
    if (B == 0 || F == 0) return SNACB_OK;             // fewer than 7 codes: nothing to do (helper returns None)
    if (!tok || !c0 || !c1 || !c2) return fail(h, SNACB_ERR_ARG, "snacb_unpack: null pointer");
    CK(h, cudaSetDevice(h->device));
    launch_unpack(tok, B, ntok, F, (flags & SNACB_RAW_IDS) ? 1 : 0, c0, c1, c2, static_cast<cudaStream_t>(stream));
    h->launches++;
    CK(h, cudaGetLastError());
    return SNACB_OK;
}

int snacb_decode(snacb_handle h, const int32_t* tok, int B, int tok_stride, int frames, int flags,
                 const float* const* noise, uint64_t seed, int16_t* pcm, float* wave, void* stream) {
    return snacb_decode_keyed(h, tok, B, tok_stride, frames, flags, noise, seed, nullptr, pcm, wave, stream);
}

int snacb_decode_keyed(snacb_handle h, const int32_t* tok, int B, int tok_stride, int frames, int flags,
                       const float* const* noise, uint64_t seed, const int32_t* stream_keys, int16_t* pcm, float* wave,
                       void* stream) {
    return snacb_decode_range(h, tok, B, tok_stride, frames, flags, noise, seed, stream_keys, 0, 0, pcm, wave, stream);
}

}  // extern "C"

namespace {
// snacb_decode_range on tokens that hold a stream from frame `origin_frames` on: rows and samples are relative to that
// frame, the NoiseBlock noise stays keyed by the absolute time step (a streaming session's end-of-stream flush)
int decode_range_impl(snacb_handle h, const int32_t* tok, int B, int tok_stride, int frames, int flags,
                      const float* const* noise, uint64_t seed, const int32_t* stream_keys, int sample_lo, int sample_hi,
                      int16_t* pcm, float* wave, void* stream, int origin_frames) {
    if (!h) return SNACB_ERR_ARG;
    const bool ranged = sample_lo != 0 || sample_hi != 0;
    if (ranged && (sample_lo < 0 || sample_hi <= sample_lo || sample_hi > 2048LL * frames))
        return fail(h, SNACB_ERR_ARG, "snacb_decode_range: bad sample range [%d, %d) for %d frames", sample_lo, sample_hi, frames);
    if (B < 0 || frames < 0 || tok_stride < frames * kFrame)
        return fail(h, SNACB_ERR_ARG, "snacb_decode: bad sizes B=%d frames=%d tok_stride=%d", B, frames, tok_stride);
    if (B == 0 || frames == 0) return SNACB_OK;
    if (!tok || !pcm) return fail(h, SNACB_ERR_ARG, "snacb_decode: null tok/pcm");
    if (frames > 16384) return fail(h, SNACB_ERR_ARG, "snacb_decode: frames=%d too large", frames);
    CK(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool f32 = (flags & SNACB_FP32) != 0;
    const bool xf32 = f32 || (flags & SNACB_STREAM_FP32);
    const size_t per_stream_elems = static_cast<size_t>(131072) * frames;      // max T*C over the blocks
    const size_t esz = xf32 ? 4 : 2;
    size_t G = h->group_bytes / (per_stream_elems * esz);
    if (G < 1) G = 1;
    if (G > static_cast<size_t>(B)) G = B;
    // workspace
    {
        const size_t need = G * per_stream_elems * esz;
        if (need > h->ws_buf_bytes) {
            h->amaps.clear();
            for (int i = 0; i < 3; ++i) {
                if (h->ws_buf[i]) { CK(h, cudaFree(h->ws_buf[i])); h->ws_buf[i] = nullptr; }
            }
            h->ws_buf_bytes = 0;
            for (int i = 0; i < 3; ++i) CK(h, cudaMalloc(&h->ws_buf[i], need));
            h->ws_buf_bytes = need;
        }
        int rc = grow(h, &h->ws_a0, &h->ws_a0_bytes, G * 4 * frames * kLatent * (f32 ? 4 : 2));
        if (rc) return rc;
        if ((flags & SNACB_BF16) && !f32 && !h->bf16_plain) {             // bf16x3: the split operand of the widest GEMM input
            rc = grow(h, &h->ws_split, &h->ws_split_bytes, G * per_stream_elems * 2 * 2);
            if (rc) return rc;
        }
        size_t cbytes = h->ws_codes_elems * sizeof(int32_t);
        rc = grow(h, reinterpret_cast<void**>(&h->ws_codes), &cbytes, G * 7 * frames * sizeof(int32_t));
        if (rc) return rc;
        h->ws_codes_elems = cbytes / sizeof(int32_t);
    }
    const int n_out = ranged ? sample_hi - sample_lo : snacb_samples_out(frames, flags);
    const int T0 = 4 * frames;
    const size_t nlen[4] = {(size_t)T0 * 8, (size_t)T0 * 64, (size_t)T0 * 256, (size_t)T0 * 512};
    for (int g0 = 0; g0 < B; g0 += static_cast<int>(G)) {
        const int S = (B - g0) < static_cast<int>(G) ? (B - g0) : static_cast<int>(G);
        const float* nz[4];
        if (noise) for (int i = 0; i < 4; ++i) nz[i] = noise[i] + static_cast<size_t>(g0) * nlen[i];
        int rc = run_group(h, tok + static_cast<size_t>(g0) * tok_stride, S, tok_stride, frames, flags,
                           noise ? nz : nullptr, seed, g0, stream_keys ? stream_keys + g0 : nullptr,
                           ranged ? sample_lo : 0, ranged ? sample_hi : 0, pcm + static_cast<size_t>(g0) * n_out,
                           wave ? wave + static_cast<size_t>(g0) * n_out : nullptr, st, nullptr, origin_frames);
        if (rc) return rc;
    }
    return SNACB_OK;
}
}  // namespace

extern "C" {

int snacb_decode_range(snacb_handle h, const int32_t* tok, int B, int tok_stride, int frames, int flags,
                       const float* const* noise, uint64_t seed, const int32_t* stream_keys, int sample_lo, int sample_hi,
                       int16_t* pcm, float* wave, void* stream) {
    return decode_range_impl(h, tok, B, tok_stride, frames, flags, noise, seed, stream_keys, sample_lo, sample_hi, pcm, wave,
                             stream, 0);
}

// =================================================================================================
// Stateful streaming session: per-slot, per-stage activations persist in HBM, a step
// with k new frames computes only the rows that became final with them.
// =================================================================================================
}  // extern "C"

struct snacb_session_s {
    snacb_handle h = nullptr;
    int n_slots = 0, max_frames = 0, flags = 0;   // max_frames: frames of every slot's window (buffers)
    std::vector<long long> frames, emitted;   // per slot: frames ingested, samples emitted (whole stream)
    std::vector<long long> origin;            // per slot: stream frame held at row 0 of the slot's buffers
    std::vector<char> finished;
    int32_t* tok = nullptr;                   // [n_slots][7 * max_frames], the window's tokens
    int32_t* codes[3] = {nullptr, nullptr, nullptr};
    int32_t* slot_keys = nullptr;             // [n_slots] = 0, 1, ...: the default NoiseBlock noise key of a slot
    int32_t* map_dev = nullptr;               // [7][n_slots]: slot, off, org, pos, keys, slide list, slide amount of a multi step
    void* a0 = nullptr;
    void* stem = nullptr;
    void* ct[4] = {}; void* nz[4] = {}; void* res[4][3] = {}; void* out[4] = {};
    std::vector<void*> allocs;
    size_t bytes = 0;
};

namespace {

// Rows of every stage that are FINAL once `F` frames of a stream are known (the stream may still grow): a row is final
// when its whole receptive field lies inside the known latent steps.  Forward propagation through the decoder:
//   stem depthwise k7:  latent step t needs z_q[t + 3]                          -> [0, 4F - 3)
//   ConvTranspose (k = 2s, stride s, pad s/2): out row m*s + p reads input rows {m - 1, m} for p < s/2 and {m, m + 1}
//                       for p >= s/2 (k_gemm_tc base_shift); the kernels work on whole input rows m, so a step
//                       processes the rows whose successor is final as well     -> [0, (vin - 1) s)
//                       (k new input rows = k rows of work: whole 128-row tiles, nothing recomputed)
//   NoiseBlock 1x1: pointwise;  ResidualUnit d: +-3d rows;  fused chain: +-39 rows;  tail conv k7: +-3 samples.
struct Frontier {
    int stem, y[4], nz[4], res[4][3], out[4], emit;
};
Frontier frontier_calc(int F, const int* strides, int chain_mask) {
    Frontier f{};
    auto pos = [](int v) { return v > 0 ? v : 0; };
    int vin = pos(4 * F - 3);
    f.stem = vin;
    for (int bi = 0; bi < 4; ++bi) {
        const int s = strides[bi];
        f.y[bi] = vin > 0 ? (vin - 1) * s : 0;
        if (chain_mask >> bi & 1) {
            f.nz[bi] = f.res[bi][0] = f.res[bi][1] = 0;
            f.out[bi] = f.res[bi][2] = pos(f.y[bi] - 39);
        } else {
            f.nz[bi] = f.y[bi];
            f.res[bi][0] = pos(f.nz[bi] - 3);
            f.res[bi][1] = pos(f.res[bi][0] - 9);
            f.res[bi][2] = pos(f.res[bi][1] - 27);
            f.out[bi] = f.res[bi][2];
        }
        vin = f.out[bi];
    }
    f.emit = pos(f.out[3] - 3);
    return f;
}
// does a session with these flags run block bi through the fused chain kernel (else: one kernel per layer)
bool block_fused(snacb_handle h, int bi, int flags) {
    if ((flags & SNACB_BF16) && !h->bf16_plain) return false;           // bf16x3 path
    return h->blk[bi].chain[(flags & SNACB_BF16) ? 0 : 1] && !h->no_chain;
}
Frontier frontier_of(snacb_handle h, int F, int flags) {
    int strides[4], mask = 0;
    for (int bi = 0; bi < 4; ++bi) {
        strides[bi] = h->blk[bi].s;
        if (block_fused(h, bi, flags)) mask |= 1 << bi;
    }
    return frontier_calc(F, strides, mask);
}

int sess_alloc(snacb_session s, void** p, size_t bytes) {
    snacb_handle h = s->h;
    if (cudaMalloc(p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(h, SNACB_ERR_NOMEM, "snacb_session_create: cudaMalloc of %zu bytes failed", bytes);
    }
    s->allocs.push_back(*p);
    s->bytes += bytes;
    return 0;
}

}  // namespace

extern "C" {

int snacb_debug_session_frontier(int frames, int chain_mask, int32_t* out, int cap) {
    if (frames < 0 || !out || cap < 22) return SNACB_ERR_ARG;
    static const int strides[4] = {8, 8, 4, 2};
    const Frontier f = frontier_calc(frames, strides, chain_mask);
    int k = 0;
    out[k++] = f.stem;
    for (int bi = 0; bi < 4; ++bi) {
        out[k++] = f.y[bi]; out[k++] = f.nz[bi];
        for (int ri = 0; ri < 3; ++ri) out[k++] = f.res[bi][ri];
    }
    out[k++] = f.emit;
    return k;
}

int snacb_session_create(snacb_handle h, int n_slots, int max_frames, int flags, snacb_session* out) {
    if (!h || !out) return SNACB_ERR_ARG;
    *out = nullptr;
    if (n_slots < 1 || max_frames < 1 || max_frames > 16384)
        return fail(h, SNACB_ERR_ARG, "snacb_session_create: bad sizes n_slots=%d max_frames=%d", n_slots, max_frames);
    if (flags & (SNACB_FP32 | SNACB_STREAM_FP32 | SNACB_KEEP_TAPS | SNACB_EXTRACT_SLICE | SNACB_UNFUSED))
        return fail(h, SNACB_ERR_ARG, "snacb_session_create: only SNACB_RAW_IDS and SNACB_BF16 apply to a session");
    CK(h, cudaSetDevice(h->device));
    snacb_session s = new (std::nothrow) snacb_session_s();
    if (!s) return SNACB_ERR_NOMEM;
    s->h = h; s->n_slots = n_slots; s->flags = flags;
    s->max_frames = (max_frames + 31) / 32 * 32;              // every stage then has >= 128 rows per slot (whole tiles)
    s->frames.assign(n_slots, 0); s->emitted.assign(n_slots, 0); s->origin.assign(n_slots, 0); s->finished.assign(n_slots, 0);
    const size_t N = static_cast<size_t>(n_slots), Fm = static_cast<size_t>(s->max_frames), T0 = 4 * Fm;
    int rc = sess_alloc(s, reinterpret_cast<void**>(&s->tok), N * 7 * Fm * sizeof(int32_t));
    for (int l = 0; l < 3 && !rc; ++l) rc = sess_alloc(s, reinterpret_cast<void**>(&s->codes[l]), N * (Fm << l) * sizeof(int32_t));
    if (!rc) rc = sess_alloc(s, &s->a0, N * T0 * kLatent * 2);
    if (!rc) rc = sess_alloc(s, &s->stem, N * T0 * kDecDim * 2);
    size_t T = T0;
    for (int bi = 0; bi < 4 && !rc; ++bi) {
        const BlockW& b = h->blk[bi];
        T *= b.s;
        const size_t sz = N * T * b.Cout * 2;
        rc = sess_alloc(s, &s->ct[bi], sz);
        if (block_fused(h, bi, flags)) {
            if (!rc) rc = sess_alloc(s, &s->out[bi], sz);
        } else {
            if (!rc) rc = sess_alloc(s, &s->nz[bi], sz);
            for (int ri = 0; ri < 3 && !rc; ++ri) rc = sess_alloc(s, &s->res[bi][ri], sz);
            s->out[bi] = s->res[bi][2];
        }
    }
    if (rc) { snacb_session_destroy(s); return rc; }
    if (!rc) rc = sess_alloc(s, reinterpret_cast<void**>(&s->slot_keys), N * sizeof(int32_t));
    if (!rc) rc = sess_alloc(s, reinterpret_cast<void**>(&s->map_dev), 7 * N * sizeof(int32_t));
    if (rc) { snacb_session_destroy(s); return rc; }
    std::vector<int32_t> iota(N);
    for (size_t i = 0; i < N; ++i) iota[i] = static_cast<int32_t>(i);
    if (cudaMemset(s->tok, 0, N * 7 * Fm * sizeof(int32_t)) != cudaSuccess ||
        cudaMemcpy(s->slot_keys, iota.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
        snacb_session_destroy(s);
        return fail(h, SNACB_ERR_CUDA, "snacb_session_create: initialising the session buffers failed");
    }
    *out = s;
    return SNACB_OK;
}

void snacb_session_destroy(snacb_session s) {
    if (!s) return;
    cudaSetDevice(s->h->device);
    cudaDeviceSynchronize();
    for (void* p : s->allocs) cudaFree(p);
    s->h->amaps.clear();                                      // maps of the freed buffers
    delete s;
}

int64_t snacb_session_bytes(snacb_session s) { return s ? static_cast<int64_t>(s->bytes) : SNACB_ERR_ARG; }
int snacb_session_max_frames(snacb_session s) { return s ? s->max_frames : SNACB_ERR_ARG; }

int snacb_session_reset(snacb_session s, int slot0, int n) {
    if (!s) return SNACB_ERR_ARG;
    if (slot0 < 0 || n < 0 || slot0 + n > s->n_slots) return fail(s->h, SNACB_ERR_ARG, "snacb_session_reset: bad slot range");
    for (int i = slot0; i < slot0 + n; ++i) { s->frames[i] = 0; s->emitted[i] = 0; s->origin[i] = 0; s->finished[i] = 0; }
    return SNACB_OK;
}

int64_t snacb_session_frames(snacb_session s, int slot) {
    return (s && slot >= 0 && slot < s->n_slots) ? s->frames[slot] : SNACB_ERR_ARG;
}
int64_t snacb_session_emitted(snacb_session s, int slot) {
    return (s && slot >= 0 && slot < s->n_slots) ? s->emitted[slot] : SNACB_ERR_ARG;
}

int snacb_session_next_emit(snacb_session s, int slot, int new_frames, int final) {
    if (!s || slot < 0 || slot >= s->n_slots || new_frames < 0) return SNACB_ERR_ARG;
    // relative to the window's origin (the frontier is translation invariant once it is past the stream's first rows)
    const long long o = s->origin[slot];
    const long long Fl = s->frames[slot] - o + new_frames, El = s->emitted[slot] - 2048 * o;
    if (Fl > 2LL * s->max_frames) return SNACB_ERR_ARG;
    const long long end = final ? 2048 * Fl : frontier_of(s->h, static_cast<int>(Fl), s->flags).emit;
    return end > El ? static_cast<int>(end - El) : 0;
}

int snacb_session_step(snacb_session s, int slot0, int n, const int32_t* new_tok, int tok_stride, int new_frames, int final,
                       uint64_t seed, const int32_t* stream_keys, int16_t* pcm, int pcm_stride, int* n_emitted,
                       void* stream) {
    if (!s) return SNACB_ERR_ARG;
    snacb_handle h = s->h;
    if (n_emitted) *n_emitted = 0;
    if (slot0 < 0 || n < 0 || slot0 + n > s->n_slots || new_frames < 0 || (new_frames > 0 && tok_stride < 7 * new_frames))
        return fail(h, SNACB_ERR_ARG, "snacb_session_step: bad sizes slot0=%d n=%d new_frames=%d tok_stride=%d", slot0, n, new_frames, tok_stride);
    if (n == 0) return SNACB_OK;
    constexpr int kKeep = 8;                   // frames a window keeps when it slides: covers the deepest stage's lag (2.5
                                               // frames), every stage's halo and the left context of the end-of-stream flush
    for (int i = slot0; i < slot0 + n; ++i) {
        if (s->finished[i]) return fail(h, SNACB_ERR_STATE, "snacb_session_step: slot %d is finished (snacb_session_reset it)", i);
        if (s->frames[i] != s->frames[slot0] || s->emitted[i] != s->emitted[slot0] || s->origin[i] != s->origin[slot0])
            return fail(h, SNACB_ERR_STATE, "snacb_session_step: slots [%d, %d) are not at the same position (slot %d: %lld frames, slot %d: %lld)",
                        slot0, slot0 + n, slot0, s->frames[slot0], i, s->frames[i]);
    }
    if (new_frames > 0 && !new_tok) return fail(h, SNACB_ERR_ARG, "snacb_session_step: null tokens");
    const int Fm = s->max_frames;
    long long o = s->origin[slot0];
    int Fp = static_cast<int>(s->frames[slot0] - o);           // frames held, relative to the window
    int Ep = static_cast<int>(s->emitted[slot0] - 2048 * o);   // samples emitted, relative to the window
    int slide = 0;
    if (Fp + new_frames > Fm) {
        // the window is full: keep its last kKeep frames, move them to the front (after the checks below)
        slide = Fp - kKeep;
        if (slide < kKeep || kKeep + new_frames > Fm)
            return fail(h, SNACB_ERR_ARG, "snacb_session_step: %d new frames do not fit a %d-frame window holding %d (a step may add at most %d)",
                        new_frames, Fm, Fp, Fm - 2 * kKeep);
    }
    CK(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t so = static_cast<size_t>(slot0);
    int32_t* tokp = s->tok + so * 7 * Fm;
    if (slide > 0) {
        // rows [slide, Fp) frames of every stage -> [0, kKeep): source and destination do not overlap (slide >= kKeep)
        // the window keeps kKeep frames of TOKENS (the end-of-stream flush re-decodes from them) but only the last
        // kKeepAct frames of the activations: no stage reads further back than 3.1 frames (block 0: 2.2 frames of lag + 27 rows of taps)
        constexpr int kKeepAct = 4;
        auto move_n = [&](void* buf, size_t row_bytes_per_frame, int keep) -> cudaError_t {
            if (!buf) return cudaSuccess;
            const size_t pitch = row_bytes_per_frame * Fm;
            char* base = static_cast<char*>(buf) + so * pitch + row_bytes_per_frame * (kKeep - keep);
            return cudaMemcpy2DAsync(base, pitch, base + row_bytes_per_frame * slide, pitch, row_bytes_per_frame * keep, n,
                                     cudaMemcpyDeviceToDevice, st);
        };
        auto move = [&](void* buf, size_t row_bytes_per_frame) { return move_n(buf, row_bytes_per_frame, kKeepAct); };
        CK(h, move_n(s->tok, 7 * sizeof(int32_t), kKeep));
        CK(h, move(s->stem, static_cast<size_t>(4) * kDecDim * 2));
        size_t rows = 4;
        for (int bi = 0; bi < 4; ++bi) {
            rows *= h->blk[bi].s;
            const size_t rb = rows * h->blk[bi].Cout * 2;
            CK(h, move(s->ct[bi], rb));
            CK(h, move(s->nz[bi], rb));
            for (int ri = 0; ri < 3; ++ri) CK(h, move(s->res[bi][ri], rb));
            if (s->out[bi] != s->res[bi][2]) CK(h, move(s->out[bi], rb));
        }
        o += slide; Fp -= slide; Ep -= 2048 * slide;
        for (int i = slot0; i < slot0 + n; ++i) s->origin[i] = o;
    }
    const int F = Fp + new_frames;
    if (new_frames > 0)
        CK(h, cudaMemcpy2DAsync(tokp + static_cast<size_t>(7) * Fp, static_cast<size_t>(7) * Fm * sizeof(int32_t), new_tok,
                                static_cast<size_t>(tok_stride) * sizeof(int32_t), static_cast<size_t>(7) * new_frames * sizeof(int32_t),
                                n, cudaMemcpyDeviceToDevice, st));
    const int dflags = s->flags & (SNACB_RAW_IDS | SNACB_BF16);
    if (!stream_keys) stream_keys = s->slot_keys + slot0;     // a slot draws the same noise whichever slots share its step
    int n_out = 0;
    if (final) {
        // end of stream: the remaining samples see the zero padding at the true end -- one stateless ranged decode of the
        // window's tokens (exact end semantics; its left edge lies more than a receptive field before the first sample
        // still to emit, so the result equals the batch decode of the whole stream bit for bit)
        n_out = 2048 * F - Ep;
        if (n_out > 0) {
            if (!pcm || pcm_stride < n_out) return fail(h, SNACB_ERR_ARG, "snacb_session_step: pcm_stride %d < %d samples", pcm_stride, n_out);
            int16_t* dst = pcm;
            if (pcm_stride != n_out && n > 1) {
                size_t pb = h->st_pcm_elems * sizeof(int16_t);
                int rc = grow(h, reinterpret_cast<void**>(&h->st_pcm), &pb, static_cast<size_t>(n) * n_out * sizeof(int16_t));
                if (rc) return rc;
                h->st_pcm_elems = pb / sizeof(int16_t);
                dst = h->st_pcm;
            }
            int rc = decode_range_impl(h, tokp, n, 7 * Fm, F, dflags, nullptr, seed, stream_keys, Ep, 2048 * F, dst, nullptr, stream,
                                       static_cast<int>(o));
            if (rc) return rc;
            if (dst != pcm)
                CK(h, cudaMemcpy2DAsync(pcm, static_cast<size_t>(pcm_stride) * 2, dst, static_cast<size_t>(n_out) * 2,
                                        static_cast<size_t>(n_out) * 2, n, cudaMemcpyDeviceToDevice, st));
        }
        for (int i = slot0; i < slot0 + n; ++i) { s->frames[i] = o + F; s->emitted[i] = 2048 * (o + F); s->finished[i] = 1; }
        if (n_emitted) *n_emitted = n_out > 0 ? n_out : 0;
        return SNACB_OK;
    }
    if (new_frames == 0) return SNACB_OK;
    const Frontier a = frontier_of(h, Fp, s->flags), b = frontier_of(h, F, s->flags);
    if ((s->flags & SNACB_BF16) && !h->bf16_plain) {          // bf16x3: split operand of the widest GEMM input, absolute rows
        int rc = grow(h, &h->ws_split, &h->ws_split_bytes, static_cast<size_t>(n) * 131072 * Fm * 2 * 2);
        if (rc) return rc;
    }
    SessionPlan pl{};
    pl.c0 = s->codes[0] + so * Fm; pl.c1 = s->codes[1] + so * 2 * Fm; pl.c2 = s->codes[2] + so * 4 * Fm;
    size_t T = 4 * static_cast<size_t>(Fm);
    pl.a0 = static_cast<char*>(s->a0) + so * T * kLatent * 2;
    pl.stem = static_cast<char*>(s->stem) + so * T * kDecDim * 2;
    pl.stem_r = Rng{a.stem, b.stem};
    int vin_a = a.stem, vin_b = b.stem;
    for (int bi = 0; bi < 4; ++bi) {
        const BlockW& bw = h->blk[bi];
        T *= bw.s;
        const size_t off = so * T * bw.Cout * 2;
        auto at = [&](void* p) -> void* { return p ? static_cast<char*>(p) + off : nullptr; };
        pl.ct[bi] = at(s->ct[bi]); pl.nz[bi] = at(s->nz[bi]); pl.out[bi] = at(s->out[bi]);
        for (int ri = 0; ri < 3; ++ri) pl.res[bi][ri] = at(s->res[bi][ri]);
        // input rows whose successor is final too (the upper output phases of row m read row m + 1)
        pl.ct_in[bi] = Rng{vin_a > 0 ? vin_a - 1 : 0, vin_b > 0 ? vin_b - 1 : 0};
        pl.post[bi] = Rng{a.out[bi], b.out[bi]};
        pl.nz_r[bi] = Rng{a.nz[bi], b.nz[bi]};
        for (int ri = 0; ri < 3; ++ri) pl.res_r[bi][ri] = Rng{a.res[bi][ri], b.res[bi][ri]};
        vin_a = a.out[bi]; vin_b = b.out[bi];
    }
    n_out = b.emit - Ep;
    if (n_out < 0) n_out = 0;
    if (n_out > 0 && (!pcm || pcm_stride < n_out)) return fail(h, SNACB_ERR_ARG, "snacb_session_step: pcm_stride %d < %d samples", pcm_stride, n_out);
    int16_t* dst = pcm;
    if (n_out > 0 && pcm_stride != n_out && n > 1) {
        size_t pb = h->st_pcm_elems * sizeof(int16_t);
        int rc = grow(h, reinterpret_cast<void**>(&h->st_pcm), &pb, static_cast<size_t>(n) * n_out * sizeof(int16_t));
        if (rc) return rc;
        h->st_pcm_elems = pb / sizeof(int16_t);
        dst = h->st_pcm;
    }
    int rc = run_group(h, tokp, n, 7 * Fm, Fm, dflags, nullptr, seed, 0, stream_keys, Ep, Ep + n_out, dst, nullptr, st, &pl,
                       static_cast<int>(o));
    if (rc) return rc;
    if (dst != pcm)
        CK(h, cudaMemcpy2DAsync(pcm, static_cast<size_t>(pcm_stride) * 2, dst, static_cast<size_t>(n_out) * 2,
                                static_cast<size_t>(n_out) * 2, n, cudaMemcpyDeviceToDevice, st));
    for (int i = slot0; i < slot0 + n; ++i) { s->frames[i] = o + F; s->emitted[i] = 2048 * o + Ep + n_out; }
    if (n_emitted) *n_emitted = n_out;
    return SNACB_OK;
}

int snacb_session_step_multi(snacb_session s, int n, const int32_t* slots_host, const int32_t* new_tok, int tok_stride,
                             int new_frames, uint64_t seed, const int32_t* stream_keys, int16_t* pcm, int pcm_stride,
                             int* n_emitted, void* stream) {
    if (!s) return SNACB_ERR_ARG;
    snacb_handle h = s->h;
    if (n_emitted) *n_emitted = 0;
    if (n < 0 || new_frames <= 0 || tok_stride < 7 * new_frames || (n > 0 && (!slots_host || !new_tok)))
        return fail(h, SNACB_ERR_ARG, "snacb_session_step_multi: bad arguments n=%d new_frames=%d tok_stride=%d", n, new_frames, tok_stride);
    if (n == 0) return SNACB_OK;
    if (n > s->n_slots) return fail(h, SNACB_ERR_ARG, "snacb_session_step_multi: %d streams for %d slots", n, s->n_slots);
    constexpr int kKeep = 8;
    const int Fm = s->max_frames;
    // ---- positions (frames held relative to each window), windows that have to slide first
    std::vector<int32_t> hs(static_cast<size_t>(7) * n);
    int32_t *v_slot = hs.data(), *v_off = v_slot + n, *v_org = v_off + n, *v_pos = v_org + n, *v_list = v_pos + 2 * n, *v_slide = v_list + n;
    std::vector<char> seen(s->n_slots, 0);
    int n_slide = 0;
    bool same = true;
    for (int i = 0; i < n; ++i) {
        const int sl = slots_host[i];
        if (sl < 0 || sl >= s->n_slots || seen[sl]) return fail(h, SNACB_ERR_ARG, "snacb_session_step_multi: bad or repeated slot %d", sl);
        seen[sl] = 1;
        if (s->finished[sl]) return fail(h, SNACB_ERR_STATE, "snacb_session_step_multi: slot %d is finished (snacb_session_reset it)", sl);
        int Fp = static_cast<int>(s->frames[sl] - s->origin[sl]);
        if (Fp + new_frames > Fm) {
            const int slide = Fp - kKeep;
            if (slide < kKeep || kKeep + new_frames > Fm)
                return fail(h, SNACB_ERR_ARG, "snacb_session_step_multi: %d new frames do not fit a %d-frame window holding %d", new_frames, Fm, Fp);
            v_list[n_slide] = sl; v_slide[n_slide] = slide; ++n_slide;
            v_org[i] = static_cast<int32_t>(s->origin[sl] + slide);
            Fp = kKeep;
        } else {
            v_org[i] = static_cast<int32_t>(s->origin[sl]);
        }
        v_slot[i] = sl; v_pos[i] = Fp;
        same = same && Fp == v_pos[0];
    }
    const int Fref = v_pos[0];
    for (int i = 0; i < n; ++i) {
        // past its first frames every stage's frontier is affine in the frame count, so streams at different positions share
        // the launch: stream i's row ranges are the reference's shifted by (pos_i - pos_ref) frames
        if (!same && v_pos[i] < 3)
            return fail(h, SNACB_ERR_STATE, "snacb_session_step_multi: slot %d holds %d frames; streams with fewer than 3 frames can only share a step with streams at the same position", v_slot[i], v_pos[i]);
        v_off[i] = v_pos[i] - Fref;
    }
    const Frontier a = frontier_of(h, Fref, s->flags), b = frontier_of(h, Fref + new_frames, s->flags);
    const int n_out = b.emit - a.emit;                   // = what every stream of the step emits (its own emitted count lags alike)
    for (int i = 0; i < n; ++i) {
        const int sl = v_slot[i];
        const long long el = s->emitted[sl] - 2048LL * v_org[i];  // samples emitted, relative to the (new) window
        if (el != frontier_of(h, v_pos[i], s->flags).emit)
            return fail(h, SNACB_ERR_STATE, "snacb_session_step_multi: slot %d is out of step with its frontier", sl);
    }
    if (n_out > 0 && (!pcm || pcm_stride < n_out)) return fail(h, SNACB_ERR_ARG, "snacb_session_step_multi: pcm_stride %d < %d samples", pcm_stride, n_out);
    CK(h, cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t N = static_cast<size_t>(s->n_slots);
    int32_t *d_slot = s->map_dev, *d_off = d_slot + N, *d_org = d_off + N, *d_pos = d_org + N, *d_keys = d_pos + N,
            *d_list = d_keys + N, *d_slide = d_list + N;
    CK(h, cudaMemcpyAsync(d_slot, v_slot, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(h, cudaMemcpyAsync(d_off, v_off, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(h, cudaMemcpyAsync(d_org, v_org, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(h, cudaMemcpyAsync(d_pos, v_pos, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (n_slide > 0) {
        CK(h, cudaMemcpyAsync(d_list, v_list, n_slide * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CK(h, cudaMemcpyAsync(d_slide, v_slide, n_slide * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        constexpr int kKeepAct = 4;      // activations: only the last 4 frames of the kept window are live (see snacb_session_step)
        auto move = [&](void* buf, size_t frame_bytes) {
            if (buf) launch_session_slide(static_cast<char*>(buf) + frame_bytes * (kKeep - kKeepAct), frame_bytes * Fm, frame_bytes, kKeepAct,
                                          d_list, d_slide, n_slide, st);
        };
        // the token rows are 28 bytes per frame (sources not 16-byte aligned): a plain copy per sliding slot
        for (int i = 0; i < n_slide; ++i) {
            int32_t* row = s->tok + static_cast<size_t>(v_list[i]) * 7 * Fm;
            CK(h, cudaMemcpyAsync(row, row + static_cast<size_t>(7) * v_slide[i], static_cast<size_t>(7) * kKeep * sizeof(int32_t),
                                  cudaMemcpyDeviceToDevice, st));
        }
        move(s->stem, static_cast<size_t>(4) * kDecDim * 2);
        size_t rows = 4;
        for (int bi = 0; bi < 4; ++bi) {
            rows *= h->blk[bi].s;
            const size_t rb = rows * h->blk[bi].Cout * 2;
            move(s->ct[bi], rb);
            move(s->nz[bi], rb);
            for (int ri = 0; ri < 3; ++ri) move(s->res[bi][ri], rb);
            if (s->out[bi] != s->res[bi][2]) move(s->out[bi], rb);
        }
        CK(h, cudaGetLastError());
    }
    launch_session_scatter(new_tok, tok_stride, 7 * new_frames, d_slot, d_pos, s->tok, 7 * Fm, stream_keys, d_keys, n, st);
    CK(h, cudaGetLastError());
    if ((s->flags & SNACB_BF16) && !h->bf16_plain) {
        int rc = grow(h, &h->ws_split, &h->ws_split_bytes, N * 131072 * Fm * 2 * 2);     // indexed by slot
        if (rc) return rc;
    }
    SessionPlan pl{};
    pl.c0 = s->codes[0]; pl.c1 = s->codes[1]; pl.c2 = s->codes[2];
    pl.a0 = s->a0; pl.stem = s->stem;
    pl.stem_r = Rng{a.stem, b.stem};
    int vin_a = a.stem, vin_b = b.stem;
    for (int bi = 0; bi < 4; ++bi) {
        pl.ct[bi] = s->ct[bi]; pl.nz[bi] = s->nz[bi]; pl.out[bi] = s->out[bi];
        for (int ri = 0; ri < 3; ++ri) pl.res[bi][ri] = s->res[bi][ri];
        pl.ct_in[bi] = Rng{vin_a > 0 ? vin_a - 1 : 0, vin_b > 0 ? vin_b - 1 : 0};
        pl.post[bi] = Rng{a.out[bi], b.out[bi]};
        pl.nz_r[bi] = Rng{a.nz[bi], b.nz[bi]};
        for (int ri = 0; ri < 3; ++ri) pl.res_r[bi][ri] = Rng{a.res[bi][ri], b.res[bi][ri]};
        vin_a = a.out[bi]; vin_b = b.out[bi];
    }
    pl.map = StreamMap{d_slot, d_off, d_org};
    pl.n_slots = s->n_slots;
    int16_t* dst = pcm;
    if (n_out > 0 && pcm_stride != n_out && n > 1) {
        size_t pb = h->st_pcm_elems * sizeof(int16_t);
        int rc = grow(h, reinterpret_cast<void**>(&h->st_pcm), &pb, static_cast<size_t>(n) * n_out * sizeof(int16_t));
        if (rc) return rc;
        h->st_pcm_elems = pb / sizeof(int16_t);
        dst = h->st_pcm;
    }
    const int dflags = s->flags & (SNACB_RAW_IDS | SNACB_BF16);
    int rc = run_group(h, s->tok, n, 7 * Fm, Fm, dflags, nullptr, seed, 0, d_keys, a.emit, a.emit + (n_out > 0 ? n_out : 0), dst,
                       nullptr, st, &pl, 0);
    if (rc) return rc;
    if (dst != pcm && n_out > 0)
        CK(h, cudaMemcpy2DAsync(pcm, static_cast<size_t>(pcm_stride) * 2, dst, static_cast<size_t>(n_out) * 2,
                                static_cast<size_t>(n_out) * 2, n, cudaMemcpyDeviceToDevice, st));
    for (int i = 0; i < n; ++i) {
        const int sl = v_slot[i];
        s->origin[sl] = v_org[i];
        s->frames[sl] += new_frames;
        s->emitted[sl] += n_out > 0 ? n_out : 0;
    }
    if (n_emitted) *n_emitted = n_out > 0 ? n_out : 0;
    return SNACB_OK;
}

int snacb_decode_host(snacb_handle h, const int32_t* tok_host, int B, int tok_stride, int frames, int flags,
                      uint64_t seed, int16_t* pcm_host) {
    if (!h) return SNACB_ERR_ARG;
    if (B < 0 || frames < 0 || tok_stride < frames * kFrame) return fail(h, SNACB_ERR_ARG, "snacb_decode_host: bad sizes");
    if (B == 0 || frames == 0) return SNACB_OK;
    if (!tok_host || !pcm_host) return fail(h, SNACB_ERR_ARG, "snacb_decode_host: null buffer");
    CK(h, cudaSetDevice(h->device));
    const size_t ntok = static_cast<size_t>(B) * tok_stride;
    const size_t npcm = static_cast<size_t>(B) * snacb_samples_out(frames, flags);
    size_t tb = h->st_tok_elems * sizeof(int32_t), pb = h->st_pcm_elems * sizeof(int16_t);
    int rc = grow(h, reinterpret_cast<void**>(&h->st_tok), &tb, ntok * sizeof(int32_t));
    if (rc) return rc;
    h->st_tok_elems = tb / sizeof(int32_t);
    rc = grow(h, reinterpret_cast<void**>(&h->st_pcm), &pb, npcm * sizeof(int16_t));
    if (rc) return rc;
    h->st_pcm_elems = pb / sizeof(int16_t);
    cudaStream_t st = h->own_stream;
    CK(h, cudaMemcpyAsync(h->st_tok, tok_host, ntok * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    rc = snacb_decode(h, h->st_tok, B, tok_stride, frames, flags, nullptr, seed, h->st_pcm, nullptr, st);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(pcm_host, h->st_pcm, npcm * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));
    return SNACB_OK;
}

// Pipelined form of the host boundary: submit() returns once the work is queued, wait() blocks until the OLDEST
// outstanding submit has its PCM in the caller's buffer.  With one submit in flight while the previous one is being
// waited on, the device->host copy of step i (copy stream) overlaps the decode of step i + 1 (compute stream).
int snacb_decode_host_submit(snacb_handle h, const int32_t* tok_host, int B, int tok_stride, int frames, int flags,
                             uint64_t seed, int16_t* pcm_host) {
    if (!h) return SNACB_ERR_ARG;
    if (B <= 0 || frames <= 0 || tok_stride < frames * kFrame) return fail(h, SNACB_ERR_ARG, "snacb_decode_host_submit: bad sizes");
    if (!tok_host || !pcm_host) return fail(h, SNACB_ERR_ARG, "snacb_decode_host_submit: null buffer");
    if (h->pl_submitted - h->pl_waited >= 2) return fail(h, SNACB_ERR_STATE, "snacb_decode_host_submit: two submits outstanding, call snacb_decode_host_wait");
    CK(h, cudaSetDevice(h->device));
    const int slot = static_cast<int>(h->pl_submitted & 1);
    if (!h->copy_stream) CK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->pl_done[slot]) {
        CK(h, cudaEventCreateWithFlags(&h->pl_done[slot], cudaEventDisableTiming));
        CK(h, cudaEventCreateWithFlags(&h->pl_copied[slot], cudaEventDisableTiming));
    }
    const size_t tbytes = static_cast<size_t>(B) * tok_stride * sizeof(int32_t);
    const size_t pbytes = static_cast<size_t>(B) * snacb_samples_out(frames, flags) * sizeof(int16_t);
    int rc = grow(h, reinterpret_cast<void**>(&h->pl_tok[slot]), &h->pl_tok_bytes[slot], tbytes);
    if (rc) return rc;
    rc = grow(h, reinterpret_cast<void**>(&h->pl_pcm[slot]), &h->pl_pcm_bytes[slot], pbytes);
    if (rc) return rc;
    cudaStream_t st = h->own_stream;
    CK(h, cudaMemcpyAsync(h->pl_tok[slot], tok_host, tbytes, cudaMemcpyHostToDevice, st));
    rc = snacb_decode(h, h->pl_tok[slot], B, tok_stride, frames, flags, nullptr, seed, h->pl_pcm[slot], nullptr, st);
    if (rc) return rc;
    CK(h, cudaEventRecord(h->pl_done[slot], st));
    CK(h, cudaStreamWaitEvent(h->copy_stream, h->pl_done[slot], 0));
    CK(h, cudaMemcpyAsync(pcm_host, h->pl_pcm[slot], pbytes, cudaMemcpyDeviceToHost, h->copy_stream));
    CK(h, cudaEventRecord(h->pl_copied[slot], h->copy_stream));
    h->pl_submitted++;
    return SNACB_OK;
}

int snacb_decode_host_wait(snacb_handle h) {
    if (!h) return SNACB_ERR_ARG;
    if (h->pl_waited == h->pl_submitted) return fail(h, SNACB_ERR_STATE, "snacb_decode_host_wait: nothing outstanding");
    CK(h, cudaSetDevice(h->device));
    const int slot = static_cast<int>(h->pl_waited & 1);
    CK(h, cudaEventSynchronize(h->pl_copied[slot]));
    h->pl_waited++;
    return SNACB_OK;
}

int snacb_profile(snacb_handle h, int enable) {
    if (!h) return SNACB_ERR_ARG;
    for (auto& r : h->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    h->prof_recs.clear();
    h->prof_names.clear();
    h->prof_on = enable != 0;
    return SNACB_OK;
}

int snacb_profile_report(snacb_handle h, char* buf, size_t cap) {
    if (!h || !buf || cap == 0) return SNACB_ERR_ARG;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    std::vector<double> tot(h->prof_names.size(), 0.0);
    std::vector<long> cnt(h->prof_names.size(), 0);
    for (auto& r : h->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { tot[r.name_id] += ms; cnt[r.name_id]++; }
    }
    std::string out;
    char line[128];
    for (size_t i = 0; i < h->prof_names.size(); ++i) {
        snprintf(line, sizeof line, "%s %ld %.6f\n", h->prof_names[i].c_str(), cnt[i], tot[i]);
        out += line;
    }
    if (out.size() + 1 > cap) return fail(h, SNACB_ERR_ARG, "snacb_profile_report: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return SNACB_OK;
}

int snacb_debug_chain_spans(int C, int16_t* out, int cap) {
    if (!out || !chain_supported(C, 1) || cap < 3 * kChainWarps * kChainSpans * 3) return SNACB_ERR_ARG;
    ChainSpan sp[3][kChainWarps][kChainSpans];
    chain_build_spans(C, sp);
    for (int l = 0; l < 3; ++l)
        for (int w = 0; w < kChainWarps; ++w)
            for (int k = 0; k < kChainSpans; ++k) {
                int16_t* o = out + ((l * kChainWarps + w) * kChainSpans + k) * 3;
                o[0] = sp[l][w][k].r_first; o[1] = sp[l][w][k].n_oct; o[2] = sp[l][w][k].kc;
            }
    return chain_tile_rows(C) | (chain_warps(C) << 16);
}

int snacb_debug_chain_spans_ex(int C, int own_end, int carry_top, int16_t* out, int cap) {
    if (!out || !chain_supported(C, 1) || cap < 3 * kChainWarps * kChainSpans * 4) return SNACB_ERR_ARG;
    if (own_end < 0 || own_end > chain_tile_rows(C) - kChainHalo) return SNACB_ERR_ARG;
    ChainSpan sp[3][kChainWarps][kChainSpans];
    if (!chain_build_spans(C, sp, own_end, carry_top != 0)) return 0;
    for (int l = 0; l < 3; ++l)
        for (int w = 0; w < kChainWarps; ++w)
            for (int k = 0; k < kChainSpans; ++k) {
                int16_t* o = out + ((l * kChainWarps + w) * kChainSpans + k) * 4;
                o[0] = sp[l][w][k].r_first; o[1] = sp[l][w][k].n_oct; o[2] = sp[l][w][k].kc; o[3] = sp[l][w][k].pad;
            }
    return chain_tile_rows(C) | (chain_warps(C) << 16);
}

int snacb_debug_chain_plan(int C, int t_n, int S, int slots, int32_t* out4) {
    if (!out4 || !chain_supported(C, 1) || t_n <= 0 || S <= 0 || slots <= 0) return SNACB_ERR_ARG;
    int k, sps, lst, lrows;
    chain_plan_strips(C, t_n, S, slots, false, &k, &sps, &lst, &lrows);
    out4[0] = k; out4[1] = sps; out4[2] = lst; out4[3] = lrows;
    return SNACB_OK;
}

int snacb_experiments_built(void) { return chain_ws_built() ? 1 : 0; }

int snacb_debug_chain_ws_spans(int C, int16_t* out, int cap) {
    if (!out || !chain_ws_supported(C, 1) || cap < 3 * kChainWarps * kChainSpans * 3) return SNACB_ERR_ARG;
    ChainSpan sp[3][kChainWarps][kChainSpans];
    chain_ws_build_spans(C, sp);
    for (int l = 0; l < 3; ++l)
        for (int w = 0; w < kChainWarps; ++w)
            for (int k = 0; k < kChainSpans; ++k) {
                int16_t* o = out + ((l * kChainWarps + w) * kChainSpans + k) * 3;
                o[0] = sp[l][w][k].r_first; o[1] = sp[l][w][k].n_oct; o[2] = sp[l][w][k].kc;
            }
    return chain_ws_tile_rows(C) | (kChainWsP << 16);
}

int snacb_debug_tap_count(snacb_handle h) { return h ? static_cast<int>(h->taps.size()) : SNACB_ERR_ARG; }

int snacb_debug_tap_info(snacb_handle h, int idx, char* name, int name_cap, int64_t* rows, int64_t* cols) {
    if (!h || idx < 0 || idx >= static_cast<int>(h->taps.size())) return SNACB_ERR_ARG;
    const Tap& t = h->taps[idx];
    if (name && name_cap > 0) { strncpy(name, t.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (rows) *rows = t.rows;
    if (cols) *cols = t.cols;
    return SNACB_OK;
}

int snacb_debug_tap_copy(snacb_handle h, int idx, float* dst_host, size_t dst_elems) {
    if (!h || idx < 0 || idx >= static_cast<int>(h->taps.size()) || !dst_host) return SNACB_ERR_ARG;
    const Tap& t = h->taps[idx];
    const size_t n = static_cast<size_t>(t.rows) * t.cols;
    if (dst_elems < n) return fail(h, SNACB_ERR_ARG, "snacb_debug_tap_copy: buffer too small");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaDeviceSynchronize());
    CK(h, cudaMemcpy(dst_host, t.dev, n * sizeof(float), cudaMemcpyDeviceToHost));
    return SNACB_OK;
}

}  // extern "C"
