// Launch-side interface between the host pipeline (snacb.cu) and the kernel translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>

#include "common.cuh"

namespace snacb {

// cudaFuncSetAttribute (opt-in dynamic shared memory, carveout) is per DEVICE: a process may hold handles on several GPUs
// and use them from several threads, so every launcher keeps one atomic bit per device instead of a process-wide flag.
// Setting an attribute twice is harmless; what must not happen is skipping it on the second device.
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    // true when the calling thread's current device has not been configured through this object yet
    bool needed(int* dev_out) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { *dev_out = -1; return true; }
        *dev_out = dev;
        return (mask.load(std::memory_order_acquire) & (1ull << dev)) == 0;
    }
    void done(int dev) { if (dev >= 0) mask.fetch_or(1ull << dev, std::memory_order_release); }
};

struct VqStemWeights {
    const float* codebook[3];  // [4096][8]
    const float* out_w[3];     // [768][8]
    const float* out_b[3];     // [768]
    const float* dw_w;         // [7][768]  (tap-major so that channels are contiguous)
    const float* dw_b;         // [768]
};

// ---- kernels_simt.cu
void launch_unpack(const int32_t* tok, int B, int ntok, int F, int raw_ids, int32_t* c0, int32_t* c1, int32_t* c2,
                   cudaStream_t st);
template <typename OutT>
void launch_vq_stem(const int32_t* c0, const int32_t* c1, const int32_t* c2, const int32_t* tok, int tok_stride, int raw_ids,
                    int S, int F, int t_lo, int t_hi, const VqStemWeights& w, OutT* out, cudaStream_t st,
                    const StreamMap& map = StreamMap{nullptr, nullptr, nullptr});
// latent steps [t_lo, t_hi) of every stream; tok != null: codes are unpacked from the token rows on the fly (c0..c2 unused)
void launch_gemm_f32(int epi, const GemmArgs& a, const float* A, const float* W, cudaStream_t st);
void launch_respre_f32(const ResUnitArgs& a, float* P, cudaStream_t st);
// streaming session: scatter each stream's new tokens to its own position; slide full windows (kernels_simt.cu)
void launch_session_scatter(const int32_t* new_tok, int tok_stride, int n_int, const int* slot, const int* pos, int32_t* tok_buf,
                            int slot_ints, const int32_t* keys_in, int32_t* keys_out, int n, cudaStream_t st);
void launch_session_slide(void* buf, size_t slot_bytes, size_t frame_bytes, int keep, const int* list, const int* slide, int n,
                          cudaStream_t st);
// bf16x3 path: fp16 rows [r_lo, r_lo + r_n) of every stream -> bf16 [rows][2K] = [hi | lo] (exact split; GemmArgs::a_wrap)
void launch_split3(const __half* x, int S, int T, int K, int r_lo, int r_n, __nv_bfloat16* out, cudaStream_t st,
                   const StreamMap& map = StreamMap{nullptr, nullptr, nullptr}, int rpf = 0);
// bf16x3 path: ResidualUnit front half from the fp16 stream, rows [r_lo, r_lo + r_n) -> bf16 [rows][C]
void launch_respre16(const ResUnitArgs& a, int r_lo, int r_n, __nv_bfloat16* out, cudaStream_t st);
template <typename InT>
void launch_tail(const InT* a, int S, int T, int t_begin, int n_out, const float* w, float bias, int16_t* pcm,
                 float* wave, cudaStream_t st, const StreamMap& map = StreamMap{nullptr, nullptr, nullptr},
                 const CUtensorMap* tm128 = nullptr, const CUtensorMap* tm8 = nullptr);   // maps: the tensor-core tail (16-bit input)
template <typename T>
void launch_to_f32(const T* in, float* out, size_t n, cudaStream_t st);

// ---- kernels_tc.cu  (tcgen05 / TMEM / TMA)
// out dtype: 0 = bf16, 1 = fp32.  Returns cudaError_t of the launch.
int gemm_tc_block_n(const GemmArgs& a);   // column-tile width the tensor-core GEMM will use for this problem
// half_fp16: 16-bit operand / storage type, 0 = bf16, 1 = fp16 (same tensor-core rate, fp32 accumulate)
cudaError_t launch_gemm_tc(int epi, int half_fp16, int out_f32, const GemmArgs& a, const CUtensorMap& tmA,
                           const CUtensorMap& tmW, const CUtensorMap& tmY, const CUtensorMap& tmO, int sm_count, cudaStream_t st);
                           // tmY / tmO: GemmArgs::tma_epi (NoiseBlock epilogue by TMA), else any valid map
cudaError_t launch_resunit_tc(int epi, int half_fp16, int x_f32, const ResUnitArgs& a, const CUtensorMap& tmW,
                              cudaStream_t st);
cudaError_t init_tc_kernels();            // opt-in shared memory sizes

// ---- kernels_res2.cu  (persistent pipelined ResidualUnit, 16-bit activations)
void resunit2_geometry(int C, int dil, int* tile_m, int* box_rows);   // x tensor-map box = (64, box_rows, 1)
bool resunit2_swizzled_x(int C);          // C = 512: the x map is 128B-swizzled (the x chunk doubles as the residual MMA's A operand)
cudaError_t launch_resunit2(int half_fp16, const ResUnitArgs& a, const CUtensorMap& tmX, const CUtensorMap& tmW,
                            const CUtensorMap& tmO, int sm_count, cudaStream_t st);   // tmO (C = 512): output, box (32, 32, 1), 64B-swizzled

// ---- kernels_convt.cu  (ConvTranspose1d with resident weights and row-shifted UMMA descriptors; block 3)
bool convt_res_supported(int Cin, int Cout, int s);
int convt_res_box_rows();             // activation tensor-map box = (64, box_rows, 1), 128B swizzle
cudaError_t launch_convt_res(int half_fp16, const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW,
                             const CUtensorMap& tmO, int sm_count, cudaStream_t st);   // tmO: output box (64, 128*s, 1)

cudaError_t launch_convt_ph2(int half_fp16, const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW,
                             const CUtensorMap& tmO, int sm_count, cudaStream_t st);   // CTA pairs (cta_group::2); tmW box (64, Cout / 2)
bool convt_ph_supported(int Cin, int Cout, int s);   // block 2: one output phase's weights resident per CTA group
cudaError_t launch_convt_ph(int half_fp16, const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW,
                            const CUtensorMap& tmO, int sm_count, cudaStream_t st);
                            // tmA box (64, 136, 1); tmW box (64, Cout); tmO: output as [slot][Tin][s*Cout], box (64, 32, 1)

// ---- kernels_chain.cu  (NoiseBlock + 3 ResidualUnits fused, residual stream in TMEM)
bool chain_supported(int C, int half_fp16);
int chain_tile_rows(int C);           // rows of a tile incl. the halo (y tensor-map box = (64, 128, 1), 128B swizzle)
int chain_warps(int C);               // warps per CTA of the launch configuration used for C channels
// Schedule of one tile type: own_end = first tile row past the owned rows (0: a full tile, rows - 40); carry_top: the
// tile owns its rows from row 0 on (classes start at or above row 0, top spans flagged in ChainSpan::pad) instead of
// from row 40.  false (table untouched) when the tile does not fit the table -- the caller then falls back.
bool chain_build_spans(int C, ChainSpan (*spans)[kChainWarps][kChainSpans], int own_end = 0, bool carry_top = false);
// strips for a row range of t_n rows per stream, S streams on `slots` CTA slots: tiles per strip, strips per stream, tiles
// of the last strip, rows the last tile owns (0 = full).  no_carry: one halo-top tile per strip (the round-1 geometry)
void chain_plan_strips(int C, int t_n, int S, int slots, bool no_carry, int* strip_tiles, int* sps, int* last_strip_tiles,
                       int* last_rows);
size_t chain_carry_bytes(int C, int sm_count);   // scratch for ChainArgs::carry
size_t chain_carry_bytes_max(int sm_count);      // the largest over the supported widths
// tm[7]: y load map box (64,128,1); out store maps box (64,128-kChainHalo,1) and (64,128,1); noise 1x1, res d=1,
// d=3, d=9 weight maps box (64, C); all 128B-swizzled.  fold = 1: the alpha-folded fp16 formulation (the three res
// weight maps then point at the copies with 1 / alpha2 folded into their K columns); see chain_fold_safe in snacb.cu
cudaError_t launch_chain(int half_fp16, int fold, const ChainArgs& a, const CUtensorMap* tm, int sm_count, cudaStream_t st);

// ---- kernels_chain_ws.cu  (the same chain, warp-specialised and pipelined over the 128-row blocks of a tile; C = 64 / 128)
bool chain_ws_built();                // false unless the library was built with SNACB_EXPERIMENTS=1 (measured slower: DESIGN.md section 6)
bool chain_ws_supported(int C, int half_fp16);
int chain_ws_tile_rows(int C);
void chain_ws_build_spans(int C, ChainSpan (*spans)[kChainWarps][kChainSpans]);   // ChainSpan::n_oct counts QUADS here
cudaError_t launch_chain_ws(int half_fp16, int fold, const ChainArgs& a, const CUtensorMap* tm, int sm_count, cudaStream_t st);

}  // namespace snacb
