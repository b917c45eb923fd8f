// The inner arithmetic of the chain kernels' in-place prologue, shared by kernels_chain.cu and kernels_chain_ws.cu:
// NOUT (8 or 4) consecutive outputs of one dilation class for one channel pair per lane,
//     a[k] = snake2( bd + sum_j w[j] * xs[k + j] ),   xs[i] = S1 of the class row (i - 3) steps from the first output.
//
// Written TAP-MAJOR: the NOUT accumulators advance together, one tap at a time, so that consecutive instructions are
// independent (NOUT-way ILP in source order) and the Snake stages run NOUT wide.  Step-major source (one output's seven
// dependent FMAs after the other) left ptxas with ~3-way interleaving and the warps at 0.13-0.17 IPC, stalled on the
// 4-cycle dependent-issue latency (profiles/r02_chain_ncu.txt: `wait` + `short_scoreboard`); the per-output operation
// order -- and therefore every bit of the result -- is the same as before.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace snacb {

__device__ __forceinline__ __half2 cs_h2(uint32_t v) { return *reinterpret_cast<const __half2*>(&v); }
__device__ __forceinline__ uint32_t cs_u32(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }

// fp16 math.  FOLD: xs / w / bd carry the alpha scalings (kernels_chain.cu), snake2 = a + sin^2 a on a = alpha2 * conv.
// !FOLD: al2 / ia2 = alpha2, 1 / (alpha2 + 1e-9) of the channel pair; Snake in fp32.
template <int NOUT, bool FOLD>
__device__ __forceinline__ void dw_snake_half(const uint32_t (&xs)[14], const __half2 (&w)[7], const __half2 bd,
                                              const float2 al2, const float2 ia2, uint32_t (&o)[8]) {
    __half2 acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = bd;
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int k = 0; k < NOUT; ++k) acc[k] = __hfma2(w[j], cs_h2(xs[k + j]), acc[k]);
    float2 t[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) t[k] = __half22float2(acc[k]);
    if (FOLD) {
        float2 s[NOUT];
#pragma unroll
        for (int k = 0; k < NOUT; ++k) s[k] = make_float2(__sinf(t[k].x), __sinf(t[k].y));
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            const __half2 sh = __floats2half2_rn(s[k].x, s[k].y);
            o[k] = cs_u32(__hfma2(sh, sh, acc[k]));
        }
    } else {
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            const float2 r = snake_pair(t[k], al2, ia2);
            o[k] = cs_u32(__floats2half2_rn(r.x, r.y));
        }
    }
}

// bf16 operands: fp32 math (an 8-bit mantissa cannot carry the depthwise accumulation)
__device__ __forceinline__ float2 cs_unpack_bf16(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));      // a shift: ALU pipe, not FMA
}
template <int NOUT>
__device__ __forceinline__ void dw_snake_bf16(const uint32_t (&xs)[14], const float2 (&w)[7], const float2 bd,
                                              const float2 al2, const float2 ia2, uint32_t (&o)[8]) {
    float2 x[NOUT + 6];
#pragma unroll
    for (int i = 0; i < NOUT + 6; ++i) x[i] = cs_unpack_bf16(xs[i]);
    float2 acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = bd;
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int k = 0; k < NOUT; ++k) acc[k] = ffma2(w[j], x[k + j], acc[k]);
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
        const float2 r = snake_pair(acc[k], al2, ia2);
        o[k] = pack2(r.x, r.y, static_cast<const __nv_bfloat16*>(nullptr));
    }
}

}  // namespace snacb
