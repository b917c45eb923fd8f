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

// The TMEM-drain epilogues of the chain kernels are bound by the XU pipe (32 MUFU.SIN per 32 x 32 piece): the column pairs
// selected by kEpiPolyMask take their sin^2 from the FMA pipe instead.
// x + sin^2 x on a half2 without the XU pipe: r = x / pi - rint(x / pi) by the magic-number trick, sin^2(pi r) as a
// degree-4 odd-free minimax polynomial in r^2 (|r| <= 1/2); 9 FMA-pipe instructions with immediate operands
__device__ __forceinline__ __half2 snake_h2_poly(__half2 xh) {
    const __half2 kInvPi = __float2half2_rn(0.318309886f), kMagic = __float2half2_rn(1536.f);
    const __half2 m = __hfma2(xh, kInvPi, kMagic);
    const __half2 n = __hsub2(m, kMagic);
    const __half2 r = __hfma2(xh, kInvPi, __hneg2(n));
    const __half2 u = __hmin2(__hmul2(r, r), __float2half2_rn(0.25f));      // |x| >= 1608 (no phase left in fp16): stay finite
    __half2 p = __hfma2(__float2half2_rn(-22.99092533f), u, __float2half2_rn(41.29496355f));
    p = __hfma2(p, u, __float2half2_rn(-32.35387252f));
    p = __hfma2(p, u, __float2half2_rn(9.86667475f));
    return __hfma2(p, u, xh);
}
#ifndef SNACB_EPI_POLY_MASK
#define SNACB_EPI_POLY_MASK 0x8888u
#endif
constexpr unsigned kEpiPolyMask = SNACB_EPI_POLY_MASK;   // bit i: column pair i of a 32-column piece takes the polynomial

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

// One span of the in-place prologue inside a 128-row block: rows r_first + k*D (block-relative), k < 4 * nquad, of one
// 64-channel chunk; lane = channel pair.  `blk` = the block's first row in the chunk plane.  The span may start at any
// row, so the 128B-swizzle phase of its steps is computed here (ph[k & 7]: 8*D rows further the phase repeats).
// h0..h2 / t0..t2: the three rows before / after the span, fetched before the block's barrier.
template <int D, bool FOLD>
__device__ __forceinline__ void ws_span_half(uint8_t* blk, int r_first, const int nquad, const uint32_t h0, const uint32_t h1,
                                             const uint32_t h2, const uint32_t t0, const uint32_t t1, const uint32_t t2,
                                             const uint32_t lane_off, const uint32_t* prm) {
    const uint4 q0 = *reinterpret_cast<const uint4*>(prm), q1 = *reinterpret_cast<const uint4*>(prm + 4);
    const __half2 bd = cs_h2(q0.x);
    const __half2 w[7] = {cs_h2(q0.y), cs_h2(q0.z), cs_h2(q0.w), cs_h2(q1.x), cs_h2(q1.y), cs_h2(q1.z), cs_h2(q1.w)};
    float2 al2 = make_float2(0.f, 0.f), ia2 = make_float2(0.f, 0.f);
    if (!FOLD) {
        const uint4 q2 = *reinterpret_cast<const uint4*>(prm + 8);
        al2 = make_float2(__uint_as_float(q2.x), __uint_as_float(q2.y));
        ia2 = make_float2(__uint_as_float(q2.z), __uint_as_float(q2.w));
    }
    uint32_t ph[8];
    const uint32_t p0s = static_cast<uint32_t>(r_first & 7) << 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) ph[j] = ((p0s + static_cast<uint32_t>(((j * D) & 7) << 4)) & 0x70u) ^ lane_off;
    uint32_t xs[14];                               // xs[i] = S1 of the class row (i - 3) steps from the octet's first row
    xs[0] = h0; xs[1] = h1; xs[2] = h2;
    uint8_t* ob = blk + r_first * 128;
#pragma unroll
    for (int j = 0; j < 3; ++j) xs[3 + j] = *reinterpret_cast<const uint32_t*>(ob + j * D * 128 + ph[j]);
    const int noct = (nquad + 1) >> 1;
#pragma unroll 1
    for (int qo = 0; qo < noct; ++qo) {
        const bool full = nquad - 2 * qo >= 2;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            xs[6 + k] = *reinterpret_cast<const uint32_t*>(ob + (k + 3) * D * 128 + ph[(k + 3) & 7]);
        if (qo == noct - 1) {
            if (full) { xs[11] = t0; xs[12] = t1; xs[13] = t2; }
            else { xs[7] = t0; xs[8] = t1; xs[9] = t2; }
        }
        uint32_t o[8];
        if (full) dw_snake_half<8, FOLD>(xs, w, bd, al2, ia2, o);
        else dw_snake_half<4, FOLD>(xs, w, bd, al2, ia2, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k == 4 && !full) break;
            if (r_first + k * D < 128)                      // the last quad of a class may run past the block
                *reinterpret_cast<uint32_t*>(ob + k * D * 128 + ph[k]) = o[k];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) xs[i] = xs[8 + i];
        r_first += 8 * D;
        ob += 8 * D * 128;
    }
}

// same span, fp32 math (bf16 operands)
template <int D>
__device__ __forceinline__ void ws_span_bf16(uint8_t* blk, int r_first, const int nquad, const uint32_t h0, const uint32_t h1,
                                             const uint32_t h2, const uint32_t t0, const uint32_t t1, const uint32_t t2,
                                             const uint32_t lane_off, const uint32_t* prm) {
    const float4* p4 = reinterpret_cast<const float4*>(prm);
    const float4 q0 = p4[0], q1 = p4[1], q2 = p4[2], q3 = p4[3], q4 = p4[4];
    const float2 w[7] = {make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y), make_float2(q1.z, q1.w),
                         make_float2(q2.x, q2.y), make_float2(q2.z, q2.w), make_float2(q3.x, q3.y)};
    const float2 bd = make_float2(q3.z, q3.w);
    const float2 al2 = make_float2(q4.x, q4.y), ia2 = make_float2(q4.z, q4.w);
    uint32_t ph[8];
    const uint32_t p0s = static_cast<uint32_t>(r_first & 7) << 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) ph[j] = ((p0s + static_cast<uint32_t>(((j * D) & 7) << 4)) & 0x70u) ^ lane_off;
    uint32_t xs[14];
    xs[0] = h0; xs[1] = h1; xs[2] = h2;
    uint8_t* ob = blk + r_first * 128;
#pragma unroll
    for (int j = 0; j < 3; ++j) xs[3 + j] = *reinterpret_cast<const uint32_t*>(ob + j * D * 128 + ph[j]);
    const int noct = (nquad + 1) >> 1;
#pragma unroll 1
    for (int qo = 0; qo < noct; ++qo) {
        const bool full = nquad - 2 * qo >= 2;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            xs[6 + k] = *reinterpret_cast<const uint32_t*>(ob + (k + 3) * D * 128 + ph[(k + 3) & 7]);
        if (qo == noct - 1) {
            if (full) { xs[11] = t0; xs[12] = t1; xs[13] = t2; }
            else { xs[7] = t0; xs[8] = t1; xs[9] = t2; }
        }
        uint32_t o[8];
        if (full) dw_snake_bf16<8>(xs, w, bd, al2, ia2, o);
        else dw_snake_bf16<4>(xs, w, bd, al2, ia2, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k == 4 && !full) break;
            if (r_first + k * D < 128)
                *reinterpret_cast<uint32_t*>(ob + k * D * 128 + ph[k]) = o[k];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) xs[i] = xs[8 + i];
        r_first += 8 * D;
        ob += 8 * D * 128;
    }
}


}  // namespace snacb
