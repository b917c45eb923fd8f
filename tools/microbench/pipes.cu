// Pipe micro-benchmark for B200 (sm_100a): issue rate and dependent-issue latency of the instructions the chain kernel's
// prologue / epilogue are made of.  For each op: ILP independent dependency chains per warp, W warps per SM sub-partition,
// cycles per warp-instruction per SMSP.  Build: nvcc -arch=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

enum Op { HFMA2_ = 0, FFMA_ = 1, FFMA2_ = 2, MUFU_ = 3, CVT_H2F = 4, PACK_F2H = 5, FMULRZ_ = 6, IMAD_ = 7, SINF_ = 8, HFMA2_MUFU = 9 };

template <int OP, int ILP>
__global__ void k(unsigned long long* out, int iters, float seed) {
    float f[ILP]; __half2 h[ILP]; unsigned long long d[ILP]; int n[ILP];
    for (int i = 0; i < ILP; ++i) { f[i] = seed + i; h[i] = __floats2half2_rn(seed + i, seed); d[i] = 0x3f8000003f800000ull + i; n[i] = i + threadIdx.x; }
    const __half2 hw = __floats2half2_rn(seed * 0.5f, seed * 0.25f), hb = __floats2half2_rn(0.001f, 0.002f);
    const unsigned long long dw = 0x3f0000003f000000ull, db = 0x3a8000003a800000ull;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == HFMA2_) h[i] = __hfma2(h[i], hw, hb);
                if (OP == FFMA_) f[i] = fmaf(f[i], seed, 0.001f * seed);
                if (OP == FFMA2_) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(dw), "l"(db));
                if (OP == MUFU_) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                if (OP == SINF_) f[i] = __sinf(f[i]);
                if (OP == CVT_H2F) { float2 t = __half22float2(h[i]); h[i] = *reinterpret_cast<__half2*>(&t.x); f[i] += t.y; }
                if (OP == PACK_F2H) { __half2 t = __floats2half2_rn(f[i], f[i]); f[i] = __uint_as_float(*reinterpret_cast<uint32_t*>(&t)); }
                if (OP == FMULRZ_) f[i] = __fmul_rz(f[i], 0.15915494f);
                if (OP == IMAD_) n[i] = n[i] * 3 + 7;
                if (OP == HFMA2_MUFU) { h[i] = __hfma2(h[i], hw, hb); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0; for (int i = 0; i < ILP; ++i) acc += f[i] + __low2float(h[i]) + (float)d[i] + n[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.456f) out[1] = 1;
}

template <int OP, int ILP>
void run(const char* name, unsigned long long* d) {
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 2000;
        k<OP, ILP><<<148, warps * 32>>>(d, iters, 1.0001f);
        k<OP, ILP><<<148, warps * 32>>>(d, iters, 1.0001f);
        unsigned long long c = 0;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        const double per_warp = (double)c / (iters * 8.0 * ILP);            // cycles per instruction of one warp
        const double per_smsp = per_warp / (warps / 4.0);                    // cycles per warp-instruction per SMSP
        printf("%-12s ILP %2d  warps/SMSP %d : %6.2f cyc/instr/warp  %5.2f cyc/instr/SMSP\n", name, ILP, warps / 4, per_warp, per_smsp);
    }
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 64);
#define RUN(op, name) run<op, 1>(name, d); run<op, 4>(name, d); run<op, 8>(name, d);
    RUN(HFMA2_, "HFMA2") RUN(FFMA_, "FFMA") RUN(FFMA2_, "FFMA2") RUN(MUFU_, "MUFU.EX2") RUN(SINF_, "FMUL.RZ+SIN")
    RUN(CVT_H2F, "HADD2.F32x2") RUN(PACK_F2H, "F2FP") RUN(FMULRZ_, "FMUL.RZ") RUN(IMAD_, "IMAD") RUN(HFMA2_MUFU, "HFMA2+MUFU")
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
