// ubench2.cu -- which PIPE do the building blocks of the Lanczos kernels run on (sm_100a)?
// Each case is a list of inline-PTX ops; a case is timed alone and mixed with an FFMA/FFMA2 stream.
// If t(mix) ~ max(t(a), t(b)) the two streams use different pipes; if ~ t(a)+t(b) they share one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench2 tools/ubench2.cu && /tmp/ubench2
// Output: cycles per iteration per SM sub-partition warp slot (8 warps per sub-partition resident),
// i.e. issue cycles consumed per warp-iteration.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITER = 1024;

struct Regs {
    float f[24];
    uint32_t u[16];
    double d[4];
};

template <class Body>
__global__ void __launch_bounds__(1024, 1) bench_kernel(const float *fin, const uint32_t *uin, float *fout, long long *cycles, Body body) {
    Regs r;
#pragma unroll
    for (int i = 0; i < 24; i++) r.f[i] = fin[(threadIdx.x + i * 37) & 1023];
#pragma unroll
    for (int i = 0; i < 16; i++) r.u[i] = uin[(threadIdx.x + i * 41) & 1023];
#pragma unroll
    for (int i = 0; i < 4; i++) r.d[i] = (double)fin[(threadIdx.x + i * 11) & 1023];
    __shared__ __align__(16) uint32_t smem[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) smem[i] = uin[i & 1023];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) body(r, smem);
    __syncthreads();
    const long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 24; i++) acc += r.f[i];
#pragma unroll
    for (int i = 0; i < 16; i++) acc += __uint_as_float(r.u[i] & 0x3fffffff);
#pragma unroll
    for (int i = 0; i < 4; i++) acc += (float)r.d[i];
    fout[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- primitive streams (each: 8 independent ops on disjoint registers) ----
__device__ __forceinline__ void s_ffma8(Regs &r, int o) {   // 8 FFMA, accumulators f[o..o+7], x = f[16..], w = f[23]
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r.f[o + i]) : "f"(r.f[16 + (i & 3)]), "f"(r.f[23]));
}
__device__ __forceinline__ void s_ffma2x4(Regs &r, int o) {  // 4 FFMA2 = 8 fma
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float2 a = make_float2(r.f[o + 2 * i], r.f[o + 2 * i + 1]);
        a = __ffma2_rn(make_float2(r.f[16 + 2 * (i & 1)], r.f[17 + 2 * (i & 1)]), make_float2(r.f[22], r.f[23]), a);
        r.f[o + 2 * i] = a.x; r.f[o + 2 * i + 1] = a.y;
    }
}
__device__ __forceinline__ void s_f2ip4(Regs &r) {   // 4 F2IP (8 floats -> 2 words)
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int a, b, c, d;
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(a) : "f"(r.f[16 + 4 * i]));
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(b) : "f"(r.f[17 + 4 * i]));
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(c) : "f"(r.f[18 + 4 * i]));
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(d) : "f"(r.f[19 + 4 * i]));
        uint32_t hi, q;
        asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(d), "r"(c));
        asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(q) : "r"(b), "r"(a), "r"(hi));
        r.u[8 + i] ^= q;
    }
}
__device__ __forceinline__ void s_fhadd8(Regs &r) {  // 8 FHADD (f16 -> f32 widening add), results into f[8..15]
#pragma unroll
    for (int i = 0; i < 8; i++)
        asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, lo, %2;\n\t}" : "=f"(r.f[8 + i]) : "r"(r.u[i & 7]), "f"(r.f[8 + i]));
}
__device__ __forceinline__ void s_lop8(Regs &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r.u[i]) : "r"(r.u[8 + (i & 3)]), "r"(r.u[12 + (i & 3)]));
}
__device__ __forceinline__ void s_prmt8(Regs &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("prmt.b32 %0, %0, %1, 0x2541;" : "+r"(r.u[i]) : "r"(r.u[8 + (i & 3)]));
}
__device__ __forceinline__ void s_i2fp8(Regs &r) {   // 8 u32 -> f32 conversions
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(r.f[8 + i]) : "r"(r.u[i]));
}
__device__ __forceinline__ void s_i2f_u16_8(Regs &r) {   // 8 u16 -> f32 conversions (low half)
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tcvt.rn.f32.u16 %0, lo;\n\t}" : "=f"(r.f[8 + i]) : "r"(r.u[i]));
}
__device__ __forceinline__ void s_hfma2_8(Regs &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(r.u[i]) : "r"(r.u[8 + (i & 3)]), "r"(r.u[12]));
}
__device__ __forceinline__ void s_hmnmx2_8(Regs &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r.u[i]) : "r"(r.u[8 + (i & 3)]));
}
__device__ __forceinline__ void s_fmnmx8(Regs &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("min.f32 %0, %0, %1;" : "+f"(r.f[8 + i]) : "f"(r.f[16 + (i & 3)]));
}
__device__ __forceinline__ void s_imad8(Regs &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r.u[i]) : "r"(r.u[8 + (i & 3)]), "r"(r.u[12 + (i & 3)]));
}
__device__ __forceinline__ void s_iadd8(Regs &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("add.s32 %0, %0, %1;" : "+r"(r.u[i]) : "r"(r.u[8 + (i & 3)]));
}
__device__ __forceinline__ void s_isetp_sel8(Regs &r) {   // setp + predicated or (the flag pattern)
#pragma unroll
    for (int i = 0; i < 8; i++)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, %2;\n\t@p or.b32 %0, %0, 0x10;\n\t}" : "+r"(r.u[i]) : "r"(r.u[8 + (i & 3)]), "r"(r.u[12 + (i & 3)]));
}
__device__ __forceinline__ void s_dfma4(Regs &r) {
#pragma unroll
    for (int i = 0; i < 4; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(r.d[i]) : "d"(r.d[(i + 1) & 3]), "d"(r.d[(i + 2) & 3]));
}
__device__ __forceinline__ void s_lds64x4(Regs &r, uint32_t *smem) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint2 v = *reinterpret_cast<uint2 *>(&smem[(threadIdx.x * 2 + i * 2048) & 8190]);
        r.u[2 * i] ^= v.x; r.u[2 * i + 1] ^= v.y;
    }
}
__device__ __forceinline__ void s_lds32x4(Regs &r, uint32_t *smem) {
#pragma unroll
    for (int i = 0; i < 4; i++) r.u[i] ^= smem[(threadIdx.x + i * 1024) & 8191];
}
__device__ __forceinline__ void s_fadd2x4(Regs &r) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float2 a = make_float2(r.f[8 + 2 * i], r.f[9 + 2 * i]);
        a = __fadd2_rn(a, make_float2(r.f[22], r.f[23]));
        r.f[8 + 2 * i] = a.x; r.f[9 + 2 * i] = a.y;
    }
}


// ---- finely interleaved streams: does a half-rate op block the issue port in its second cycle? ----
#define FFMA_I(i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r.f[i]) : "f"(r.f[16 + ((i) & 3)]), "f"(r.f[23]));
#define LOP_I(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r.u[i]) : "r"(r.u[8 + ((i) & 3)]), "r"(r.u[12 + ((i) & 3)]));
#define PRMT_I(i) asm volatile("prmt.b32 %0, %0, %1, 0x2541;" : "+r"(r.u[i]) : "r"(r.u[8 + ((i) & 3)]));
#define FHADD_I(i) asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, lo, %0;\n\t}" : "+f"(r.f[8 + (i)]) : "r"(r.u[(i) & 7]));
#define DFMA_I(i) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(r.d[i]) : "d"(r.d[((i) + 1) & 3]), "d"(r.d[((i) + 2) & 3]));
#define FFMA2_I(i) { float2 a = make_float2(r.f[2 * (i)], r.f[2 * (i) + 1]); \
        a = __ffma2_rn(make_float2(r.f[16 + 2 * ((i) & 1)], r.f[17 + 2 * ((i) & 1)]), make_float2(r.f[22], r.f[23]), a); \
        asm volatile("" : "+f"(a.x), "+f"(a.y)); r.f[2 * (i)] = a.x; r.f[2 * (i) + 1] = a.y; }
#define F2IP_I(i) { int a_, b_; uint32_t q_; \
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(a_) : "f"(r.f[16 + (i)])); \
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(b_) : "f"(r.f[17 + (i)])); \
        asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(q_) : "r"(b_), "r"(a_), "r"(r.u[8 + (i)])); r.u[i] ^= q_; }
#define I2FP_I(i) asm volatile("{\n\t.reg .f32 t;\n\tcvt.rn.f32.u32 t, %1;\n\tadd.f32 %0, %0, t;\n\t}" : "+f"(r.f[8 + (i)]) : "r"(r.u[i]));

#define CASE(NAME, OPS, ...)                                                         \
    struct NAME {                                                                    \
        static constexpr int ops = OPS;                                              \
        static constexpr const char *name = #NAME;                                   \
        __device__ void operator()(Regs &r, uint32_t *smem) const { __VA_ARGS__ }    \
    };

CASE(ffma16, 16, s_ffma8(r, 0); s_ffma8(r, 8);)
CASE(ffma2x8, 8, s_ffma2x4(r, 0); s_ffma2x4(r, 8);)
CASE(f2ip4, 4, s_f2ip4(r);)
CASE(ffma16_f2ip4, 20, s_ffma8(r, 0); s_f2ip4(r); s_ffma8(r, 8);)
CASE(ffma8_f2ip4, 12, s_ffma8(r, 0); s_f2ip4(r);)
CASE(fhadd8, 8, s_fhadd8(r);)
CASE(ffma8_fhadd8, 16, s_ffma8(r, 0); s_fhadd8(r);)
CASE(lop8, 8, s_lop8(r);)
CASE(ffma8_lop8, 16, s_ffma8(r, 0); s_lop8(r);)
CASE(ffma16_lop8, 24, s_ffma8(r, 0); s_lop8(r); s_ffma8(r, 8);)
CASE(ffma2x8_lop8, 16, s_ffma2x4(r, 0); s_lop8(r); s_ffma2x4(r, 8);)
CASE(ffma2x8_lop4_prmt4, 16, s_ffma2x4(r, 0); s_lop8(r); s_ffma2x4(r, 8);)
CASE(prmt8, 8, s_prmt8(r);)
CASE(ffma8_prmt8, 16, s_ffma8(r, 0); s_prmt8(r);)
CASE(i2fp8, 8, s_i2fp8(r);)
CASE(ffma8_i2fp8, 16, s_ffma8(r, 0); s_i2fp8(r);)
CASE(lop8_i2fp8, 16, s_lop8(r); s_i2fp8(r);)
CASE(i2f_u16_8, 8, s_i2f_u16_8(r);)
CASE(hfma2_8, 8, s_hfma2_8(r);)
CASE(ffma8_hfma2_8, 16, s_ffma8(r, 0); s_hfma2_8(r);)
CASE(hmnmx2_8, 8, s_hmnmx2_8(r);)
CASE(ffma8_hmnmx2_8, 16, s_ffma8(r, 0); s_hmnmx2_8(r);)
CASE(fmnmx8, 8, s_fmnmx8(r);)
CASE(ffma8_fmnmx8, 16, s_ffma8(r, 0); s_fmnmx8(r);)
CASE(lop8_fmnmx8, 16, s_lop8(r); s_fmnmx8(r);)
CASE(imad8, 8, s_imad8(r);)
CASE(ffma8_imad8, 16, s_ffma8(r, 0); s_imad8(r);)
CASE(iadd8, 8, s_iadd8(r);)
CASE(lop8_iadd8, 16, s_lop8(r); s_iadd8(r);)
CASE(isetp_sel8, 16, s_isetp_sel8(r);)
CASE(ffma16_isetp_sel8, 32, s_ffma8(r, 0); s_isetp_sel8(r); s_ffma8(r, 8);)
CASE(dfma4, 4, s_dfma4(r);)
CASE(ffma2x8_dfma4, 12, s_ffma2x4(r, 0); s_dfma4(r); s_ffma2x4(r, 8);)
CASE(ffma16_dfma4, 20, s_ffma8(r, 0); s_dfma4(r); s_ffma8(r, 8);)
CASE(lds64x4, 4, s_lds64x4(r, smem);)
CASE(ffma16_lds64x4, 20, s_ffma8(r, 0); s_lds64x4(r, smem); s_ffma8(r, 8);)
CASE(lds32x4, 4, s_lds32x4(r, smem);)
CASE(fadd2x4, 4, s_fadd2x4(r);)
CASE(ffma8_f2ip4_lop8, 20, s_ffma8(r, 0); s_f2ip4(r); s_lop8(r);)
CASE(f2ip4_lop8, 12, s_f2ip4(r); s_lop8(r);)
CASE(fhadd8_lop8, 16, s_fhadd8(r); s_lop8(r);)
// the V-pass shape per new row of an 8-byte column: 4 PRMT + 8 FHADD + 32 FFMA2 + 4 FADD2 + 8 F2IP + 8 LOP3
CASE(vshape, 64, s_prmt8(r); s_fhadd8(r); s_ffma2x4(r, 0); s_ffma2x4(r, 8); s_ffma2x4(r, 0); s_ffma2x4(r, 8);
     s_ffma2x4(r, 0); s_ffma2x4(r, 8); s_ffma2x4(r, 0); s_ffma2x4(r, 8); s_fadd2x4(r); s_f2ip4(r); s_f2ip4(r); s_lop8(r);)


CASE(il_ffma_lop, 16, FFMA_I(0) LOP_I(0) FFMA_I(1) LOP_I(1) FFMA_I(2) LOP_I(2) FFMA_I(3) LOP_I(3) FFMA_I(4) LOP_I(4) FFMA_I(5) LOP_I(5) FFMA_I(6) LOP_I(6) FFMA_I(7) LOP_I(7))
CASE(il_ffma2_lop, 16, FFMA2_I(0) LOP_I(0) FFMA2_I(1) LOP_I(1) FFMA2_I(2) LOP_I(2) FFMA2_I(3) LOP_I(3) FFMA2_I(4) LOP_I(4) FFMA2_I(5) LOP_I(5) FFMA2_I(6) LOP_I(6) FFMA2_I(7) LOP_I(7))
CASE(il_2ffma_lop, 24, FFMA_I(0) FFMA_I(8) LOP_I(0) FFMA_I(1) FFMA_I(9) LOP_I(1) FFMA_I(2) FFMA_I(10) LOP_I(2) FFMA_I(3) FFMA_I(11) LOP_I(3) FFMA_I(4) FFMA_I(12) LOP_I(4) FFMA_I(5) FFMA_I(13) LOP_I(5) FFMA_I(6) FFMA_I(14) LOP_I(6) FFMA_I(7) FFMA_I(15) LOP_I(7))
CASE(il_ffma2_prmt, 16, FFMA2_I(0) PRMT_I(0) FFMA2_I(1) PRMT_I(1) FFMA2_I(2) PRMT_I(2) FFMA2_I(3) PRMT_I(3) FFMA2_I(4) PRMT_I(4) FFMA2_I(5) PRMT_I(5) FFMA2_I(6) PRMT_I(6) FFMA2_I(7) PRMT_I(7))
CASE(il_ffma2_dfma, 12, FFMA2_I(0) FFMA2_I(1) DFMA_I(0) FFMA2_I(2) FFMA2_I(3) DFMA_I(1) FFMA2_I(4) FFMA2_I(5) DFMA_I(2) FFMA2_I(6) FFMA2_I(7) DFMA_I(3))
CASE(il_ffma_dfma, 12, FFMA_I(0) FFMA_I(1) DFMA_I(0) FFMA_I(2) FFMA_I(3) DFMA_I(1) FFMA_I(4) FFMA_I(5) DFMA_I(2) FFMA_I(6) FFMA_I(7) DFMA_I(3))
CASE(il_fhadd_lop, 16, FHADD_I(0) LOP_I(0) FHADD_I(1) LOP_I(1) FHADD_I(2) LOP_I(2) FHADD_I(3) LOP_I(3) FHADD_I(4) LOP_I(4) FHADD_I(5) LOP_I(5) FHADD_I(6) LOP_I(6) FHADD_I(7) LOP_I(7))
CASE(il_f2ip_lop, 8, F2IP_I(0) LOP_I(4) F2IP_I(1) LOP_I(5) F2IP_I(2) LOP_I(6) F2IP_I(3) LOP_I(7))
CASE(il_f2ip_ffma, 12, F2IP_I(0) FFMA_I(0) FFMA_I(1) F2IP_I(1) FFMA_I(2) FFMA_I(3) F2IP_I(2) FFMA_I(4) FFMA_I(5) F2IP_I(3) FFMA_I(6) FFMA_I(7))
CASE(f2ip_only4, 4, F2IP_I(0) F2IP_I(1) F2IP_I(2) F2IP_I(3))
CASE(i2fp_fadd8, 16, I2FP_I(0) I2FP_I(1) I2FP_I(2) I2FP_I(3) I2FP_I(4) I2FP_I(5) I2FP_I(6) I2FP_I(7))
CASE(il_i2fp_lop, 24, I2FP_I(0) LOP_I(0) I2FP_I(1) LOP_I(1) I2FP_I(2) LOP_I(2) I2FP_I(3) LOP_I(3) I2FP_I(4) LOP_I(4) I2FP_I(5) LOP_I(5) I2FP_I(6) LOP_I(6) I2FP_I(7) LOP_I(7))
CASE(il_ffma2_lop_ffma2_prmt_fhadd, 20, FFMA2_I(0) LOP_I(0) FFMA2_I(1) PRMT_I(1) FHADD_I(0) FFMA2_I(2) LOP_I(2) FFMA2_I(3) PRMT_I(3) FHADD_I(1) FFMA2_I(4) LOP_I(4) FFMA2_I(5) PRMT_I(5) FHADD_I(2) FFMA2_I(6) LOP_I(6) FFMA2_I(7) PRMT_I(7) FHADD_I(3))

template <class Body>
int run(const float *fin, const uint32_t *uin, float *fout, long long *cyc, int sms) {
    bench_kernel<<<sms, 1024>>>(fin, uin, fout, cyc, Body());
    CK(cudaDeviceSynchronize());
    bench_kernel<<<sms, 1024>>>(fin, uin, fout, cyc, Body());
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    const double c = (double)h[sms / 2];
    // 1024 threads = 32 warps = 8 per sub-partition; cycles per warp-iteration per sub-partition
    const double per_iter = c / ITER / 8.0;
    printf("%-24s ops/iter %3d   cycles/warp-iter %7.2f   cycles/op %5.2f\n", Body::name, Body::ops, per_iter, per_iter / Body::ops);
    return 0;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *fin, *fout; uint32_t *uin; long long *cyc;
    CK(cudaMalloc(&fin, 1024 * 4)); CK(cudaMalloc(&uin, 1024 * 4));
    CK(cudaMalloc(&fout, (size_t)sms * 1024 * 4)); CK(cudaMalloc(&cyc, sms * 8));
    std::vector<float> hf(1024); std::vector<uint32_t> hu(1024);
    for (int i = 0; i < 1024; i++) { hf[i] = 0.5f + (i % 7) * 0.01f; hu[i] = 0x00330012u + i; }
    CK(cudaMemcpy(fin, hf.data(), 4096, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(uin, hu.data(), 4096, cudaMemcpyHostToDevice));
#define RUN(B) if (run<B>(fin, uin, fout, cyc, sms)) return 1;
    RUN(ffma16) RUN(ffma2x8) RUN(fadd2x4) RUN(f2ip4) RUN(ffma8_f2ip4) RUN(ffma16_f2ip4) RUN(f2ip4_lop8) RUN(ffma8_f2ip4_lop8)
    RUN(fhadd8) RUN(ffma8_fhadd8) RUN(fhadd8_lop8)
    RUN(lop8) RUN(ffma8_lop8) RUN(ffma16_lop8) RUN(ffma2x8_lop8) RUN(prmt8) RUN(ffma8_prmt8)
    RUN(i2fp8) RUN(ffma8_i2fp8) RUN(lop8_i2fp8) RUN(i2f_u16_8)
    RUN(hfma2_8) RUN(ffma8_hfma2_8) RUN(hmnmx2_8) RUN(ffma8_hmnmx2_8)
    RUN(fmnmx8) RUN(ffma8_fmnmx8) RUN(lop8_fmnmx8) RUN(imad8) RUN(ffma8_imad8) RUN(iadd8) RUN(lop8_iadd8)
    RUN(isetp_sel8) RUN(ffma16_isetp_sel8)
    RUN(dfma4) RUN(ffma2x8_dfma4) RUN(ffma16_dfma4)
    RUN(lds64x4) RUN(ffma16_lds64x4) RUN(lds32x4)
    RUN(vshape)
    RUN(il_ffma_lop) RUN(il_ffma2_lop) RUN(il_2ffma_lop) RUN(il_ffma2_prmt) RUN(il_ffma2_dfma) RUN(il_ffma_dfma) RUN(il_fhadd_lop)
    RUN(f2ip_only4) RUN(il_f2ip_lop) RUN(il_f2ip_ffma) RUN(i2fp_fadd8) RUN(il_i2fp_lop) RUN(il_ffma2_lop_ffma2_prmt_fhadd)
    return 0;
}
