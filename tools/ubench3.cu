// ubench3.cu -- issue cost of the Lanczos kernels' building blocks BY OPERAND FORM (sm_100a), round 2.
// ubench2.cu measured 3-register forms only and some of its cases were loop-invariant (ptxas hoisted them).
// Here every op reads a value that changes every iteration (a per-iteration XOR/ADD on the inputs is part of
// the baseline `base8` and subtracted), and each form is listed separately: immediate / constant-bank /
// uniform-register operands, predicate destinations, three-input min/max, conversions, LDS/STS widths, and the
// legacy tensor path (mma.sync s8 and f16) alone and interleaved with FFMA -- the evidence behind DESIGN.md's
// "floor of the scalar pipes" section.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench3 tools/ubench3.cu
// Output: issue cycles per warp-instruction with 8 warps per SM sub-partition resident.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITER = 2048;

struct Regs {
    float f[32];
    uint32_t u[32];
};
struct KArgs {
    float w[8];
    uint32_t k[8];
};

template <class Body>
__global__ void __launch_bounds__(1024, 1) bench_kernel(const float *fin, const uint32_t *uin, float *fout, long long *cycles, const __grid_constant__ KArgs ka, Body body) {
    Regs r;
#pragma unroll
    for (int i = 0; i < 32; i++) r.f[i] = fin[(threadIdx.x + i * 37) & 1023];
#pragma unroll
    for (int i = 0; i < 32; i++) r.u[i] = uin[(threadIdx.x + i * 41) & 1023];
    __shared__ __align__(16) uint32_t smem[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) smem[i] = uin[i & 1023];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) body(r, smem, ka, it);
    __syncthreads();
    const long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) acc += r.f[i];
#pragma unroll
    for (int i = 0; i < 32; i++) acc += __uint_as_float(r.u[i] & 0x3fffffff);
    fout[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Register roles: f[0..15] accumulators, f[16..23] operands, u[0..15] accumulators, u[16..23] operands.
#define REP8(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)

// ---- FMA-pipe forms ----
#define FFMA_RRR(i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r.f[i]) : "f"(r.f[16 + ((i) & 3)]), "f"(r.f[20 + ((i) & 3)]));
#define FFMA_RCR(i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r.f[i]) : "f"(r.f[16 + ((i) & 3)]), "f"(ka.w[(i) & 7]));
#define FFMA_RIR(i) asm volatile("fma.rn.f32 %0, %1, 0f3F000123, %0;" : "+f"(r.f[i]) : "f"(r.f[16 + ((i) & 3)]));
#define FFMA_B(i)   asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r.f[8 + (i)]) : "f"(r.f[16 + ((i) & 3)]), "f"(ka.w[(i) & 7]));
#define FADD_RR(i)  asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r.f[i]) : "f"(r.f[16 + ((i) & 3)]));
#define FMUL_RC(i)  asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r.f[i]) : "f"(ka.w[(i) & 7]));
#define FFMA2_RCR(i) { float2 a = make_float2(r.f[2 * ((i) & 3)], r.f[2 * ((i) & 3) + 1]);                                  \
        a = __ffma2_rn(make_float2(r.f[16 + 2 * ((i) & 1)], r.f[17 + 2 * ((i) & 1)]), make_float2(ka.w[(i) & 7], ka.w[(i) & 7]), a); \
        asm volatile("" : "+f"(a.x), "+f"(a.y)); r.f[2 * ((i) & 3)] = a.x; r.f[2 * ((i) & 3) + 1] = a.y; }
#define FHADD_R(i)  asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, lo, %0;\n\t}" : "+f"(r.f[i]) : "r"(r.u[i]));
#define FHADD_Z(i)  asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, hi, 0f00000000;\n\t}" : "=f"(r.f[i]) : "r"(r.u[i]));
#define HFMA2_RRR(i) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]), "r"(r.u[20 + ((i) & 3)]));
#define HFMA2_RIR(i) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]), "r"(0xB600B600u));
#define HFMA2_RIR2(i) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]), "r"(0xB600B600u));
#define HADD2_RR(i) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]));
#define HMNMX2_RR(i) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]));
#define IMAD_RRR(i) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]), "r"(r.u[20 + ((i) & 3)]));
#define IMAD_RIR(i) asm volatile("mad.lo.s32 %0, %1, 3, %0;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]));
#define IMAD_SHL(i) asm volatile("mad.lo.s32 %0, %0, 256, %1;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]));
#define FSETP_SEL(i) asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %0, %1;\n\tselp.f32 %0, %2, %0, p;\n\t}" : "+f"(r.f[i]) : "f"(r.f[16 + ((i) & 3)]), "f"(r.f[20 + ((i) & 3)]));
#define FSETP_ACC(i) asm volatile("setp.lt.or.f32 pacc, %0, %1, pacc;" :: "f"(r.f[i]), "f"(r.f[16 + ((i) & 3)]));
#define ISETP_ACC(i) asm volatile("setp.ne.or.u32 pacc, %0, %1, pacc;" :: "r"(r.u[i]), "r"(r.u[16 + ((i) & 3)]));
#define FMNMX_RR(i) asm volatile("min.f32 %0, %0, %1;" : "+f"(r.f[i]) : "f"(r.f[16 + ((i) & 3)]));
#ifdef UB3_MNMX3
#define FMNMX3(i)   asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(r.f[i]) : "f"(r.f[16 + ((i) & 3)]), "f"(r.f[20 + ((i) & 3)]));
#endif

// ---- ALU-pipe forms ----
#define LOP3_RRR(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]), "r"(r.u[20 + ((i) & 3)]));
#define LOP3_RRI(i) asm volatile("lop3.b32 %0, %0, %1, 0x80008000, 0xf8;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]));
#define LOP2_RR(i)  asm volatile("xor.b32 %0, %0, %1;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]));
#define LOP2_RI(i)  asm volatile("xor.b32 %0, %0, 0x00ff00ff;" : "+r"(r.u[i]));
#define PRMT_RIZ(i) asm volatile("prmt.b32 %0, %0, 0, 0x4341;" : "+r"(r.u[i]));
#define PRMT_RRI(i) asm volatile("prmt.b32 %0, %0, %1, 0x2541;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]));
#define PRMT_RRR(i) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]), "r"(r.u[20 + ((i) & 3)]));
#define IADD_RR(i)  asm volatile("add.s32 %0, %0, %1;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]));
#define IADD_RI(i)  asm volatile("add.s32 %0, %0, 0x1234;" : "+r"(r.u[i]));
#define IADD3_RRR(i) asm volatile("{\n\t.reg .u32 t;\n\tadd.s32 t, %1, %2;\n\tadd.s32 %0, %0, t;\n\t}" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]), "r"(r.u[20 + ((i) & 3)]));
#define SHF_RI(i)   asm volatile("shr.u32 %0, %0, 1;" : "+r"(r.u[i]));
#define ISETP_SEL(i) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, %1;\n\tselp.b32 %0, %0, %2, p;\n\t}" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]), "r"(r.u[20 + ((i) & 3)]));
#define F2IP_RRR(i) { int a_, b_; uint32_t q_;                                                      \
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(a_) : "f"(r.f[i]));                           \
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(b_) : "f"(r.f[8 + (i)]));                     \
        asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(q_) : "r"(b_), "r"(a_), "r"(r.u[i])); r.u[i] = q_; }
#define F2I_I2FP(i) asm volatile("{\n\t.reg .s32 t;\n\tcvt.rzi.s32.f32 t, %0;\n\tcvt.rn.f32.s32 %0, t;\n\t}" : "+f"(r.f[i]));
#define F2IP_FHADD(i) { int a_, b_; uint32_t q_;                                                    \
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(a_) : "f"(r.f[i]));                           \
        asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(b_) : "f"(r.f[8 + (i)]));                     \
        asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(q_) : "r"(b_), "r"(a_));       \
        asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, lo, %0;\n\t}" : "+f"(r.f[i]) : "r"(q_)); }
#define F2I_I2FU8(i) asm volatile("{\n\t.reg .s32 t;\n\t.reg .b16 h;\n\tcvt.rzi.s32.f32 t, %0;\n\tcvt.u16.u32 h, t;\n\tcvt.rn.f32.u8 %0, h;\n\t}" : "+f"(r.f[i]));
#define F2FP_FHADD(i) { uint32_t q_; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(q_) : "f"(r.f[i]), "f"(r.f[8 + (i)]));   \
        asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, lo, %0;\n\t}" : "+f"(r.f[i]) : "r"(q_)); }
#define FRND_R(i)   asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(r.f[i]));
#define DP4A_RRR(i) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(r.u[i]) : "r"(r.u[16 + ((i) & 3)]), "r"(r.u[20 + ((i) & 3)]));
#define SHFL_R(i)   asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(r.u[i]));

// Every measured op is loop-carried through its own destination (acc = op(acc, ...)), so ptxas cannot hoist it.
// Two cheap ops per iteration perturb u[16] / f[16] for the few cases that compare against them; they and the
// loop overhead are measured alone as `base8` and subtracted.
#define PERTURB                                                                                       \
    asm volatile("add.s32 %0, %0, 0x00010001;" : "+r"(r.u[16]));                                      \
    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r.f[16]) : "f"(ka.w[0]));

#define CASE(NAME, OPS, ...)                                                                          \
    struct NAME {                                                                                     \
        static constexpr int ops = OPS;                                                               \
        static constexpr const char *name = #NAME;                                                    \
        __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const { PERTURB __VA_ARGS__ } \
    };

CASE(base8, 0, )
CASE(ffma_rrr, 16, REP8(FFMA_RRR) REP8(FFMA_RRR))
CASE(ffma_rcr, 16, REP8(FFMA_RCR) REP8(FFMA_RCR))
CASE(ffma_rir, 16, REP8(FFMA_RIR) REP8(FFMA_RIR))
CASE(fadd_rr, 16, REP8(FADD_RR) REP8(FADD_RR))
CASE(fmul_rc, 16, REP8(FMUL_RC) REP8(FMUL_RC))
CASE(ffma2_rcr, 8, REP8(FFMA2_RCR))
CASE(fhadd_r, 16, REP8(FHADD_R) REP8(FHADD_R))
CASE(hfma2_rrr, 16, REP8(HFMA2_RRR) REP8(HFMA2_RRR))
CASE(hfma2_rir, 16, REP8(HFMA2_RIR) REP8(HFMA2_RIR))
CASE(hfma2_rir_newdst, 8, REP8(HFMA2_RIR2))
CASE(hadd2_rr, 8, REP8(HADD2_RR))
CASE(hmnmx2_rr, 8, REP8(HMNMX2_RR))
CASE(imad_rrr, 16, REP8(IMAD_RRR) REP8(IMAD_RRR))
CASE(imad_rir, 16, REP8(IMAD_RIR) REP8(IMAD_RIR))
CASE(imad_shl, 16, REP8(IMAD_SHL) REP8(IMAD_SHL))
CASE(fsetp_fsel_pair, 16, REP8(FSETP_SEL))
CASE(isetp_sel_pair, 16, REP8(ISETP_SEL))
CASE(fmnmx_rr, 8, REP8(FMNMX_RR))
#ifdef UB3_MNMX3
CASE(fmnmx3, 16, REP8(FMNMX3) REP8(FMNMX3))
#endif
CASE(lop3_rrr, 16, REP8(LOP3_RRR) REP8(LOP3_RRR))
CASE(lop3_rri, 16, REP8(LOP3_RRI) REP8(LOP3_RRI))
CASE(lop2_rr, 16, REP8(LOP2_RR) REP8(LOP2_RR))
CASE(lop2_ri, 16, REP8(LOP2_RI) REP8(LOP2_RI))
CASE(prmt_riz, 8, REP8(PRMT_RIZ))
CASE(prmt_rri, 16, REP8(PRMT_RRI) REP8(PRMT_RRI))
CASE(prmt_rrr, 16, REP8(PRMT_RRR) REP8(PRMT_RRR))
CASE(iadd_rr, 16, REP8(IADD_RR) REP8(IADD_RR))
CASE(iadd_ri, 8, REP8(IADD_RI))
CASE(iadd3_rrr, 16, REP8(IADD3_RRR) REP8(IADD3_RRR))
CASE(shf_ri, 8, REP8(SHF_RI))
CASE(f2ip_rrr, 8, REP8(F2IP_RRR))
CASE(f2i_i2fp_pair, 16, REP8(F2I_I2FP))
CASE(f2ip_fhadd_pair, 16, REP8(F2IP_FHADD))
CASE(f2i_i2fu8_pair, 16, REP8(F2I_I2FU8))
CASE(f2fp_fhadd_pair, 12, REP8(F2FP_FHADD))
CASE(frnd_r, 8, REP8(FRND_R))
CASE(dp4a_rrr, 16, REP8(DP4A_RRR) REP8(DP4A_RRR))
CASE(shfl_r, 8, REP8(SHFL_R))
// mixes: do the two streams overlap (max) or add (sum)?
CASE(ffma16_lop3rrr8, 24, REP8(FFMA_RCR) REP8(LOP3_RRR) REP8(FFMA_B))
CASE(ffma16_lop2ri8, 24, REP8(FFMA_RCR) REP8(LOP2_RI) REP8(FFMA_B))
CASE(ffma16_prmtriz8, 24, REP8(FFMA_RCR) REP8(PRMT_RIZ) REP8(FFMA_B))
CASE(ffma16_hfma2rir8, 24, REP8(FFMA_RCR) REP8(HFMA2_RIR) REP8(FFMA_B))
CASE(ffma16_f2ip8, 24, REP8(FFMA_RCR) REP8(F2IP_RRR) REP8(FFMA_B))
CASE(ffma16_fhadd8, 24, REP8(FFMA_RCR) REP8(FHADD_R) REP8(FFMA_B))
CASE(ffma16_imadrir8, 24, REP8(FFMA_RCR) REP8(IMAD_RIR) REP8(FFMA_B))
CASE(ffma16_shfl8, 24, REP8(FFMA_RCR) REP8(SHFL_R) REP8(FFMA_B))
CASE(lop3rrr8_hfma2rir8, 16, REP8(LOP3_RRR) REP8(HFMA2_RIR))
CASE(prmtriz8_hfma2rir8, 16, REP8(PRMT_RIZ) REP8(HFMA2_RIR))

// predicate accumulation (setp.X.or p, a, b, p): 8 per iteration into one predicate, consumed once
struct fsetp_acc8 {
    static constexpr int ops = 8;
    static constexpr const char *name = "fsetp_acc8";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
        asm volatile("{\n\t.reg .pred pacc;\n\tsetp.lt.f32 pacc, %1, %2;\n\t"
                     "setp.lt.or.f32 pacc, %3, %2, pacc;\n\tsetp.lt.or.f32 pacc, %4, %2, pacc;\n\tsetp.lt.or.f32 pacc, %5, %2, pacc;\n\t"
                     "setp.lt.or.f32 pacc, %6, %2, pacc;\n\tsetp.lt.or.f32 pacc, %7, %2, pacc;\n\tsetp.lt.or.f32 pacc, %8, %2, pacc;\n\t"
                     "setp.lt.or.f32 pacc, %9, %2, pacc;\n\t@pacc add.s32 %0, %0, 1;\n\t}"
                     : "+r"(r.u[0]) : "f"(r.f[0]), "f"(r.f[16]), "f"(r.f[1]), "f"(r.f[2]), "f"(r.f[3]), "f"(r.f[4]), "f"(r.f[5]), "f"(r.f[6]), "f"(r.f[7]));
    }
};
struct isetp_acc8 {
    static constexpr int ops = 8;
    static constexpr const char *name = "isetp_acc8";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
        asm volatile("{\n\t.reg .pred pacc;\n\tsetp.ne.u32 pacc, %1, %2;\n\t"
                     "setp.ne.or.u32 pacc, %3, %2, pacc;\n\tsetp.ne.or.u32 pacc, %4, %2, pacc;\n\tsetp.ne.or.u32 pacc, %5, %2, pacc;\n\t"
                     "setp.ne.or.u32 pacc, %6, %2, pacc;\n\tsetp.ne.or.u32 pacc, %7, %2, pacc;\n\tsetp.ne.or.u32 pacc, %8, %2, pacc;\n\t"
                     "setp.ne.or.u32 pacc, %9, %2, pacc;\n\t@pacc add.s32 %0, %0, 1;\n\t}"
                     : "+r"(r.u[0]) : "r"(r.u[1]), "r"(r.u[16]), "r"(r.u[2]), "r"(r.u[3]), "r"(r.u[4]), "r"(r.u[5]), "r"(r.u[6]), "r"(r.u[7]), "r"(r.u[8]));
    }
};

// shared memory: addresses depend on the iteration (no hoisting), conflict-free
struct lds32x4 {
    static constexpr int ops = 4; static constexpr const char *name = "lds32x4";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
#pragma unroll
        for (int i = 0; i < 4; i++) r.u[i] ^= smem[(threadIdx.x + i * 1024 + it * 32) & 8191];
    }
};
struct lds64x4 {
    static constexpr int ops = 4; static constexpr const char *name = "lds64x4";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint2 v = *reinterpret_cast<const uint2 *>(&smem[(threadIdx.x * 2 + i * 2048 + it * 64) & 8190]); r.u[2 * i] ^= v.x; r.u[2 * i + 1] ^= v.y; }
    }
};
struct lds128x4 {
    static constexpr int ops = 4; static constexpr const char *name = "lds128x4";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint4 v = *reinterpret_cast<const uint4 *>(&smem[(threadIdx.x * 4 + i * 4096 + it * 128) & 8188]); r.u[4 * i] ^= v.x; r.u[4 * i + 1] ^= v.y; r.u[4 * i + 2] ^= v.z; r.u[4 * i + 3] ^= v.w; }
    }
};
struct sts128x2 {
    static constexpr int ops = 2; static constexpr const char *name = "sts128x2";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
#pragma unroll
        for (int i = 0; i < 2; i++) *reinterpret_cast<uint4 *>(&smem[(threadIdx.x * 4 + i * 4096 + it * 128) & 8188]) = make_uint4(r.u[16], r.u[17], r.u[20], r.u[21]);
    }
};
struct ffma16_lds64x4 {
    static constexpr int ops = 20; static constexpr const char *name = "ffma16_lds64x4";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
        REP8(FFMA_RCR)
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint2 v = *reinterpret_cast<const uint2 *>(&smem[(threadIdx.x * 2 + i * 2048 + it * 64) & 8190]); r.u[2 * i] ^= v.x; r.u[2 * i + 1] ^= v.y; }
        REP8(FFMA_B)
    }
};

// legacy tensor path: mma.sync m16n8k32 u8 x s8 -> s32 and m16n8k16 f16 -> f32, 4 independent accumulator tiles
__device__ __forceinline__ void imma(uint32_t (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void hmma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
struct imma4 {
    static constexpr int ops = 4; static constexpr const char *name = "imma_m16n8k32_x4";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
        uint32_t a[4] = {r.u[16], r.u[17], r.u[20], r.u[21]}, b[2] = {r.u[18], r.u[19]};
#pragma unroll
        for (int t = 0; t < 4; t++) { uint32_t c[4] = {r.u[4 * t], r.u[4 * t + 1], r.u[4 * t + 2], r.u[4 * t + 3]}; imma(c, a, b); r.u[4 * t] = c[0]; r.u[4 * t + 1] = c[1]; r.u[4 * t + 2] = c[2]; r.u[4 * t + 3] = c[3]; }
    }
};
struct hmma4 {
    static constexpr int ops = 4; static constexpr const char *name = "hmma_m16n8k16_x4";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
        uint32_t a[4] = {r.u[16] & 0x3bff3bffu, r.u[17] & 0x3bff3bffu, r.u[20] & 0x3bff3bffu, r.u[21] & 0x3bff3bffu}, b[2] = {r.u[18] & 0x3bff3bffu, r.u[19] & 0x3bff3bffu};
#pragma unroll
        for (int t = 0; t < 4; t++) { float c[4] = {r.f[4 * t], r.f[4 * t + 1], r.f[4 * t + 2], r.f[4 * t + 3]}; hmma(c, a, b); r.f[4 * t] = c[0]; r.f[4 * t + 1] = c[1]; r.f[4 * t + 2] = c[2]; r.f[4 * t + 3] = c[3]; }
    }
};
struct imma4_ffma16 {
    static constexpr int ops = 20; static constexpr const char *name = "imma_x4+ffma16";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
        uint32_t a[4] = {r.u[16], r.u[17], r.u[20], r.u[21]}, b[2] = {r.u[18], r.u[19]};
#pragma unroll
        for (int t = 0; t < 4; t++) {
            uint32_t c[4] = {r.u[4 * t], r.u[4 * t + 1], r.u[4 * t + 2], r.u[4 * t + 3]}; imma(c, a, b); r.u[4 * t] = c[0]; r.u[4 * t + 1] = c[1]; r.u[4 * t + 2] = c[2]; r.u[4 * t + 3] = c[3];
            FFMA_B(0) FFMA_B(1) FFMA_B(2) FFMA_B(3)
        }
    }
};
struct hmma4_ffma16 {
    static constexpr int ops = 20; static constexpr const char *name = "hmma_x4+ffma16";
    __device__ __forceinline__ void operator()(Regs &r, uint32_t *smem, const KArgs &ka, int it) const {
        PERTURB
        uint32_t a[4] = {r.u[16] & 0x3bff3bffu, r.u[17] & 0x3bff3bffu, r.u[20] & 0x3bff3bffu, r.u[21] & 0x3bff3bffu}, b[2] = {r.u[18] & 0x3bff3bffu, r.u[19] & 0x3bff3bffu};
#pragma unroll
        for (int t = 0; t < 4; t++) {
            float c[4] = {r.f[4 * t], r.f[4 * t + 1], r.f[4 * t + 2], r.f[4 * t + 3]}; hmma(c, a, b); r.f[4 * t] = c[0]; r.f[4 * t + 1] = c[1]; r.f[4 * t + 2] = c[2]; r.f[4 * t + 3] = c[3];
            asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r.f[24 + 0]) : "f"(r.f[16]), "f"(ka.w[0]));
            asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r.f[24 + 1]) : "f"(r.f[17]), "f"(ka.w[1]));
            asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r.f[24 + 2]) : "f"(r.f[20]), "f"(ka.w[2]));
            asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r.f[24 + 3]) : "f"(r.f[21]), "f"(ka.w[3]));
        }
    }
};

static double g_base = 0.0;

template <class Body>
int run(const float *fin, const uint32_t *uin, float *fout, long long *cyc, int sms, const KArgs &ka) {
    bench_kernel<<<sms, 1024>>>(fin, uin, fout, cyc, ka, Body());
    CK(cudaDeviceSynchronize());
    bench_kernel<<<sms, 1024>>>(fin, uin, fout, cyc, ka, Body());
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    const double c = (double)h[sms / 2];
    const double per_iter = c / ITER / 8.0;    // 32 warps = 8 per sub-partition
    if (Body::ops == 0) { g_base = per_iter; printf("%-24s cycles/warp-iter %7.2f  (2 perturbation ops + loop, subtracted below)\n", Body::name, per_iter); return 0; }
    printf("%-24s ops/iter %3d   cycles/warp-iter %7.2f   net cycles/op %5.2f\n", Body::name, Body::ops, per_iter, (per_iter - g_base) / Body::ops);
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
    KArgs ka;
    for (int i = 0; i < 8; i++) { ka.w[i] = 0.999f - 0.0001f * i; ka.k[i] = 0x01010101u * (i + 1); }
#define RUN(B) if (run<B>(fin, uin, fout, cyc, sms, ka)) return 1;
    RUN(base8)
    RUN(ffma_rrr) RUN(ffma_rcr) RUN(ffma_rir) RUN(fadd_rr) RUN(fmul_rc) RUN(ffma2_rcr) RUN(fhadd_r)
    RUN(hfma2_rrr) RUN(hfma2_rir) RUN(hfma2_rir_newdst) RUN(hadd2_rr) RUN(hmnmx2_rr)
    RUN(imad_rrr) RUN(imad_rir) RUN(imad_shl)
    RUN(fsetp_fsel_pair) RUN(isetp_sel_pair) RUN(fsetp_acc8) RUN(isetp_acc8) RUN(fmnmx_rr)
#ifdef UB3_MNMX3
    RUN(fmnmx3)
#endif
    RUN(lop3_rrr) RUN(lop3_rri) RUN(lop2_rr) RUN(lop2_ri) RUN(prmt_riz) RUN(prmt_rri) RUN(prmt_rrr)
    RUN(iadd_rr) RUN(iadd_ri) RUN(iadd3_rrr) RUN(shf_ri)
    RUN(f2ip_rrr) RUN(f2i_i2fp_pair) RUN(f2ip_fhadd_pair) RUN(f2i_i2fu8_pair) RUN(f2fp_fhadd_pair) RUN(frnd_r) RUN(dp4a_rrr) RUN(shfl_r)
    RUN(lds32x4) RUN(lds64x4) RUN(lds128x4) RUN(sts128x2) RUN(ffma16_lds64x4)
    RUN(ffma16_lop3rrr8) RUN(ffma16_lop2ri8) RUN(ffma16_prmtriz8) RUN(ffma16_hfma2rir8) RUN(ffma16_f2ip8) RUN(ffma16_fhadd8)
    RUN(ffma16_imadrir8) RUN(ffma16_shfl8) RUN(lop3rrr8_hfma2rir8) RUN(prmtriz8_hfma2rir8)
    RUN(imma4) RUN(hmma4) RUN(imma4_ffma16) RUN(hmma4_ffma16)
    return 0;
}
