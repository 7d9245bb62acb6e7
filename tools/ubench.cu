// ubench.cu -- instruction-throughput microbenchmarks for the ops the Lanczos kernels are
// built from (sm_100a).  Prints thread-ops per clock per SM for each body, measured with
// clock64() over a resident grid (1 CTA of 1024 threads per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench
// Results are recorded in profiles/ and drive the instruction budget in DESIGN.md.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITER = 2048;

struct Regs {
    float f[16];
    uint32_t u[8];
};

template <class Body>
__global__ void __launch_bounds__(1024, 1) bench_kernel(const float *fin, const uint32_t *uin, float *fout, long long *cycles, Body body) {
    Regs r;
#pragma unroll
    for (int i = 0; i < 16; i++) r.f[i] = fin[(threadIdx.x + i * 37) & 1023];
#pragma unroll
    for (int i = 0; i < 8; i++) r.u[i] = uin[(threadIdx.x + i * 41) & 1023];
    __shared__ __align__(16) uint32_t smem[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) smem[i] = uin[i & 1023];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) body(r, smem);
    __syncthreads();
    const long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc += r.f[i];
#pragma unroll
    for (int i = 0; i < 8; i++) acc += __uint_as_float(r.u[i] & 0x3fffffff);
    fout[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- bodies (each states how many "ops" it issues per iteration per thread) ----
struct FFMA8 {  // 8 independent 3-register FFMA
    static constexpr int ops = 8; static constexpr const char *name = "FFMA (3 reg) x8";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.f[i] = fmaf(r.f[i], r.f[8 + (i & 3)], r.f[12 + (i & 3)]);
    }
};
struct FFMA8_shared_w {  // the filter pattern: acc_i += x_i * w, same w for all i
    static constexpr int ops = 8; static constexpr const char *name = "FFMA acc+=x*w (w shared) x8";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.f[i] = fmaf(r.f[8 + (i & 7)], r.f[15], r.f[i]);
    }
};
struct FFMA2x4 {  // 4 packed FFMA2 = 8 FMAs
    static constexpr int ops = 8; static constexpr const char *name = "FFMA2 x4 (=8 fma)";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float2 a = make_float2(r.f[2 * i], r.f[2 * i + 1]);
            float2 b = make_float2(r.f[8 + 2 * (i & 1)], r.f[9 + 2 * (i & 1)]);
            float2 c = make_float2(r.f[12 + 2 * (i & 1)], r.f[13 + 2 * (i & 1)]);
            a = __ffma2_rn(a, b, c);
            r.f[2 * i] = a.x; r.f[2 * i + 1] = a.y;
        }
    }
};
struct FFMA2x8 {
    static constexpr int ops = 16; static constexpr const char *name = "FFMA2 x8 (=16 fma) acc+=x*w";
    __device__ void operator()(Regs &r, uint32_t *) const {
        float2 w = make_float2(r.f[14], r.f[15]);
#pragma unroll
        for (int rep = 0; rep < 2; rep++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float2 a = make_float2(r.f[2 * i], r.f[2 * i + 1]);
            float2 x = make_float2(r.f[8 + 2 * (i & 1)], r.f[9 + 2 * (i & 1)]);
            a = __ffma2_rn(x, w, a);
            r.f[2 * i] = a.x; r.f[2 * i + 1] = a.y;
        }
    }
};
struct FADD8 {
    static constexpr int ops = 8; static constexpr const char *name = "FADD x8";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.f[i] = r.f[i] + r.f[8 + (i & 7)];
    }
};
struct FMNMX8 {
    static constexpr int ops = 8; static constexpr const char *name = "FMNMX x8";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.f[i] = fminf(r.f[i], r.f[8 + (i & 7)]);
    }
};
struct I2F_U8x8 {  // 8 byte->float conversions with byte select, consumed by 8 FADD
    static constexpr int ops = 8; static constexpr const char *name = "I2F.U8 x8 (+8 FADD consumers)";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.f[i] += (float)((r.u[i >> 2] >> (8 * (i & 3))) & 0xff);
        r.u[0] += 0x01010101u; r.u[1] += 0x01010101u;
    }
};
struct F2IPx8 {  // 8 float->u8 truncating saturating conversions, consumed by 4 IADD3
    static constexpr int ops = 8; static constexpr const char *name = "F2IP.U8.F32 x8 (+4 IADD3 consumers)";
    __device__ void operator()(Regs &r, uint32_t *) const {
        uint32_t q[8];
#pragma unroll
        for (int i = 0; i < 8; i++) asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(q[i]) : "f"(r.f[i]));
#pragma unroll
        for (int i = 0; i < 4; i++) r.u[i] = r.u[i] + q[2 * i] + q[2 * i + 1];
        r.f[0] += 1.0f;  // keep inputs changing (1 FADD)
    }
};
struct F2I_S32x8 {
    static constexpr int ops = 8; static constexpr const char *name = "F2I.S32.TRUNC x8 (+4 IADD3 consumers)";
    __device__ void operator()(Regs &r, uint32_t *) const {
        int q[8];
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = __float2int_rz(r.f[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) r.u[i] = r.u[i] + q[2 * i] + q[2 * i + 1];
        r.f[0] += 1.0f;
    }
};
struct PRMT8 {
    static constexpr int ops = 8; static constexpr const char *name = "PRMT x8";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.u[i] = __byte_perm(r.u[i], r.u[(i + 1) & 7], 0x2541);
    }
};
struct LOP8 {
    static constexpr int ops = 8; static constexpr const char *name = "LOP3 x8";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.u[i] = (r.u[i] & r.u[(i + 1) & 7]) ^ r.u[(i + 3) & 7];
    }
};
struct IADD8 {
    static constexpr int ops = 8; static constexpr const char *name = "IADD3 x8";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.u[i] = r.u[i] + r.u[(i + 1) & 7] + r.u[(i + 3) & 7];
    }
};
struct IMAD8 {
    static constexpr int ops = 8; static constexpr const char *name = "IMAD x8";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.u[i] = r.u[i] * r.u[(i + 1) & 7] + r.u[(i + 3) & 7];
    }
};
struct IMADSHL8 {  // pack step: t = t*256 + u
    static constexpr int ops = 8; static constexpr const char *name = "IMAD x*256+y x8";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.u[i] = r.u[i] * 256u + r.u[(i + 1) & 7];
    }
};
struct MIX_FFMA_PRMT {  // 8 FFMA + 4 PRMT: do they co-issue?
    static constexpr int ops = 12; static constexpr const char *name = "mix 8 FFMA + 4 PRMT";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 8; i++) r.f[i] = fmaf(r.f[8 + (i & 7)], r.f[15], r.f[i]);
#pragma unroll
        for (int i = 0; i < 4; i++) r.u[i] = __byte_perm(r.u[i], r.u[(i + 1) & 7], 0x2541);
    }
};
struct MIX_FFMA2_PRMT {  // 8 FFMA2 (16 fma) + 8 PRMT
    static constexpr int ops = 24; static constexpr const char *name = "mix 8 FFMA2(16 fma) + 8 PRMT";
    __device__ void operator()(Regs &r, uint32_t *) const {
        float2 w = make_float2(r.f[14], r.f[15]);
#pragma unroll
        for (int rep = 0; rep < 2; rep++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float2 a = make_float2(r.f[2 * i], r.f[2 * i + 1]);
            float2 x = make_float2(r.f[8 + 2 * (i & 1)], r.f[9 + 2 * (i & 1)]);
            a = __ffma2_rn(x, w, a);
            r.f[2 * i] = a.x; r.f[2 * i + 1] = a.y;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) r.u[i] = __byte_perm(r.u[i], r.u[(i + 1) & 7], 0x2541);
    }
};
struct MIX_FFMA_I2F {  // 12 FFMA + 4 I2F.U8
    static constexpr int ops = 16; static constexpr const char *name = "mix 12 FFMA + 4 I2F.U8";
    __device__ void operator()(Regs &r, uint32_t *) const {
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = (float)((r.u[0] >> (8 * i)) & 0xff);
#pragma unroll
        for (int i = 0; i < 12; i++) r.f[i & 7] = fmaf(x[i & 3], r.f[8 + (i & 7)], r.f[i & 7]);
        r.u[0] += 0x01010101u;
    }
};
struct MIX_FFMA_F2IP {  // 12 FFMA + 4 F2IP + 3 IMAD pack
    static constexpr int ops = 19; static constexpr const char *name = "mix 12 FFMA + 4 F2IP + 3 IMAD(pack)";
    __device__ void operator()(Regs &r, uint32_t *) const {
#pragma unroll
        for (int i = 0; i < 12; i++) r.f[i & 7] = fmaf(r.f[8 + (i & 7)], r.f[15], r.f[i & 7]);
        uint32_t q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(q[i]) : "f"(r.f[i]));
        r.u[0] ^= ((q[3] * 256u + q[2]) * 256u + q[1]) * 256u + q[0];
    }
};
struct LDS32 {
    static constexpr int ops = 4; static constexpr const char *name = "LDS.32 x4 (conflict-free)";
    __device__ void operator()(Regs &r, uint32_t *s) const {
#pragma unroll
        for (int i = 0; i < 4; i++) r.u[i] += s[(threadIdx.x + r.u[4 + i]) & 4095];
    }
};
struct LDS128 {
    static constexpr int ops = 2; static constexpr const char *name = "LDS.128 x2 (16 B/thread each)";
    __device__ void operator()(Regs &r, uint32_t *s) const {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            uint4 v = *reinterpret_cast<const uint4 *>(&s[((threadIdx.x + r.u[4 + i]) * 4) & 4095]);
            r.u[2 * i] += v.x ^ v.y; r.u[2 * i + 1] += v.z ^ v.w;
        }
    }
};
struct STS128 {
    static constexpr int ops = 2; static constexpr const char *name = "STS.128 x2";
    __device__ void operator()(Regs &r, uint32_t *s) const {
#pragma unroll
        for (int i = 0; i < 2; i++)
            *reinterpret_cast<uint4 *>(&s[((threadIdx.x + i * 1024) * 4) & 4095]) = make_uint4(r.u[0], r.u[1], r.u[2], r.u[3]);
        r.u[0]++;
    }
};
struct DADD4 {
    static constexpr int ops = 4; static constexpr const char *name = "DADD x4";
    __device__ void operator()(Regs &r, uint32_t *) const {
        double *d = reinterpret_cast<double *>(r.f);
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = __dadd_rn(d[i], d[4 + (i & 3)]);
    }
};
struct DMUL4 {
    static constexpr int ops = 4; static constexpr const char *name = "DMUL x4";
    __device__ void operator()(Regs &r, uint32_t *) const {
        double *d = reinterpret_cast<double *>(r.f);
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = __dmul_rn(d[i], d[4 + (i & 3)]);
    }
};
// the vertical-pass inner body for one interpolated output word (4 bytes): 24 FFMA, quantise,
// guard test (2nd quantisation at +2g), pack, and the conversions of one new input word
struct VBODY {
    static constexpr int ops = 1; static constexpr const char *name = "V-pass body / output word (24 FFMA+4 FADD+8 F2IP+6 IMAD+4 I2F)";
    __device__ void operator()(Regs &r, uint32_t *s) const {
        float x[4];
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = (float)((r.u[0] >> (8 * i)) & 0xff);
        float acc[4] = {r.f[12], r.f[12], r.f[12], r.f[12]};
#pragma unroll
        for (int k = 0; k < 6; k++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[i] = fmaf(k == 0 ? x[i] : r.f[(k * 4 + i) & 7], r.f[8 + (k & 3)], acc[i]);
        uint32_t qa[4], qb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(qa[i]) : "f"(acc[i]));
            asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(qb[i]) : "f"(acc[i] + r.f[13]));
        }
        const uint32_t wa = ((qa[3] * 256u + qa[2]) * 256u + qa[1]) * 256u + qa[0];
        const uint32_t wb = ((qb[3] * 256u + qb[2]) * 256u + qb[1]) * 256u + qb[0];
        if (wa != wb) r.u[7]++;
        r.u[0] = wa + r.u[1];
#pragma unroll
        for (int i = 0; i < 4; i++) r.f[i] = acc[i];
    }
};

template <class Body>
int run(const float *fin, const uint32_t *uin, float *fout, long long *cyc, int sms, int clock_khz) {
    const int grid = sms;
    bench_kernel<Body><<<grid, 1024>>>(fin, uin, fout, cyc, Body());  // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench_kernel<Body><<<grid, 1024>>>(fin, uin, fout, cyc, Body());
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    const double med = (double)h[grid / 2];
    const double ops = (double)1024 * ITER * Body::ops;
    printf("%-78s %8.1f ops/clk/SM   (%6.2f clk/iter/warp-slot, %.3f ms, ~%.0f MHz)\n", Body::name, ops / med,
           med / ITER, ms, med / (ms * 1e-3) / 1e6);
    return 0;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, clock %d kHz, L2 %d MB, smem/SM %zu KB\n", prop.name, prop.multiProcessorCount,
           prop.clockRate, prop.l2CacheSize >> 20, prop.sharedMemPerMultiprocessor >> 10);
    float *fin, *fout; uint32_t *uin; long long *cyc;
    CK(cudaMalloc(&fin, 1024 * 4)); CK(cudaMalloc(&uin, 1024 * 4));
    CK(cudaMalloc(&fout, (size_t)prop.multiProcessorCount * 1024 * 4)); CK(cudaMalloc(&cyc, prop.multiProcessorCount * 8));
    std::vector<float> hf(1024); std::vector<uint32_t> hu(1024);
    for (int i = 0; i < 1024; i++) { hf[i] = 1.0f + 1e-3f * (i % 17); hu[i] = 0x01020304u * (i + 1); }
    CK(cudaMemcpy(fin, hf.data(), 4096, cudaMemcpyHostToDevice)); CK(cudaMemcpy(uin, hu.data(), 4096, cudaMemcpyHostToDevice));
    const int sms = prop.multiProcessorCount, khz = prop.clockRate;
    run<FFMA8>(fin, uin, fout, cyc, sms, khz);
    run<FFMA8_shared_w>(fin, uin, fout, cyc, sms, khz);
    run<FFMA2x4>(fin, uin, fout, cyc, sms, khz);
    run<FFMA2x8>(fin, uin, fout, cyc, sms, khz);
    run<FADD8>(fin, uin, fout, cyc, sms, khz);
    run<FMNMX8>(fin, uin, fout, cyc, sms, khz);
    run<I2F_U8x8>(fin, uin, fout, cyc, sms, khz);
    run<F2IPx8>(fin, uin, fout, cyc, sms, khz);
    run<F2I_S32x8>(fin, uin, fout, cyc, sms, khz);
    run<PRMT8>(fin, uin, fout, cyc, sms, khz);
    run<LOP8>(fin, uin, fout, cyc, sms, khz);
    run<IADD8>(fin, uin, fout, cyc, sms, khz);
    run<IMAD8>(fin, uin, fout, cyc, sms, khz);
    run<IMADSHL8>(fin, uin, fout, cyc, sms, khz);
    run<MIX_FFMA_PRMT>(fin, uin, fout, cyc, sms, khz);
    run<MIX_FFMA2_PRMT>(fin, uin, fout, cyc, sms, khz);
    run<MIX_FFMA_I2F>(fin, uin, fout, cyc, sms, khz);
    run<MIX_FFMA_F2IP>(fin, uin, fout, cyc, sms, khz);
    run<LDS32>(fin, uin, fout, cyc, sms, khz);
    run<LDS128>(fin, uin, fout, cyc, sms, khz);
    run<STS128>(fin, uin, fout, cyc, sms, khz);
    run<DADD4>(fin, uin, fout, cyc, sms, khz);
    run<DMUL4>(fin, uin, fout, cyc, sms, khz);
    run<VBODY>(fin, uin, fout, cyc, sms, khz);
    return 0;
}
