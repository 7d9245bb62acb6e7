// fast_common.cuh -- device/host helpers shared by the TMA + register-window kernels
// (lanczos_v6.cu: static-phase H pass for a few ratios; lanczos_dyn.cu: any ratio with N <= 32).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace lzb {
namespace {

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

// u8 -> fp32 without the 16/clk I2F unit: PRMT places two bytes into the low bytes of two fp16
// lanes (0x00bb = the fp16 SUBNORMAL b * 2^-24), and one FHADD per byte widens it to fp32 exactly.
// All fp32 pixel values in these kernels therefore carry a factor 2^-24 (kPixScale); the float
// weight tables are pre-multiplied by 2^24 on the host (exact, powers of two), so sums come out in
// pixel units with exactly the rounding they would have unscaled.
constexpr float kPixUnscale = 16777216.f;  // 2^24
__device__ __forceinline__ float h2_lo_to_f32(uint32_t h2) {
    float f;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, lo, 0f00000000;\n\t}" : "=f"(f) : "r"(h2));
    return f;
}
__device__ __forceinline__ float h2_hi_to_f32(uint32_t h2) {
    float f;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, hi, 0f00000000;\n\t}" : "=f"(f) : "r"(h2));
    return f;
}
__device__ __forceinline__ void word_to_f32x4(uint32_t w, float &f0, float &f1, float &f2, float &f3) {
    const uint32_t a = __byte_perm(w, 0u, 0x4140), b = __byte_perm(w, 0u, 0x4342);
    f0 = h2_lo_to_f32(a); f1 = h2_hi_to_f32(a); f2 = h2_lo_to_f32(b); f3 = h2_hi_to_f32(b);
}

// double_to_uint8 (full_TB.h:29-37) on four fp32 values, packed little-endian: clamp to [0,255],
// truncate toward zero. ptxas fuses each cvt.rzi pair + cvt.pack into ONE F2IP.U8.F32.TRUNC.
__device__ __forceinline__ uint32_t quantise4(float a, float b, float c, float d) {
    const int ia = __float2int_rz(a), ib = __float2int_rz(b), ic = __float2int_rz(c), id = __float2int_rz(d);
    uint32_t hi, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(id), "r"(ic));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(ib), "r"(ia), "r"(hi));
    return r;
}

// Exact restatement of full_TB.h:58-63 for one sample whose 2a taps are `stride` bytes apart in
// shared memory (taps outside the image were zero-filled by TMA: 0*w adds +-0, same bits).
// (double)byte is formed as (2^52 + b) - 2^52 on the FP64 pipe: I2F.F64 runs on the slow XU unit.
template <int TAPS, class W>
__device__ __forceinline__ uint8_t exact_taps(const uint8_t *tap0, int stride, W weight) {
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < TAPS; k++) {
        const double v = __hiloint2double(0x43300000, (int)tap0[k * stride]) - 4503599627370496.0;
        sum = __dadd_rn(sum, __dmul_rn(v, weight(k)));
    }
    return quantise_f64(sum);
}
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(sym);
    }
    return fn;
}


}  // namespace
}  // namespace lzb
