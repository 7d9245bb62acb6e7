// lanczos_fast.cu -- specialised fused H->V Lanczos kernels for sm_100a.
//
// What it replaces in the reference (software-path arithmetic, HLS-path structure):
//   cyclic_buffer/cyclic_buffer.h:4-69  2a(+1)-line cyclic buffer   -> shared-memory ring of H-pass rows
//                                                                       + a 2a-row register window per thread
//   worker.cpp:138-155 ColWorkers::exec / :225-247 RowWorkers::exec  -> the V-pass and H-pass below
//   lanczos.cpp:68-83  process_channel block loop (DATAFLOW)          -> chunk loop, TMA double buffering
//   kernel.cpp:40-58   coefficient LUT                                -> polyphase table in kernel params
//                                                                       (H pass, static phase) and smem (V pass)
//   full_TB.h:55-77    the arithmetic that must be matched bit for bit
//
// One CTA owns a strip of output byte-columns and a vertical segment of the image and streams
// down it in chunks of RB input rows:
//   1. TMA (cp.async.bulk.tensor, 3-D map x/y/frame, out-of-bounds = 0 = the reference's dropped
//      taps) stages RB input rows + halo bytes into shared memory, double buffered;
//   2. H pass: each thread takes PH ratio-periods of one row, converts its input bytes to fp32
//      once (PRMT to fp16 magic + FHADD), runs the 2a-tap FFMA chains with the phase weights as
//      constant-bank operands, quantises 4 samples per 2 F2IP (clamp+truncate+pack) and stores
//      16-byte vectors into the shared ring of uint8 intermediate rows;
//   3. V pass: each thread owns one 32-bit word (4 byte-columns) of the strip, keeps the last 2a
//      intermediate rows as fp32 in registers (statically rotated), and for every new row emits the
//      output rows that became computable, as coalesced 32-bit stores.
// Exactness: sums are accumulated from -guard; a sample whose truncation differs between x-guard
// and x+guard (guard >= 2x the rigorous fp32 error bound, plan.cpp) is recomputed with the
// reference's double arithmetic.  Phase-0 samples are copies, checked with the "cannot flip"
// filter of plan.cpp and recomputed exactly when it fails.  Recomputations are deferred to
// per-CTA lists so that they run 32 lanes wide.
#include <algorithm>
#include <cuda.h>

#include "../../include/lanczos_b200.h"
#include "kernels.cuh"

namespace lzb {

namespace {

constexpr int F_THREADS = 256;
constexpr int F_SW_MAX = 1024;   // strip width in output bytes (one 32-bit word per V thread)
constexpr int F_LIST = 1536;     // deferred-recompute list entries per pass

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

// two bytes of `w` (selected by SEL, e.g. 0x7170 = bytes 0 and 1) as a pair of fp16 values 1024+b
template <uint32_t SEL>
__device__ __forceinline__ uint32_t bytes_to_h2(uint32_t w) { return __byte_perm(w, 0x64006400u, SEL); }
// fp32 = fp16 (low/high half of h2) - 1024 : one FHADD each, exact
__device__ __forceinline__ float h2_lo_to_f32(uint32_t h2) {
    float f;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, lo, %2;\n\t}" : "=f"(f) : "r"(h2), "f"(-1024.f));
    return f;
}
__device__ __forceinline__ float h2_hi_to_f32(uint32_t h2) {
    float f;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.f16 %0, hi, %2;\n\t}" : "=f"(f) : "r"(h2), "f"(-1024.f));
    return f;
}
__device__ __forceinline__ void word_to_f32x4(uint32_t w, float &f0, float &f1, float &f2, float &f3) {
    const uint32_t a = bytes_to_h2<0x7170>(w), b = bytes_to_h2<0x7372>(w);
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

// ---------------------------------------------------------------------------------------------
// parameters
// ---------------------------------------------------------------------------------------------
struct FastParams {
    const uint8_t *in;    // input row `in_row0` of frame 0 (exact recomputation reads it directly)
    uint8_t *out;         // output row `out_row0` of frame 0
    long long in_pitch, out_pitch, in_frame_stride, out_frame_stride;
    int in_w, in_h, out_w, out_h;
    int out_row0, out_rows, in_row0, in_rows;
    int sw;               // strip width in output bytes (groups * OUT_B)
    int groups;           // H-pass thread groups per strip row
    int seg_periods;      // vertical ratio-periods per segment
    int vperiod0;         // first vertical period covered by the launch (floor(out_row0 / N))
    const double *wdx, *wdy;
    float guard;
    int exact_x, exact_y;     // every phase-0 coordinate exactly integral (plan AxisTables.aligned_exact)
    int strict_v_identity;    // 0 with LANCZOS_FLAG_FAST_ALIGNED
    float align_k[8];         // phase-0 "cannot flip" constants
    float wtab[32 * 8];       // polyphase table [N][8] (padded to 8 taps), N <= 32
    unsigned long long *strict_counter;
};

template <int C, int A, int N, int D, int PH>
struct Geo {
    static constexpr int TAPS = 2 * A;
    static constexpr int IN_B = PH * D * C;     // input bytes owned by one H item
    static constexpr int OUT_B = PH * N * C;    // output bytes produced by one H item
    static constexpr int HALO_L = (A - 1) * C;
    static constexpr int B_LAST = ((N - 1) * D) / N;
    static constexpr int HALO_R = (B_LAST + A + 1 - D) * C > 0 ? (B_LAST + A + 1 - D) * C : 0;
    static constexpr int PAD_L = 16 * ((HALO_L + 15) / 16);       // TMA box starts PAD_L bytes left of the strip
    static constexpr int WIN_B = HALO_L + IN_B + HALO_R;          // bytes one H item reads
    static constexpr int MIS = (PAD_L - HALO_L) % 4;              // window start inside its first word
    static constexpr int WIN0 = PAD_L - HALO_L - MIS;             // first word (byte offset) for group 0
    static constexpr int NWORDS = (MIS + WIN_B + 3) / 4;
    static constexpr int MAX_GROUPS = F_SW_MAX / OUT_B;
    static constexpr int BOX_B = 16 * ((PAD_L + MAX_GROUPS * IN_B + HALO_R + 15) / 16);  // TMA box row bytes
    static constexpr int RB = 12;                                 // input rows per chunk (multiple of TAPS)
    static constexpr int RING = RB + TAPS + 2;                    // intermediate rows kept in smem
    static constexpr int STAGE_B = 128 * ((RB * BOX_B + 127) / 128);  // TMA destinations must be 128-byte aligned
    static_assert(IN_B % 4 == 0, "H item input must be word aligned");
    static_assert(OUT_B % 16 == 0, "H item output must be 16-byte aligned");
    static_assert(RB % TAPS == 0, "chunk must be a multiple of the register-window rotation");
    static_assert(BOX_B / 4 <= 256, "TMA box too wide");
    static_assert(N <= 32, "phase table too large for kernel params");
};

template <class G>
struct __align__(128) FastSmem {
    uint8_t in[2][G::STAGE_B];            // TMA destinations (double buffered), row lr at lr * BOX_B
    uint8_t ring[G::RING][F_SW_MAX];      // H-pass results (uint8), row r lives in slot (r - rs) % RING
    float wv[32][8];                      // V-pass copy of the polyphase table
    uint32_t hlist[F_LIST], vlist[F_LIST];
    unsigned long long bar[2];
    int hcount, vcount;
};

// exact H sample: the reference's loop full_TB.h:58-63 for output byte `obyte` of input row gy
template <int C, int A, int N, int D>
__device__ __noinline__ uint8_t exact_h(const FastParams &p, const uint8_t *in_frame, int gy, int obyte) {
    if (gy < 0 || gy >= p.in_h) return 0;
    const int xx = obyte / C, c = obyte - xx * C;
    const int first = (int)(((long long)xx * D) / N) - A + 1;
    const uint8_t *row = in_frame + (long long)(gy - p.in_row0) * p.in_pitch;
    const double *w = p.wdx + (long long)xx * (2 * A);
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 2 * A; k++) {
        const int px = first + k;
        const uint8_t v = (px >= 0 && px < p.in_w) ? row[(long long)px * C + c] : (uint8_t)0;
        sum = __dadd_rn(sum, __dmul_rn((double)v, w[k]));
    }
    return quantise_f64(sum);
}

// exact V sample (full_TB.h:71-75 arithmetic) from the uint8 intermediate rows in the shared ring
template <int TAPS, int RING>
__device__ __noinline__ uint8_t exact_v(const uint8_t (*ring)[F_SW_MAX], const double *wd, int first_slot_row, int b) {
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < TAPS; k++) {
        const uint8_t v = ring[(first_slot_row + k) % RING][b];
        sum = __dadd_rn(sum, __dmul_rn((double)v, wd[k]));
    }
    return quantise_f64(sum);
}

template <int C, int A, int N, int D, int PH, int KM>
__global__ void __launch_bounds__(F_THREADS, 2)
lanczos_fast_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ FastParams p) {
    using G = Geo<C, A, N, D, PH>;
    constexpr int TAPS = G::TAPS;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FastSmem<G> &sm = *reinterpret_cast<FastSmem<G> *>(smem_raw);

    const int tid = threadIdx.x;
    const int strip = blockIdx.x, seg = blockIdx.y, frame = blockIdx.z;
    const uint8_t *in_frame = p.in + (long long)frame * p.in_frame_stride;
    uint8_t *out_frame = p.out + (long long)frame * p.out_frame_stride;

    // horizontal extent
    const int obyte0 = strip * p.sw;                          // first output byte column of the strip
    const int row_bytes = p.out_w * C;
    const int valid_bytes = min(p.sw, row_bytes - obyte0);    // > 0 by construction of the grid
    const int groups = min(p.groups, (valid_bytes + G::OUT_B - 1) / G::OUT_B);
    const int ibyte0 = (obyte0 / (N * C)) * (D * C);          // first input byte column of the strip
    // vertical extent: periods [pv0, pv1) -> output rows [N*pv0, N*pv1), clipped to the band
    const int pv0 = p.vperiod0 + seg * p.seg_periods;
    const int y_end_band = p.out_row0 + p.out_rows;
    const int pv1 = min(pv0 + p.seg_periods, (y_end_band + N - 1) / N);
    const int ys = max(N * pv0, p.out_row0), ye = min(N * pv1, y_end_band);
    if (ys >= ye) return;
    const int rs = D * pv0 - A + 1;                           // first intermediate row pushed
    const int nrows = D * (pv1 - pv0) + TAPS - 1;             // rows to push
    const int nchunks = (nrows + G::RB - 1) / G::RB;

    const uint32_t bar0 = smem_u32(&sm.bar[0]), bar1 = smem_u32(&sm.bar[1]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.hcount = 0;
        sm.vcount = 0;
    }
    for (int i = tid; i < 32 * 8; i += F_THREADS) (&sm.wv[0][0])[i] = p.wtab[i];
    __syncthreads();

    constexpr uint32_t kStageBytes = G::RB * G::BOX_B;
    auto issue = [&](int chunk) {
        const int st = chunk & 1;
        const uint32_t bar = st ? bar1 : bar0;
        mbar_expect_tx(bar, kStageBytes);
        tma_load_3d(smem_u32(&sm.in[st][0]), &in_map, (ibyte0 - G::PAD_L) / 4, rs + chunk * G::RB - p.in_row0, frame, bar);
    };
    if (tid == 0) {
        issue(0);
        if (nchunks > 1) issue(1);
    }

    // V-pass state: the last TAPS intermediate rows of this thread's word column, as fp32
    float win[TAPS][4];
#pragma unroll
    for (int j = 0; j < TAPS; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) win[j][e] = 0.f;
    const bool v_active = 4 * tid < valid_bytes;
    const float g2 = 2.f * p.guard;

    for (int chunk = 0; chunk < nchunks; chunk++) {
        const int st = chunk & 1;
        mbar_wait(st ? bar1 : bar0, (chunk >> 1) & 1);
        const int r0 = rs + chunk * G::RB;                    // first row of this chunk

        // ------------------------------ H pass ------------------------------
        for (int item = tid; item < G::RB * groups; item += F_THREADS) {
            const int lr = item / groups, g = item - lr * groups;
            const int gy = r0 + lr;
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&sm.in[st][lr * G::BOX_B + G::WIN0 + g * G::IN_B]);
            float f[G::NWORDS * 4];
#pragma unroll
            for (int wi = 0; wi < G::NWORDS; wi++)
                word_to_f32x4(src[wi], f[4 * wi], f[4 * wi + 1], f[4 * wi + 2], f[4 * wi + 3]);
            // window byte k (k = 0 is HALO_L bytes left of the item's own input) = f[MIS + k]
            uint32_t outw[G::OUT_B / 4];
            uint32_t need_fix = 0;        // bit per output word
#pragma unroll
            for (int ow = 0; ow < G::OUT_B / 4; ow++) {
                float xa[4], xb[4];
                bool flag = false;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int o = 4 * ow + e;              // output byte of the item
                    const int px = o / C, c = o % C;       // pixel within the item, channel
                    const int per = px / N, r = px % N;    // period, position in period
                    const int ph = (r * D) % N;            // phase
                    const int base = (per * D + (r * D) / N) * C + c;   // window index of tap 0
                    if (ph == 0) {
                        const float v = f[G::MIS + base + (A - 1) * C];
                        xa[e] = v;
                        xb[e] = v;
                        if (p.exact_x) {
                            float z = v;
#pragma unroll
                            for (int k = 0; k < TAPS; k++)
                                if ((KM >> k) & 1) z = fmaf(f[G::MIS + base + k * C], -p.align_k[k], z);
                            flag |= (z < 0.02f) && (v > 0.5f);
                        } else {
                            flag = true;                    // inexact alignment: always recompute
                        }
                    } else {
                        float acc = -p.guard;
#pragma unroll
                        for (int k = 0; k < TAPS; k++) acc = fmaf(f[G::MIS + base + k * C], p.wtab[ph * 8 + k], acc);
                        xa[e] = acc;
                        xb[e] = acc + g2;
                    }
                }
                const uint32_t qa = quantise4(xa[0], xa[1], xa[2], xa[3]);
                const uint32_t qb = quantise4(xb[0], xb[1], xb[2], xb[3]);
                outw[ow] = qa;
                if ((flag || qa != qb) && g * G::OUT_B + 4 * ow < valid_bytes) need_fix |= 1u << ow;
            }
            const int slot = (gy - rs) % G::RING;
            uint4 *dst = reinterpret_cast<uint4 *>(&sm.ring[slot][g * G::OUT_B]);
#pragma unroll
            for (int v4 = 0; v4 < G::OUT_B / 16; v4++)
                dst[v4] = make_uint4(outw[4 * v4], outw[4 * v4 + 1], outw[4 * v4 + 2], outw[4 * v4 + 3]);
            while (need_fix) {                               // rare: queue the word for exact recomputation
                const int ow = __ffs(need_fix) - 1;
                need_fix &= need_fix - 1;
                const int pos = atomicAdd(&sm.hcount, 1);
                const uint32_t entry = ((uint32_t)lr << 16) | (uint32_t)(g * G::OUT_B + 4 * ow);
                if (pos < F_LIST) {
                    sm.hlist[pos] = entry;
                } else {                                     // list full: recompute in place
                    for (int e = 0; e < 4; e++)
                        sm.ring[slot][g * G::OUT_B + 4 * ow + e] =
                            exact_h<C, A, N, D>(p, in_frame, gy, obyte0 + g * G::OUT_B + 4 * ow + e);
                }
            }
        }
        __syncthreads();
        {   // deferred exact H samples, one lane per byte
            const int n = min(sm.hcount, F_LIST);
            if (n > 0) {
                for (int i = tid; i < 4 * n; i += F_THREADS) {
                    const uint32_t entry = sm.hlist[i >> 2];
                    const int lr = entry >> 16, b = (entry & 0xffff) + (i & 3);
                    const int gy = r0 + lr;
                    sm.ring[(gy - rs) % G::RING][b] = exact_h<C, A, N, D>(p, in_frame, gy, obyte0 + b);
                }
                if (p.strict_counter && tid == 0) atomicAdd(p.strict_counter, (unsigned long long)(4 * sm.hcount));
                __syncthreads();
                if (tid == 0) sm.hcount = 0;
            }
        }
        // the input stage is free again: prefetch chunk+2 into it
        if (tid == 0 && chunk + 2 < nchunks) issue(chunk + 2);

        // ------------------------------ V pass ------------------------------
        if (v_active) {
            uint8_t *ocol = out_frame + obyte0 + 4 * tid;
#pragma unroll 1
            for (int lb = 0; lb < G::RB; lb += TAPS) {
#pragma unroll
            for (int j = 0; j < TAPS; j++) {                   // j = static window slot of the new row
                const int r = r0 + lb + j;
                const uint32_t w = *reinterpret_cast<const uint32_t *>(&sm.ring[(r - rs) % G::RING][4 * tid]);
                word_to_f32x4(w, win[j][0], win[j][1], win[j][2], win[j][3]);
                const int m = r - A;                           // floor(y*D/N) of the rows that complete now
                if (m < D * pv0) continue;
                const int y_lo = (m * N + D - 1) / D, y_hi = min(((m + 1) * N + D - 1) / D, ye);
#pragma unroll 1
                for (int y = max(y_lo, ys); y < y_hi; y++) {
                    const int ph = (y * D) % N;
                    uint32_t q;
                    bool fix = false;
                    if (ph == 0 && p.exact_y) {
                        // phase 0: the centre tap (window row m) is the result
                        q = *reinterpret_cast<const uint32_t *>(&sm.ring[(m - rs) % G::RING][4 * tid]);
                        if (p.strict_v_identity) {
                            float z[4];
#pragma unroll
                            for (int e = 0; e < 4; e++) z[e] = win[(j + 1 + (A - 1)) % TAPS][e];
                            bool dark = false;
#pragma unroll
                            for (int k = 0; k < TAPS; k++) {
                                if (!((KM >> k) & 1)) continue;
                                const float kk = -p.align_k[k];
#pragma unroll
                                for (int e = 0; e < 4; e++) z[e] = fmaf(win[(j + 1 + k) % TAPS][e], kk, z[e]);
                            }
#pragma unroll
                            for (int e = 0; e < 4; e++) dark |= (z[e] < 0.02f) && (win[(j + 1 + (A - 1)) % TAPS][e] > 0.5f);
                            fix = dark;
                        }
                    } else {
                        const float4 wlo = *reinterpret_cast<const float4 *>(&sm.wv[ph][0]);
                        const float4 whi = *reinterpret_cast<const float4 *>(&sm.wv[ph][4]);
                        const float wk[8] = {wlo.x, wlo.y, wlo.z, wlo.w, whi.x, whi.y, whi.z, whi.w};
                        float xa[4], xb[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) xa[e] = -p.guard;
#pragma unroll
                        for (int k = 0; k < TAPS; k++)
#pragma unroll
                            for (int e = 0; e < 4; e++) xa[e] = fmaf(win[(j + 1 + k) % TAPS][e], wk[k], xa[e]);
#pragma unroll
                        for (int e = 0; e < 4; e++) xb[e] = xa[e] + g2;
                        q = quantise4(xa[0], xa[1], xa[2], xa[3]);
                        fix = q != quantise4(xb[0], xb[1], xb[2], xb[3]);
                    }
                    *reinterpret_cast<uint32_t *>(ocol + (long long)(y - p.out_row0) * p.out_pitch) = q;
                    if (fix) {
                        const int pos = atomicAdd(&sm.vcount, 1);
                        if (pos < F_LIST) {
                            sm.vlist[pos] = ((uint32_t)(y - ys) << 16) | (uint32_t)(4 * tid);
                        } else {                               // list full: recompute in place
                            for (int e = 0; e < 4; e++)
                                ocol[(long long)(y - p.out_row0) * p.out_pitch + e] = exact_v<TAPS, G::RING>(
                                    sm.ring, p.wdy + (long long)y * TAPS, m - A + 1 - rs, 4 * tid + e);
                        }
                    }
                }
            }
            }
        }
        __syncthreads();
        {   // deferred exact V samples (full_TB.h:71-75 arithmetic on the uint8 intermediate rows)
            const int n = min(sm.vcount, F_LIST);
            if (n > 0) {
                for (int i = tid; i < 4 * n; i += F_THREADS) {
                    const uint32_t entry = sm.vlist[i >> 2];
                    const int y = ys + (int)(entry >> 16), b = (int)(entry & 0xffff) + (i & 3);
                    const int first = (int)(((long long)y * D) / N) - A + 1;
                    out_frame[(long long)(y - p.out_row0) * p.out_pitch + obyte0 + b] =
                        exact_v<TAPS, G::RING>(sm.ring, p.wdy + (long long)y * TAPS, first - rs, b);
                }
                if (p.strict_counter && tid == 0) atomicAdd(p.strict_counter, (unsigned long long)(4 * sm.vcount));
                __syncthreads();
                if (tid == 0) sm.vcount = 0;
                __syncthreads();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
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

template <int C, int A, int N, int D, int PH, int KM>
int launch_one(const KParams &k, const float *phase_w, const float *align_k, int exact_x, int exact_y, cudaStream_t s) {
    using G = Geo<C, A, N, D, PH>;
    EncodeFn encode = get_encode();
    if (!encode) return -1;
    const int row_bytes = k.out_w * C;
    // strip width: a multiple of OUT_B (<= 1024 bytes, one word per V thread) whose input start stays
    // 16-byte aligned for every strip (TMA needs a 16-byte aligned box start), wasting the fewest threads
    int best_groups = 0;
    double best_eff = -1;
    for (int gr = G::MAX_GROUPS; gr >= 1; gr--) {
        if ((gr * G::IN_B) % 16 != 0) continue;
        const int sw_c = gr * G::OUT_B;
        const int strips_c = (row_bytes + sw_c - 1) / sw_c;
        const double eff = (double)row_bytes / ((double)strips_c * F_SW_MAX);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_groups = gr; }
    }
    if (best_groups == 0) return -1;
    const int sw = best_groups * G::OUT_B;
    const int strips = (row_bytes + sw - 1) / sw;
    const int vperiod0 = k.out_row0 / N;
    const int vperiods = (k.out_row0 + k.out_rows + N - 1) / N - vperiod0;
    // vertical segments: enough CTAs for >= ~8 waves of 2 CTAs/SM, but at least 4 chunks per segment
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int min_periods = std::max(1, (4 * G::RB) / D);
    int segs = (int)((8LL * 2 * sms + (long long)strips * k.n_frames - 1) / ((long long)strips * k.n_frames));
    segs = std::max(1, std::min(segs, std::max(1, vperiods / min_periods)));
    int seg_periods = (vperiods + segs - 1) / segs;
    segs = (vperiods + seg_periods - 1) / seg_periods;

    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)(k.in_w * C / 4), (cuuint64_t)k.in_rows, (cuuint64_t)k.n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)k.in_pitch, (cuuint64_t)(k.n_frames > 1 ? k.in_frame_stride : k.in_pitch * k.in_rows)};
    const cuuint32_t box[3] = {(cuuint32_t)(G::BOX_B / 4), (cuuint32_t)G::RB, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t *>(k.in), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -1;

    FastParams p{};
    p.in = k.in; p.out = k.out;
    p.in_pitch = k.in_pitch; p.out_pitch = k.out_pitch;
    p.in_frame_stride = k.in_frame_stride; p.out_frame_stride = k.out_frame_stride;
    p.in_w = k.in_w; p.in_h = k.in_h; p.out_w = k.out_w; p.out_h = k.out_h;
    p.out_row0 = k.out_row0; p.out_rows = k.out_rows; p.in_row0 = k.in_row0; p.in_rows = k.in_rows;
    p.sw = sw; p.groups = best_groups; p.seg_periods = seg_periods; p.vperiod0 = vperiod0;
    p.wdx = k.wdx; p.wdy = k.wdy; p.guard = k.guard;
    p.exact_x = exact_x; p.exact_y = exact_y;
    p.strict_v_identity = (k.flags & LANCZOS_FLAG_FAST_ALIGNED) ? 0 : 1;
    for (int i = 0; i < 8; i++) p.align_k[i] = i < 2 * A ? align_k[i] : 0.f;
    for (int ph = 0; ph < N; ph++)
        for (int t = 0; t < 8; t++) p.wtab[ph * 8 + t] = t < 2 * A ? phase_w[ph * 2 * A + t] : 0.f;
    p.strict_counter = k.strict_counter;

    auto kern = lanczos_fast_kernel<C, A, N, D, PH, KM>;
    const size_t smem = sizeof(FastSmem<G>) + 128;
    static bool attr_set[64] = {};
    if (!attr_set[dev & 63]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        attr_set[dev & 63] = true;
    }
    dim3 grid(strips, segs, k.n_frames);
    kern<<<grid, F_THREADS, smem, s>>>(map, p);
    return (int)cudaGetLastError();
}

}  // namespace

// Returns 0 on launch, >0 cudaError, -1 when no specialised kernel applies (caller falls back).
int launch_fast(const KParams &k, const float *phase_w_host, const float *align_k_host, int exact_x, int exact_y,
                int *kernel_id, cudaStream_t s) {
    // layout requirements of the TMA map and of the 32-bit/128-bit accesses
    if ((k.in_w * k.channels) % 4 != 0 || (k.out_w * k.channels) % 4 != 0) return -1;
    if (k.in_pitch % 16 != 0 || k.out_pitch % 4 != 0) return -1;
    if ((reinterpret_cast<uintptr_t>(k.in) & 15) != 0 || (reinterpret_cast<uintptr_t>(k.out) & 3) != 0) return -1;
    if (k.n_frames > 1 && (k.in_frame_stride % 16 != 0 || k.out_frame_stride % 4 != 0)) return -1;
    if (k.out_rows >= 65536) return -1;   // list entries hold the row in 16 bits per segment; keep it simple
    const int C = k.channels, A = k.a, N = k.scale_n, D = k.scale_d;
    // KM: taps whose phase-0 residue is negative (nonzero filter constant); the host table must agree
    int km = 0;
    for (int t = 0; t < 2 * A; t++)
        if (align_k_host[t] != 0.f) km |= 1 << t;
#define LZ_CASE(c, a, n, d, ph, kmask, id)                                                             \
    if (C == c && A == a && N == n && D == d && (km & ~(kmask)) == 0) {                                 \
        *kernel_id = id;                                                                                \
        return launch_one<c, a, n, d, ph, kmask>(k, phase_w_host, align_k_host, exact_x, exact_y, s);   \
    }
    // a = 3: sin(2*pi) < 0 in double, so the |d| = 2 taps (k = 0 and k = 4) carry negative residues
    LZ_CASE(3, 3, 2, 1, 8, 0x11, 1)
    LZ_CASE(4, 3, 2, 1, 4, 0x11, 2)
    LZ_CASE(4, 3, 3, 2, 4, 0x11, 3)
    LZ_CASE(3, 2, 2, 1, 8, 0x8, 4)
#undef LZ_CASE
    return -1;
}

}  // namespace lzb
