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
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <cuda.h>

#include "../../include/lanczos_b200.h"
#include "fast_common.cuh"

#ifndef LZB_ALWAYS_CHECK
#define LZB_ALWAYS_CHECK 0
#endif

namespace lzb {

namespace {


// ---------------------------------------------------------------------------------------------
// parameters
// ---------------------------------------------------------------------------------------------
struct FastParams {
    const uint8_t *in;    // input row `in_row0` of frame 0 (not dereferenced by the kernel; TMA reads it)
    uint8_t *out;         // output row `out_row0` of frame 0
    long long in_pitch, out_pitch, in_frame_stride, out_frame_stride;
    int in_w, in_h, out_w, out_h;
    int out_row0, out_rows, in_row0, in_rows;
    int sw;               // strip width in output bytes (groups * OUT_B)
    int groups;           // H-pass thread groups per strip row
    int seg_periods;      // vertical ratio-periods per segment
    int vperiod0;         // first vertical period covered by the launch (floor(out_row0 / N))
    const double *wdx, *wdy;  // per-coordinate double weights (exact recomputation, non-uniform phases)
    float guard;
    int exact_x, exact_y;     // every phase-0 coordinate exactly integral (plan AxisTables.aligned_exact)
    int uniform_x, uniform_y; // double weights identical for all coordinates of a phase -> wdtab usable
    int strict_v_identity;    // 0 with LANCZOS_FLAG_FAST_ALIGNED
    float align_k[8];         // phase-0 "cannot flip" constants
    int align_ki[8];          // the same, ceil(K * 2^16), for the integer re-check in the slow paths
    float wtab[32 * 8];       // polyphase table [N][8] (padded to 8 taps), N <= 32
    double wdtab[8 * 8];      // double polyphase table [N][8] for N <= 8 (valid when uniform_*)
    unsigned long long *strict_counter;
};

template <int C, int A, int N, int D, int PH, int NT>
struct Geo {
    static constexpr int THREADS = NT;
    static constexpr int SW_MAX = 4 * NT;       // strip width in output bytes (one 32-bit word per V thread)
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
    static constexpr int MAX_GROUPS = SW_MAX / OUT_B;
    static constexpr int BOX_B = 16 * ((PAD_L + MAX_GROUPS * IN_B + HALO_R + 15) / 16);  // TMA box row bytes
    static constexpr int RB = 12;                                 // input rows per chunk
    static constexpr int RING = 3 * RB;                           // intermediate rows kept in smem
    static constexpr int STAGE_B = 128 * ((RB * BOX_B + 127) / 128);  // TMA destinations must be 128-byte aligned
    // the first row pushed by a segment is rs = D*pv0 - A + 1, so (row - A) mod D is static per chunk row
    static constexpr int S0 = (((1 - 2 * A) % D) + D) % D;
    static constexpr int UNR = (TAPS % D == 0) ? TAPS : TAPS * D; // rows per statically unrolled V block (multiple of TAPS and D)
    static constexpr int YSPAN = (UNR + S0) * N / D + 2;          // bound on output rows touched per V block
    static_assert(IN_B % 4 == 0, "H item input must be word aligned");
    static_assert(OUT_B % 16 == 0, "H item output must be 16-byte aligned");
    static_assert(RB % UNR == 0, "chunk must be a multiple of the unrolled V block");
    static_assert(RING >= 2 * RB + TAPS && RING % RB == 0, "ring too small for barrier-free V/H overlap");
    static_assert(BOX_B / 4 <= 256, "TMA box too wide");
    static_assert(N <= 32, "phase table too large for kernel params");
    static constexpr int VROWS = ((S0 + UNR) * N + D - 1) / D - (S0 * N + D - 1) / D;  // output rows per V block
    static_assert(YSPAN <= 32, "fix mask (bit per output row) too small");
    static_assert(VROWS <= 16 && OUT_B / 4 <= 16, "fix queue (16 entries per lane) too small");
};

template <class G>
struct __align__(128) FastSmem {
    uint8_t in[2][G::STAGE_B];            // TMA destinations (double buffered), row lr at lr * BOX_B
    uint8_t ring[G::RING][G::SW_MAX];     // H-pass results (uint8), row r lives in slot (r - rs) % RING
    unsigned long long bar[2];
    uint16_t fixq[G::THREADS / 32][32 * 16];  // per-warp queue of words to look at again (lane << 5 | index)
    uint16_t fixb[G::THREADS / 32][32 * 4];   // per-warp queue of bytes to recompute exactly (batch lane << 5 | byte)
};

__host__ __device__ constexpr int cdiv_c(int a, int b) { return (a + b - 1) / b; }

template <int C, int A, int N, int D, int PH, int KM, int NT>
__global__ void __launch_bounds__(NT, (NT == 128 ? 7 : 3))
lanczos_fast_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ FastParams p) {
    using G = Geo<C, A, N, D, PH, NT>;
    constexpr int TAPS = G::TAPS;
    constexpr int SWM = G::SW_MAX;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FastSmem<G> &sm = *reinterpret_cast<FastSmem<G> *>(smem_raw);

    const int tid = threadIdx.x;
    const int strip = blockIdx.x, seg = blockIdx.y, frame = blockIdx.z;
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
    }
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

    // V-pass state: the last TAPS intermediate rows of this thread's word column, as fp32 pairs
    // (value * 2^-24), plus the raw words (phase-0 rows are copies of the centre tap)
    float2 win[TAPS][2];
    uint32_t raw[TAPS];
#pragma unroll
    for (int j = 0; j < TAPS; j++) {
        win[j][0] = win[j][1] = make_float2(0.f, 0.f);
        raw[j] = 0;
    }
    const bool v_active = 4 * tid < valid_bytes;
    const float guard = p.guard, g2 = 2.f * p.guard;
    const long long opitch = p.out_pitch;
    const uint8_t *vcol = &sm.ring[0][4 * tid];               // this thread's word column of the ring
    uint8_t *ocol = out_frame + obyte0 + 4 * tid;
    unsigned long long n_strict = 0;
    // flag handling without loop-invariant branches (they would multiply the unrolled code):
    //   *_force    sign bit set -> every phase-0 sample is recomputed (inexact alignment, e.g. 17/10)
    //   v_signmask 0            -> phase-0 rows are never recomputed (LANCZOS_FLAG_FAST_ALIGNED)
    const uint32_t v_force = p.exact_y ? 0u : 0x80000000u;
    const uint32_t v_signmask = (p.exact_y && !p.strict_v_identity) ? 0u : 0x80000000u;
    const uint32_t h_force = p.exact_x ? 0u : 0x80000000u;

    // H-pass work distribution is the same for every chunk: item = round * NT + tid -> (row, group)
    constexpr int HROUNDS = (G::RB * G::MAX_GROUPS + NT - 1) / NT;
    int h_lr[HROUNDS], h_g[HROUNDS];
#pragma unroll
    for (int r = 0; r < HROUNDS; r++) {
        const int item = r * NT + tid;
        h_lr[r] = item / groups;
        h_g[r] = item - h_lr[r] * groups;
    }
    const int lane = tid & 31;
    uint16_t *wq = sm.fixq[tid >> 5], *wb = sm.fixb[tid >> 5];
    // running output pointer of the V pass: row (ybase - out_row0) of this thread's column, advanced per block
    const int t0_first = (rs - A - G::S0) / D;                 // exact division (also for negative values)
    // uniform part of the output address (row ybase of the strip) + a 32-bit per-thread offset
    uint8_t *obase_u = out_frame + obyte0 + (long long)(N * t0_first - p.out_row0) * opitch;
    const uint32_t ocoff = 4u * (uint32_t)tid, opitch32 = (uint32_t)opitch;
    const long long oblock = (long long)(N * (G::UNR / D)) * opitch;   // output rows per V block
    int ybase = N * t0_first;

    for (int chunk = 0; chunk < nchunks; chunk++) {
        const int st = chunk & 1;
        mbar_wait(st ? bar1 : bar0, (chunk >> 1) & 1);
        const int slot0 = (chunk % (G::RING / G::RB)) * G::RB;   // ring slot of this chunk's first row (no wrap inside)

        // ------------------------------ H pass ------------------------------
#pragma unroll
        for (int hr = 0; hr < HROUNDS; hr++) {
            const int item = hr * NT + tid;
            const unsigned hmask = __ballot_sync(0xffffffffu, item < G::RB * groups);
            if (item >= G::RB * groups) continue;
            const int lr = h_lr[hr], g = h_g[hr];
            const uint8_t *srow = &sm.in[st][lr * G::BOX_B];
            const uint32_t *src = reinterpret_cast<const uint32_t *>(srow + G::WIN0 + g * G::IN_B);
            float f[G::NWORDS * 4];
#pragma unroll
            for (int wi = 0; wi < G::NWORDS; wi++)
                word_to_f32x4(src[wi], f[4 * wi], f[4 * wi + 1], f[4 * wi + 2], f[4 * wi + 3]);
            // window byte k (k = 0 is HALO_L bytes left of the item's own input) = f[MIS + k] * 2^24
            uint32_t outw[G::OUT_B / 4];
            uint32_t fix_g = 0, fix_z = 0;      // bit per output word: truncation in doubt / phase-0 sample may flip
#pragma unroll
            for (int ow = 0; ow < G::OUT_B / 4; ow++) {
                float xa[4], xb[4];
                uint32_t zor = 0;          // sign bit set <=> some phase-0 sample may flip (or must be recomputed)
                uint32_t p0mask = 0;       // which bytes of the word are phase-0 samples (static)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int o = 4 * ow + e;              // output byte of the item
                    const int px = o / C, c = o % C;       // pixel within the item, channel
                    const int per = px / N, r = px % N;    // period, position in period
                    const int ph = (r * D) % N;            // phase
                    const int base = (per * D + (r * D) / N) * C + c;   // window index of tap 0
                    if (ph == 0) {
                        const float v = f[G::MIS + base + (A - 1) * C];
                        xa[e] = xb[e] = v * kPixUnscale;
                        // "cannot flip" filter of plan.cpp, one test per side of the centre tap:
                        // v - sum K_k*b_k >= 0 over the negative residues before / after it -> output is v
                        float zpre = v, zpost = v;
#pragma unroll
                        for (int k = 0; k < TAPS; k++)
                            if ((KM >> k) & 1) {
                                if (k < A - 1) zpre = fmaf(f[G::MIS + base + k * C], -p.align_k[k], zpre);
                                else zpost = fmaf(f[G::MIS + base + k * C], -p.align_k[k], zpost);
                            }
                        zor |= __float_as_uint(zpre) | __float_as_uint(zpost);
                        p0mask |= 1u << e;
                    } else {
                        float acc = -guard;
#pragma unroll
                        for (int k = 0; k < TAPS; k++) acc = fmaf(f[G::MIS + base + k * C], p.wtab[ph * 8 + k], acc);
                        xa[e] = acc;
                        xb[e] = acc + g2;
                    }
                }
                const uint32_t qa = quantise4(xa[0], xa[1], xa[2], xa[3]);
                const uint32_t qb = quantise4(xb[0], xb[1], xb[2], xb[3]);
                outw[ow] = qa;
                if (p0mask) zor |= h_force;                  // inexact alignment: phase-0 samples always recomputed
                // branch-free flag: the word is recomputed exactly if a truncation is in doubt (qa != qb)
                // or a phase-0 sample may flip (sign of zor)
                fix_g |= (qa != qb ? 1u : 0u) << ow;
                fix_z |= (zor >> 31) << ow;
            }
            uint8_t *drow = &sm.ring[slot0 + lr][g * G::OUT_B];
            uint4 *dst = reinterpret_cast<uint4 *>(drow);
#pragma unroll
            for (int v4 = 0; v4 < G::OUT_B / 16; v4++)
                dst[v4] = make_uint4(outw[4 * v4], outw[4 * v4 + 1], outw[4 * v4 + 2], outw[4 * v4 + 3]);
            // rare: words whose truncation is in doubt or whose phase-0 samples may flip.  The warp pools
            // them (uniform noise flags ~1/4 of the words): stage A re-checks every byte of a flagged word
            // with the integer phase-0 filter, stage B pools the surviving bytes, stage C recomputes them
            // exactly, one byte per lane.
            if (__ballot_sync(hmask, (fix_g | fix_z) != 0)) {
                const int nfix = warp_enqueue(fix_g | fix_z, fix_g, hmask, lane, wq);
                const int nact = __popc(hmask);
                auto decode = [&](uint32_t ent, int e, const uint8_t *&tap0, uint8_t *&dst, int &ph, int &xx) -> bool {
                    const int item_l = item - lane + (int)((ent >> 5) & 31u), ow = (int)(ent & 31u);
                    const int lr_l = item_l / groups, g_l = item_l - lr_l * groups;
                    const int b = 4 * ow + e;
                    if (g_l * G::OUT_B + b >= valid_bytes) return false;
                    const int ob = obyte0 + g_l * G::OUT_B + b;              // global output byte column
                    xx = ob / C;
                    const int c = ob - xx * C;
                    const int first = (xx * D) / N - A + 1;                    // first tap pixel (full_TB.h:59)
                    ph = (xx * D) % N;
                    tap0 = &sm.in[st][lr_l * G::BOX_B] + G::PAD_L + first * C + c - ibyte0;
                    dst = &sm.ring[slot0 + lr_l][g_l * G::OUT_B + b];
                    return true;
                };
                for (int base = 0; base < nfix; base += nact) {
                    uint32_t m4 = 0;
                    if (base + lane < nfix) {
                        const uint32_t ent = wq[base + lane];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const uint8_t *tap0; uint8_t *dst; int ph, xx;
                            if (!decode(ent, e, tap0, dst, ph, xx)) continue;
                            const bool need = (ph == 0) ? !(p.exact_x && phase0_safe<TAPS, KM>(tap0, C, p.align_ki))
                                                        : (ent >> 15) != 0;
                            m4 |= (need ? 1u : 0u) << e;
                        }
                    }
                    const int nb = warp_enqueue(m4, 0u, hmask, lane, wb);
                    for (int j = lane; j < nb; j += nact) {
                        const uint32_t eb = wb[j];
                        const uint8_t *tap0; uint8_t *dst; int ph, xx;
                        decode(wq[base + (int)(eb >> 5)], (int)(eb & 31u), tap0, dst, ph, xx);
                        if (p.uniform_x && N <= 8) *dst = exact_taps<TAPS>(tap0, C, [&](int k) { return p.wdtab[ph * 8 + k]; });
                        else *dst = exact_taps<TAPS>(tap0, C, [&](int k) { return p.wdx[(long long)xx * TAPS + k]; });
                        n_strict++;
                    }
                    __syncwarp(hmask);
                }
            }
        }
        __syncthreads();
        // the input stage is free again: prefetch chunk+2 into it
        if (tid == 0 && chunk + 2 < nchunks) issue(chunk + 2);

        // ------------------------------ V pass ------------------------------
        const unsigned vmask = __ballot_sync(0xffffffffu, v_active);
        if (v_active) {
#pragma unroll 1
            for (int sub = 0; sub < G::RB / G::UNR; sub++, ybase += N * (G::UNR / D), obase_u += oblock) {
                uint8_t *obase = obase_u + ocoff;
                // rows rb..rb+UNR-1 arrive (rb = rs + chunk*RB + sub*UNR); row r completes the outputs y with
                // floor(y*D/N) = r - A.  rb - A = D*t0 + S0 exactly, so everything relative to ybase = N*t0 is
                // static; ybase and the output pointer obase advance by one block per iteration.
                const uint8_t *vrow = vcol + (slot0 + sub * G::UNR) * SWM;   // no ring wrap inside a block
                uint32_t fixrows = 0;                              // bit per output row: recompute this thread's word exactly
                const bool interior = (ybase >= ys) && (ybase + G::YSPAN <= ye);
                auto body = [&](auto check_tag) {
                    constexpr bool CHECK = decltype(check_tag)::value;
#pragma unroll
                    for (int lr = 0; lr < G::UNR; lr++) {
                        const int j = lr % TAPS;                    // static window slot of the new row
                        const uint32_t w = *reinterpret_cast<const uint32_t *>(vrow + lr * SWM);
                        raw[j] = w;
                        word_to_f32x4(w, win[j][0].x, win[j][0].y, win[j][1].x, win[j][1].y);
                        const int s = (G::S0 + lr) % D, tq = (G::S0 + lr) / D;     // r - A = D*(t0+tq) + s
                        const int yfirst = N * tq + cdiv_c(s * N, D), ylast = N * tq + cdiv_c((s + 1) * N, D);
#pragma unroll
                        for (int yy = yfirst; yy < ylast; yy++) {
                            const int ph = (yy * D) % N;
                            if (CHECK && (ybase + yy < ys || ybase + yy >= ye)) continue;
                            uint32_t q;
                            if (ph == 0) {
                                // phase 0: the centre tap (row r - A) is the result; flag it unless the
                                // "cannot flip" filter of plan.cpp proves the reference returns it too
                                q = raw[(j + 1 + (A - 1)) % TAPS];
                                float2 z0 = win[(j + 1 + (A - 1)) % TAPS][0], z1 = win[(j + 1 + (A - 1)) % TAPS][1];
#pragma unroll
                                for (int k = 0; k < TAPS; k++)
                                    if ((KM >> k) & 1) {
                                        const float2 kk = make_float2(-p.align_k[k], -p.align_k[k]);
                                        z0 = __ffma2_rn(win[(j + 1 + k) % TAPS][0], kk, z0);
                                        z1 = __ffma2_rn(win[(j + 1 + k) % TAPS][1], kk, z1);
                                    }
                                const uint32_t zor = v_force | __float_as_uint(z0.x) | __float_as_uint(z0.y) |
                                                     __float_as_uint(z1.x) | __float_as_uint(z1.y);
                                fixrows |= ((zor & v_signmask) != 0 ? 1u : 0u) << yy;
                            } else {
                                float2 a0 = make_float2(-guard, -guard), a1 = a0;
#pragma unroll
                                for (int k = 0; k < TAPS; k++) {
                                    const float2 wk = make_float2(p.wtab[ph * 8 + k], p.wtab[ph * 8 + k]);
                                    a0 = __ffma2_rn(win[(j + 1 + k) % TAPS][0], wk, a0);
                                    a1 = __ffma2_rn(win[(j + 1 + k) % TAPS][1], wk, a1);
                                }
                                const float2 gg = make_float2(g2, g2);
                                const float2 b0 = __fadd2_rn(a0, gg), b1 = __fadd2_rn(a1, gg);
                                q = quantise4(a0.x, a0.y, a1.x, a1.y);
                                fixrows |= (q != quantise4(b0.x, b0.y, b1.x, b1.y) ? 1u : 0u) << yy;
                            }
                            *reinterpret_cast<uint32_t *>(obase_u + (ocoff + (uint32_t)yy * opitch32)) = q;
                        }
                    }
                };
                if (interior && !LZB_ALWAYS_CHECK) body(std::false_type{}); else body(std::true_type{});
                // rare: rows whose truncation is in doubt / whose phase-0 word may flip, pooled over the warp
                // (same three stages as in the H pass)
                if (__ballot_sync(vmask, fixrows != 0)) {
                    const int nfix = warp_enqueue(fixrows, 0u, vmask, lane, wq);
                    const int nact = __popc(vmask);
                    auto decode = [&](uint32_t ent, int &y, int &ph, int &s0, const uint8_t *&col, uint8_t *&orow) {
                        const int dl = (int)((ent >> 5) & 31u) - lane, yy = (int)(ent & 31u);   // owner lane offset, row
                        y = ybase + yy;
                        ph = (y * D) % N;
                        s0 = ((y * D) / N - A + 1 - rs) % G::RING;       // ring slot of the first tap row (full_TB.h:72)
                        col = vcol + 4 * dl;
                        orow = obase + (long long)yy * opitch + 4 * dl;
                    };
                    for (int base = 0; base < nfix; base += nact) {
                        uint32_t m4 = 0;
                        if (base + lane < nfix) {
                            int y, ph, s0; const uint8_t *col; uint8_t *orow;
                            decode(wq[base + lane], y, ph, s0, col, orow);
                            m4 = 15u;
                            if (ph == 0 && p.exact_y && s0 + TAPS <= G::RING) {
                                m4 = 0;
#pragma unroll
                                for (int e = 0; e < 4; e++)
                                    m4 |= (phase0_safe<TAPS, KM>(col + s0 * SWM + e, SWM, p.align_ki) ? 0u : 1u) << e;
                            }
                        }
                        const int nb = warp_enqueue(m4, 0u, vmask, lane, wb);
                        for (int j = lane; j < nb; j += nact) {
                            const uint32_t eb = wb[j];
                            const int e = (int)(eb & 31u);
                            int y, ph, s0; const uint8_t *col; uint8_t *orow;
                            decode(wq[base + (int)(eb >> 5)], y, ph, s0, col, orow);
                            // full_TB.h:71-75 on the uint8 intermediate rows; the 2a tap rows sit in consecutive
                            // ring slots unless the ring wraps inside the window
                            uint8_t taps_b[TAPS];
#pragma unroll
                            for (int k = 0; k < TAPS; k++) taps_b[k] = col[((s0 + k) % G::RING) * SWM + e];
                            if (p.uniform_y && N <= 8) orow[e] = exact_taps<TAPS>(taps_b, 1, [&](int k) { return p.wdtab[ph * 8 + k]; });
                            else orow[e] = exact_taps<TAPS>(taps_b, 1, [&](int k) { return p.wdy[(long long)y * TAPS + k]; });
                            n_strict++;
                        }
                        __syncwarp(vmask);
                    }
                }
            }
        }
        // no barrier here: RING >= 2*RB + TAPS keeps the rows this chunk's V pass reads intact while the
        // next chunk's H pass writes; the barrier after that H pass orders everything else.
    }
    if (p.strict_counter && n_strict) atomicAdd(p.strict_counter, n_strict);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int C, int A, int N, int D, int PH, int KM, int NT>
int launch_one(const KParams &k, const FastHostTables &t, cudaStream_t s) {
    using G = Geo<C, A, N, D, PH, NT>;
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
        const double eff = (double)row_bytes / ((double)strips_c * G::SW_MAX);
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
    const int ctas_per_sm = NT == 128 ? 7 : 3;
    int segs = (int)((8LL * ctas_per_sm * sms + (long long)strips * k.n_frames - 1) / ((long long)strips * k.n_frames));
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
    p.exact_x = t.exact_x; p.exact_y = t.exact_y;
    p.uniform_x = t.uniform_x; p.uniform_y = t.uniform_y;
    p.strict_v_identity = (k.flags & LANCZOS_FLAG_FAST_ALIGNED) ? 0 : 1;
    for (int i = 0; i < 8; i++) p.align_k[i] = i < 2 * A ? t.align_k[i] : 0.f;
    for (int i = 0; i < 8; i++) p.align_ki[i] = (int)std::ceil((double)p.align_k[i] * 65536.0 * 1.0001);
    for (int ph = 0; ph < N; ph++)
        for (int q = 0; q < 8; q++) p.wtab[ph * 8 + q] = q < 2 * A ? t.phase_w[ph * 2 * A + q] * 16777216.f : 0.f;  // x 2^24, see kPixUnscale
    for (int ph = 0; ph < N && ph < 8; ph++)
        for (int q = 0; q < 8; q++) p.wdtab[ph * 8 + q] = q < 2 * A ? t.phase_wd[ph * 2 * A + q] : 0.0;
    p.strict_counter = k.strict_counter;

    auto kern = lanczos_fast_kernel<C, A, N, D, PH, KM, NT>;
    const size_t smem = sizeof(FastSmem<G>) + 128;
    static std::atomic<bool> attr_set[64] = {};     // per device; the call is idempotent, a race only repeats it
    if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        attr_set[dev & 63].store(true, std::memory_order_release);
    }
    dim3 grid(strips, segs, k.n_frames);
    kern<<<grid, NT, smem, s>>>(map, p);
    return (int)cudaGetLastError();
}

}  // namespace

// Returns 0 on launch, >0 cudaError, -1 when no specialised kernel applies (caller falls back).
int launch_fast(const KParams &k, const FastHostTables &t, int *kernel_id, cudaStream_t s) {
    // layout requirements of the TMA map and of the 32-bit/128-bit accesses
    if ((k.in_w * k.channels) % 4 != 0 || (k.out_w * k.channels) % 4 != 0) return -1;
    if (k.in_pitch % 16 != 0 || k.out_pitch % 4 != 0) return -1;
    if ((reinterpret_cast<uintptr_t>(k.in) & 15) != 0 || (reinterpret_cast<uintptr_t>(k.out) & 3) != 0) return -1;
    if (k.n_frames > 1 && (k.in_frame_stride % 16 != 0 || k.out_frame_stride % 4 != 0)) return -1;
    if (k.out_pitch >= (1LL << 26)) return -1;   // V-pass stores use 32-bit offsets inside a block of rows
    const int C = k.channels, A = k.a, N = k.scale_n, D = k.scale_d;
    // tuning knob (tools only): LZB_NT=256 selects 256-thread CTAs with 1024-byte strips
    static const int nt = [] { const char *e = getenv("LZB_NT"); return e ? atoi(e) : 128; }();
    // KM: taps whose phase-0 residue is negative (nonzero filter constant); the host table must agree
    int km = 0;
    for (int q = 0; q < 2 * A; q++)
        if (t.align_k[q] != 0.f) km |= 1 << q;
#define LZ_CASE(c, a, n, d, ph, kmask, id)                                                             \
    if (C == c && A == a && N == n && D == d && (km & ~(kmask)) == 0) {                                 \
        *kernel_id = id;                                                                                \
        if (nt == 256) return launch_one<c, a, n, d, ph, kmask, 256>(k, t, s);                          \
        return launch_one<c, a, n, d, ph, kmask, 128>(k, t, s);                                         \
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
