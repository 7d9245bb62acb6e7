// mma_proto.cu -- measured prototype: the H pass of config 2 (RGB, 2x, a = 3) as u8 x s8 -> s32 tensor-core MMAs.
// (Round 2: the review asked for one measured prototype of the banded operator on the tensor cores before the claim "the
// kernel is bound by instruction dispatch, tensor cores would not help" is accepted.  This is that measurement; it is a
// development tool, not part of the library.)
//
// Formulation.  An interpolated sample of row r whose first tap is input byte f is sum_j w_j * in[r][f + 3 j], j = 0..5 (RGB:
// taps are 3 bytes apart; at 2x every interpolated sample has the same six weights).  For a window of K = 32 consecutive
// input bytes starting at b0 the samples f = b0 .. b0 + 15 have all their taps inside the window, so
//     D[16 rows][16 samples] = A[16 rows][32 bytes] * B[32 bytes][16 samples],   B[k][n] = w_j if k = n + 3 j else 0
// is two mma.sync.m16n8k32 (N = 8 each); the next window starts 16 bytes on and re-uses half of the A fragment.  The
// weights are fixed point with 22 fractional bits in three balanced base-256 digits (s8), one MMA per digit plane:
// six IMMAs per 16 x 16 samples, combined as d2 * 65536 + d1 * 256 + d0 -- an EXACT integer dot product.  A sample whose
// fractional part is within G = 6 * 255 * 2^-23 (the weight rounding) of an integer is flagged (it would take the exact
// double path of the library); the others are truncated and clamped like full_TB.h:29-37.
//
// What is timed (16 one-warp CTAs per SM like lanczos_v6_kernel, rows resident in shared memory as after a TMA stage):
//   mode 0: MMAs + integer epilogue, interpolated bytes written to shared memory compactly (2-byte stores)
//   mode 1: the same + the copies and the 3-byte interleave of the real ring layout with byte stores
//   mode 2: MMAs only (accumulators XOR-ed into one register): the tensor-pipe share
// and, for scale, the library's own H pass: 398 M interpolated samples in 0.266 ms on 148 SMs (profiles/r02_ablation_*).
// Verification: every output byte against the integer dot product on the CPU, and against the reference's double sum for
// every sample that was not flagged.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_proto tools/mma_proto.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ROWS = 16;            // rows per tile (M)
constexpr int TILES = 16;           // windows per warp and pass: 16 * 16 = 256 interpolated samples per row (13 KB of shared memory per warp)
constexpr int ROWB = 16 * TILES + 16 + 16;   // staged bytes per row (+ halo, + pad)
constexpr int FRAC = 22;
constexpr int GUARD = 768;          // ceil(6 * 255 * 2^-23 * 2^22) + margin

__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// digits[plane][j]: balanced base-256 digit of tap j
template <int MODE>
__global__ void __launch_bounds__(32, 16) hpass_mma(const uint8_t *in, uint8_t *out, unsigned *flags, const int8_t *digits, int iters) {
    __shared__ __align__(16) uint8_t rows[ROWS][ROWB];
    __shared__ __align__(16) uint8_t ring[ROWS][MODE == 1 ? 2 * 16 * TILES + 32 : 16 * TILES + 16];
    const int t = threadIdx.x, g = t >> 2, q = t & 3;
    for (int i = t; i < ROWS * ROWB / 4; i += 32)
        reinterpret_cast<uint32_t *>(&rows[0][0])[i] = reinterpret_cast<const uint32_t *>(in + (size_t)blockIdx.x % 64 * ROWS * ROWB)[i];
    __syncwarp();
    // B fragments: thread holds B[k = 4q + i (+16)][n = g] for both n halves (sample n + 8 h): w_j if k == n + 8 h + 3 j
    uint32_t bw[3][2][2];
    for (int pl = 0; pl < 3; pl++)
        for (int h = 0; h < 2; h++)
            for (int kh = 0; kh < 2; kh++) {
                uint32_t w = 0;
                for (int i = 0; i < 4; i++) {
                    const int k = 4 * q + i + 16 * kh, d = k - (g + 8 * h);
                    int8_t v = 0;
                    if (d >= 0 && d % 3 == 0 && d / 3 < 6) v = digits[pl * 6 + d / 3];
                    w |= (uint32_t)(uint8_t)v << (8 * i);
                }
                bw[pl][h][kh] = w;
            }
    unsigned doubt = 0, sink = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t a[4];
        a[2] = *reinterpret_cast<const uint32_t *>(&rows[g][4 * q]);
        a[3] = *reinterpret_cast<const uint32_t *>(&rows[g + 8][4 * q]);
#pragma unroll 2
        for (int tile = 0; tile < TILES; tile++) {
            a[0] = a[2]; a[1] = a[3];
            a[2] = *reinterpret_cast<const uint32_t *>(&rows[g][16 * tile + 16 + 4 * q]);
            a[3] = *reinterpret_cast<const uint32_t *>(&rows[g + 8][16 * tile + 16 + 4 * q]);
            int d[3][2][4];
#pragma unroll
            for (int pl = 0; pl < 3; pl++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
#pragma unroll
                    for (int i = 0; i < 4; i++) d[pl][h][i] = 0;
                    imma(d[pl][h], a, bw[pl][h][0], bw[pl][h][1]);
                }
            if (MODE == 2) {
#pragma unroll
                for (int pl = 0; pl < 3; pl++)
#pragma unroll
                    for (int h = 0; h < 2; h++)
#pragma unroll
                        for (int i = 0; i < 4; i++) sink ^= (unsigned)d[pl][h][i];
                continue;
            }
            // epilogue: combine the digit planes, flag what is within GUARD of an integer, truncate + clamp, pack
            unsigned mn = 0xffffffffu;
            uint32_t pk[2][2];                   // [row half][n half]: two adjacent samples as bytes
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int rh = 0; rh < 2; rh++) {
                    int qv[2];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int i = 2 * rh + e;
                        const int acc = d[2][h][i] * 65536 + d[1][h][i] * 256 + d[0][h][i];
                        mn = min(mn, (unsigned)(acc + GUARD) & ((1u << FRAC) - 1u));
                        qv[e] = acc >> FRAC;
                    }
                    uint32_t p2;
                    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(p2) : "r"(qv[1]), "r"(qv[0]));
                    pk[rh][h] = p2;
                }
            doubt |= (mn < 2u * GUARD) ? 1u : 0u;
            if (MODE == 0) {
                // compact: sample n of the tile at ring[row][16 tile + n]
#pragma unroll
                for (int rh = 0; rh < 2; rh++)
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        *reinterpret_cast<uint16_t *>(&ring[g + 8 * rh][16 * tile + 8 * h + 2 * q]) = (uint16_t)pk[rh][h];
            } else {
                // ring layout of the real kernel: output pixel 2p = copy of input pixel p, 2p + 1 = interpolated.
                // Sample with first tap f sits between input bytes f + 6 and f + 9: output byte 2 (f + 6) + 3 - (f + 6) % 3 ... in
                // byte terms: input byte i -> output byte 6 (i / 3) + i % 3; sample f -> 6 ((f + 6) / 3) + 3 + f % 3
#pragma unroll
                for (int rh = 0; rh < 2; rh++) {
#pragma unroll
                    for (int h = 0; h < 2; h++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int f = 16 * tile + 8 * h + 2 * q + e;
                            ring[g + 8 * rh][6 * ((f + 6) / 3) + 3 + f % 3 - 12] = (uint8_t)(pk[rh][h] >> (8 * e));
                        }
                    // the copies: this thread's four new input bytes of the row (a[2] / a[3])
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int i = 16 * tile + 16 + 4 * q + e;
                        ring[g + 8 * rh][6 * (i / 3) + i % 3 - 24] = (uint8_t)(a[2 + rh] >> (8 * e));
                    }
                }
            }
        }
        __syncwarp();
    }
    if (doubt) atomicOr(flags, 1u);
    if (MODE == 2 && sink == 0x12345u) atomicOr(flags, 2u);
    // the last pass's results of block 0 go out for verification
    if (blockIdx.x == 0 && MODE != 2)
        for (int i = t; i < (int)sizeof(ring); i += 32) out[i] = (&ring[0][0])[i];
}

static double sinc(double x) { return x == 0 ? 1.0 : std::sin(x) / x; }

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 200;
    // weights of phase 1/2 at a = 3: taps at distance 2.5, 1.5, 0.5, -0.5, -1.5, -2.5 (full_TB.h:39-53)
    double w[6];
    int W[6];
    int8_t digits[18];
    for (int j = 0; j < 6; j++) {
        const double x = 2.5 - j;
        w[j] = sinc(M_PI * x) * sinc(M_PI * x / 3.0);
        W[j] = (int)std::llround(w[j] * (1 << FRAC));
        int rem = W[j];
        for (int pl = 0; pl < 3; pl++) {                 // balanced base-256 digits, least significant first
            int dgt = ((rem % 256) + 256) % 256;
            if (dgt >= 128) dgt -= 256;
            digits[pl * 6 + j] = (int8_t)dgt;
            rem = (rem - dgt) / 256;
        }
        if (rem != 0) { printf("weight %d does not fit three digits\n", j); return 1; }
    }
    std::vector<uint8_t> h_in((size_t)64 * ROWS * ROWB);
    uint32_t s = 12345;
    for (auto &b : h_in) { s = s * 1664525u + 1013904223u; b = (uint8_t)(s >> 24); }
    uint8_t *d_in, *d_out; unsigned *d_flags; int8_t *d_digits;
    CK(cudaMalloc(&d_in, h_in.size())); CK(cudaMalloc(&d_out, 1 << 16)); CK(cudaMalloc(&d_flags, 4)); CK(cudaMalloc(&d_digits, 18));
    CK(cudaMemcpy(d_in, h_in.data(), h_in.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_digits, digits, 18, cudaMemcpyHostToDevice));
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; mode++) {
        CK(cudaMemset(d_flags, 0, 4));
        CK(cudaMemset(d_out, 0, 1 << 16));
        auto launch = [&](int n) {
            if (mode == 0) hpass_mma<0><<<grid, 32>>>(d_in, d_out, d_flags, d_digits, n);
            else if (mode == 1) hpass_mma<1><<<grid, 32>>>(d_in, d_out, d_flags, d_digits, n);
            else hpass_mma<2><<<grid, 32>>>(d_in, d_out, d_flags, d_digits, n);
        };
        launch(2);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        const double samples = (double)grid * iters * TILES * 256.0;
        printf("mode %d: %.3f ms for %.0f M interpolated samples: %.2f G samples/s per SM (library H pass: 10.1 incl. copies, filter, splice)\n",
               mode, best, samples / 1e6, samples / (best * 1e-3) / sms / 1e9);
        if (mode == 2) continue;
        // verification of block 0 (rows of input block 0)
        std::vector<uint8_t> h_out(1 << 16);
        unsigned h_flags = 0;
        CK(cudaMemcpy(h_out.data(), d_out, 1 << 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&h_flags, d_flags, 4, cudaMemcpyDeviceToHost));
        const int ringb = mode == 1 ? 2 * 16 * TILES + 32 : 16 * TILES + 16;
        long bad_int = 0, bad_ref = 0, flagged = 0, copies_bad = 0;
        for (int r = 0; r < ROWS; r++)
            for (int f = 0; f < 16 * TILES; f++) {
                long long acc = 0;
                double ref = 0.0;
                for (int j = 0; j < 6; j++) {
                    const int b = h_in[(size_t)r * ROWB + f + 3 * j];
                    acc += (long long)W[j] * b;
                    ref = ref + (double)b * w[j];
                }
                long long qi = acc >> FRAC;
                qi = qi < 0 ? 0 : (qi > 255 ? 255 : qi);
                const bool fl = (unsigned)((acc + GUARD) & ((1 << FRAC) - 1)) < 2u * GUARD;
                const int qr = ref < 0 ? 0 : (ref > 255 ? 255 : (int)ref);
                const int pos = mode == 1 ? 6 * ((f + 6) / 3) + 3 + f % 3 - 12 : f;
                const int got = h_out[(size_t)r * ringb + pos];
                if (got != qi) bad_int++;
                if (fl) flagged++;
                else if (got != qr) bad_ref++;
            }
        if (mode == 1)
            for (int r = 0; r < ROWS; r++)
                for (int i = 16; i < 16 * TILES + 16; i++)      // (bytes 0..15 are the first window's low half: never new)
                    if (h_out[(size_t)r * ringb + 6 * (i / 3) + i % 3 - 24] != h_in[(size_t)r * ROWB + i]) copies_bad++;
        printf("   block 0: %ld of %d samples differ from the integer dot product, %ld unflagged samples differ from the double sum, "
               "%ld flagged (%.3f %%), %ld copies wrong, device doubt flag %u\n",
               bad_int, ROWS * 16 * TILES, bad_ref, flagged, 100.0 * flagged / (ROWS * 16 * TILES), copies_bad, h_flags);
    }
    return 0;
}
