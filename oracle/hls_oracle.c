/*
 * hls_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the reference's fixed-point HLS path.
 *
 * PINNING: the per-sample arithmetic below (oracle_hls_mac1 / _mac2 / _to_byte = worker.cpp compute / compute_ /
 * clamp_to_byte) is pinned to the reference's OWN code: oracle/Makefile compiles worker.cpp:10-130 as it is
 * against the integer-backed ap_fixed/ap_uint stand-ins of oracle/ap_shim.h into oracle/_ref/libref_hls_*.so and
 * tests/test_hls_mode.py::test_sample_arithmetic_against_compiled_reference compares the two on random windows,
 * every LUT phase and random kernel values.  The image loops of oracle_hls_upscale go through the same three helpers.
 * What stays UNPINNED is the LUT content: the reference evaluates it with Xilinx hls::sinpi in fixed point
 * (kernel.cpp:12-18), which is neither in the reference tree nor installed, and the reference holds no golden
 * vector for it; the border rules and the window stepping are restated from the sources.
 *
 * Restated (reference LanczosUpscaler/):
 *   kernel.cpp:40-45  init_lanczos_kernel: ROM[i] = L(kernel_t(i)/SCALE_N), i < A*N; ROM[A*N] = 0.
 *                     kernel_t = ap_fixed<8+BP,8>, AP_TRN: the argument is floor(i*2^BP/N)/2^BP and
 *                     the stored value floor(L*2^BP)/2^BP.  L itself is evaluated in double here
 *                     (the reference uses hls::sinpi in fixed point: UNPINNED).
 *   kernel.cpp:50-67  weight = LUT[|out_idx*SCALE_D - in_idx*SCALE_N|]
 *   lanczos.cpp:96    vertical ("column lengthening") pass first, then horizontal
 *   worker.cpp:45-78  compute : acc = sum_j k_j * v_j exactly (BP fraction bits), clamped to
 *                     [min,max] of the two central taps A-1, A (de-ring)
 *   worker.cpp:81-115 compute_: same on the fixed-point intermediates, each product floored to BP
 *                     fraction bits by the += into num_el_t (AP_TRN)
 *   worker.cpp:118-130 clamp_to_byte: byte = raw >> BP
 *   worker.cpp:170-198, cyclic_buffer.h:30-42  vertical window: A-1 zero rows above the image,
 *                     the last row replicated below it
 *   worker.cpp:239-275 horizontal window: zeros on the left, last column replicated on the right
 * Valid for integer scales (SCALE_D = 1): for other ratios the reference's BP-bit step condition
 * (worker.cpp:140,234) drifts from the ideal window (SURVEY.md 8a "HLS-path validity limit").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static double sinc_d(double x) { return x == 0 ? 1 : sin(x) / x; }

/* lut must hold a*n+1 entries; values in units of 2^-bp */
int oracle_hls_lut(int a, int n, int bp, int32_t *lut) {
    if (a < 1 || n < 1 || bp < 1 || bp > 12 || a * n > 127) return -1; /* kernel_t(i) wraps for i >= 128 */
    for (int i = 0; i < a * n; i++) {
        const double x = floor((double)i * (1 << bp) / n) / (1 << bp);
        const double l = sinc_d(M_PI * x) * sinc_d(M_PI * x / a);
        lut[i] = (int32_t)floor(l * (1 << bp));
    }
    lut[a * n] = 0;
    return 0;
}

static int iabs_(int v) { return v < 0 ? -v : v; }

/* num_el_t = ap_fixed<10+bp, 10>: AP_WRAP drops integer bits beyond 10 (never reached with Lanczos weights:
 * sum |w| <= 1.6, so |acc| <= 410 < 512; kept so that the pin against the compiled reference holds for any kernel) */
static int32_t wrap_num(int64_t x, int bp) {
    const int w = 10 + bp;
    const uint64_t m = (1ull << w) - 1;
    uint64_t u = (uint64_t)x & m;
    if ((u >> (w - 1)) & 1) u |= ~m;
    return (int32_t)(int64_t)u;
}
static int32_t dering(int32_t acc, int32_t c0, int32_t c1) {      /* worker.cpp:63-74 / :100-111 */
    const int32_t lo = c0 < c1 ? c0 : c1, hi = c0 < c1 ? c1 : c0;
    return acc < lo ? lo : (acc > hi ? hi : acc);
}
/* worker.cpp:45-78 compute, one channel: v[2a] bytes, k[2a] kernel_t raw (units 2^-bp) -> num_el_t raw.
 * k * v is exact in ap_fixed arithmetic (bp fraction bits), acc += wraps to 10 integer bits. */
int32_t oracle_hls_mac1(const uint8_t *v, const int32_t *k, int a, int bp) {
    int32_t acc = 0;
    for (int j = 0; j < 2 * a; j++) acc = wrap_num((int64_t)acc + (int64_t)k[j] * v[j], bp);
    return dering(acc, (int32_t)v[a - 1] << bp, (int32_t)v[a] << bp);
}
/* worker.cpp:81-115 compute_, one channel: v[2a] num_el_t raw -> num_el_t raw.  kern * in has 2 bp fraction bits;
 * the += into num_el_t (AP_TRN) floors the sum, i.e. each product, to bp fraction bits. */
int32_t oracle_hls_mac2(const int32_t *v, const int32_t *k, int a, int bp) {
    int32_t acc = 0;
    for (int j = 0; j < 2 * a; j++) acc = wrap_num((int64_t)acc + (((int64_t)k[j] * v[j]) >> bp), bp);
    return dering(acc, v[a - 1], v[a]);
}
/* worker.cpp:118-130 clamp_to_byte: byte_el_t(num_el_t) = integer part like a C cast (toward zero), wrapped to 8 bits.
 * After the de-ring clamp the value is >= 0 and < 256, where this is raw >> bp. */
uint8_t oracle_hls_to_byte(int32_t raw, int bp) {
    int32_t q = raw >> bp;
    if (raw < 0 && (raw & ((1 << bp) - 1)) != 0) q += 1;
    return (uint8_t)(q & 0xff);
}

/* interleaved uint8 in/out, integer scale n (out = in * n) */
int oracle_hls_upscale(const uint8_t *in, uint8_t *out, int channels, int in_w, int in_h, int out_w,
                       int out_h, int a, int n, int bp) {
    if (channels < 1 || in_w < 1 || in_h < 1 || out_w < 1 || out_h < 1) return -1;
    int32_t lut[128];
    if (oracle_hls_lut(a, n, bp, lut)) return -1;
    const int taps = 2 * a;
    /* vertical pass: mid[y][x*C+c], fixed point with bp fraction bits, >= 0 after the clamp */
    int32_t *mid = (int32_t *)malloc(sizeof(int32_t) * (size_t)out_h * in_w * channels);
    if (!mid) return -2;
    for (int y = 0; y < out_h; y++) {
        const int base = y / n; /* floor(y * D / N), D = 1 */
        int32_t k[8];
        for (int j = 0; j < taps; j++) k[j] = lut[iabs_(y - (base - a + 1 + j) * n)];   /* nominal row indexes the LUT, also when replicated */
        for (int xb = 0; xb < in_w * channels; xb++) {
            uint8_t v[8];
            for (int j = 0; j < taps; j++) {
                const int row = base - a + 1 + j;
                v[j] = row >= 0 ? in[(size_t)(row < in_h ? row : in_h - 1) * in_w * channels + xb] : 0;
            }
            mid[(size_t)y * in_w * channels + xb] = oracle_hls_mac1(v, k, a, bp);
        }
    }
    /* horizontal pass */
    for (int y = 0; y < out_h; y++) {
        const int32_t *m = mid + (size_t)y * in_w * channels;
        for (int x = 0; x < out_w; x++) {
            const int base = x / n;
            int32_t k[8];
            for (int j = 0; j < taps; j++) k[j] = lut[iabs_(x - (base - a + 1 + j) * n)];
            for (int c = 0; c < channels; c++) {
                int32_t v[8];
                for (int j = 0; j < taps; j++) {
                    const int col = base - a + 1 + j;
                    v[j] = col >= 0 ? m[(size_t)(col < in_w ? col : in_w - 1) * channels + c] : 0;
                }
                out[((size_t)y * out_w + x) * channels + c] = oracle_hls_to_byte(oracle_hls_mac2(v, k, a, bp), bp);
            }
        }
    }
    free(mid);
    return 0;
}
