/*
 * hls_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the reference's fixed-point HLS path.
 *
 * PARITY UNPINNED: the HLS path needs Xilinx ap_fixed.h / hls_math.h (hls::sinpi), which are not in
 * the reference tree and not installed, and the reference holds no golden vector for it
 * (SURVEY.md 8c).  This file restates the integer arithmetic the sources specify and documents
 * the one place that cannot be pinned (the LUT values produced by hls::sinpi).
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
        for (int xb = 0; xb < in_w * channels; xb++) {
            int32_t acc = 0, c0 = 0, c1 = 0;
            for (int j = 0; j < taps; j++) {
                const int row = base - a + 1 + j; /* nominal row: also indexes the LUT when replicated */
                int32_t v = 0;
                if (row >= 0) v = in[(size_t)(row < in_h ? row : in_h - 1) * in_w * channels + xb];
                acc += lut[iabs_(y - row * n)] * v;
                if (j == a - 1) c0 = v << bp;
                if (j == a) c1 = v << bp;
            }
            const int32_t lo = c0 < c1 ? c0 : c1, hi = c0 < c1 ? c1 : c0;
            mid[(size_t)y * in_w * channels + xb] = acc < lo ? lo : (acc > hi ? hi : acc);
        }
    }
    /* horizontal pass */
    for (int y = 0; y < out_h; y++) {
        const int32_t *m = mid + (size_t)y * in_w * channels;
        for (int x = 0; x < out_w; x++) {
            const int base = x / n;
            for (int c = 0; c < channels; c++) {
                int32_t acc = 0, c0 = 0, c1 = 0;
                for (int j = 0; j < taps; j++) {
                    const int col = base - a + 1 + j;
                    int32_t v = 0;
                    if (col >= 0) v = m[(size_t)(col < in_w ? col : in_w - 1) * channels + c];
                    /* floor to bp fraction bits (arithmetic shift of the signed product) */
                    acc += (int32_t)(((int64_t)lut[iabs_(x - col * n)] * v) >> bp);
                    if (j == a - 1) c0 = v;
                    if (j == a) c1 = v;
                }
                const int32_t lo = c0 < c1 ? c0 : c1, hi = c0 < c1 ? c1 : c0;
                acc = acc < lo ? lo : (acc > hi ? hi : acc);
                out[((size_t)y * out_w + x) * channels + c] = (uint8_t)(acc >> bp);
            }
        }
    }
    free(mid);
    return 0;
}
