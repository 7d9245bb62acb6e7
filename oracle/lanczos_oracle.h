/*
 * lanczos_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * Plain-C restatement of the reference *software path* `lanczos_expected()`
 * (reference LanczosUpscaler/full_TB.h:29-96).  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load
 * this library; the product (lanczos_hls_b200/) never links or calls it.
 *
 * Parity pinning: the reference ships no golden vectors.  This
 * oracle is pinned against (1) the reference itself, compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/, (2) the golden fixtures
 * in tests/golden/ that were generated from that compiled reference, and
 * (3) the FNV-1a known-answer hashes of SURVEY.md Appendix A.
 */
#ifndef LANCZOS_ORACLE_H
#define LANCZOS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Variant selectors for the vertical pass. */
enum {
    ORACLE_VERBATIM = 0, /* in-place bottom-up column pass, full_TB.h:67-77 (top rows alias) */
    ORACLE_CLEAN    = 1  /* same arithmetic, but reads the untouched H-pass plane (ping-pong) */
};

/* Literal restatement, planar layout byte[C][H][W], per-tap libm sin().
 * Follows full_TB.h:79-96 loop for loop.  `variant` picks VERBATIM/CLEAN. */
int oracle_expected_planar(const uint8_t *in, uint8_t *out, int channels,
                           int in_w, int in_h, int out_w, int out_h,
                           int a, int scale_n, int scale_d, int variant);

/* Same results (bit-identical), but weights are evaluated once per output
 * coordinate instead of once per tap per row, and rows/columns are spread over
 * `threads` OpenMP threads (0 = all).  Used for big parity cases. */
int oracle_expected_planar_fast(const uint8_t *in, uint8_t *out, int channels,
                                int in_w, int in_h, int out_w, int out_h,
                                int a, int scale_n, int scale_d, int variant,
                                int threads);

/* Interleaved front end (what sim_tb does around the call, full_TB.h:127-138
 * and :146-165): pixel-interleaved rows with byte pitches, channel 0 first. */
int oracle_upscale_interleaved(const uint8_t *in, int64_t in_pitch,
                               uint8_t *out, int64_t out_pitch, int channels,
                               int in_w, int in_h, int out_w, int out_h,
                               int a, int scale_n, int scale_d, int variant,
                               int threads);

/* Row band of the full-image result: output rows [row0,row0+rows) of the image
 * described by the arguments, computed from the FULL input (the band result is
 * by definition the corresponding slice of the full result). */
int oracle_upscale_interleaved_rows(const uint8_t *in, int64_t in_pitch,
                                    uint8_t *out_band, int64_t out_pitch,
                                    int channels, int in_w, int in_h, int out_w,
                                    int out_h, int a, int scale_n, int scale_d,
                                    int variant, int threads, int row0, int rows);

/* Reference kernel value L(x) = sinc(pi x) sinc(pi x / a), full_TB.h:39-53. */
double oracle_kernel(double x, int a);

/* Synthetic inputs shared by tests and bench. */
void oracle_fill_xorshift(uint8_t *dst, int64_t n, uint64_t seed);
uint64_t oracle_fnv1a64(const uint8_t *p, int64_t n);

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
