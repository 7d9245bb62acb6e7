/* ref_shim_pre.h -- prologue for compiling the reference software path
 * (/root/reference/LanczosUpscaler/full_TB.h lines 29-96) *where it lies*.
 * oracle/Makefile pipes: this file, then `sed -n 29,96p full_TB.h`, then
 * ref_shim_post.h into g++.  No reference source is stored in this repo.
 * The reference needs Xilinx ap_uint<8> only as 8-bit storage (full_TB.h:18),
 * and the size macros normally come from its git-ignored params.h
 * (template lanczos.h:9-31, MIN/MAX lanczos.h:63-64, SCALE lanczos.h:112). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
typedef uint8_t byte;
#define MIN(a,b) ((a)<(b)?(a):(b))
#define MAX(a,b) ((a)>(b)?(a):(b))
#define SCALE ((double)SCALE_N/SCALE_D)
#if !defined(IN_WIDTH) || !defined(IN_HEIGHT) || !defined(OUT_WIDTH) || !defined(OUT_HEIGHT) || \
    !defined(NUM_CHANNELS) || !defined(LANCZOS_A) || !defined(SCALE_N) || !defined(SCALE_D)
#error "pass -DIN_WIDTH= -DIN_HEIGHT= -DOUT_WIDTH= -DOUT_HEIGHT= -DNUM_CHANNELS= -DLANCZOS_A= -DSCALE_N= -DSCALE_D="
#endif
