/* ref_hls_pre.h -- prologue for compiling the reference's fixed-point sample arithmetic
 * (/root/reference/LanczosUpscaler/worker.cpp lines 10-130) *where it lies*, against oracle/ap_shim.h.
 * oracle/Makefile pipes: this file, `sed -n 10,130p worker.cpp`, ref_hls_post.h.  No reference source is stored here.
 * Restated configuration (the reference takes it from its git-ignored params.h, template lanczos.h:9-31):
 *   MIN/MAX lanczos.h:63-64, INTEGER_BITS lanczos.h:74, the five typedefs lanczos.h:79-82,90-91,
 *   cyclic_buffer_t::Slice (cyclic_buffer.h:49-61) reduced to what compute() uses: operator[] over 2a packed pixels. */
#include <stdint.h>
#include <string.h>
#include "ap_shim.h"
#if !defined(NUM_CHANNELS) || !defined(LANCZOS_A) || !defined(BIT_PRECISION)
#error "pass -DNUM_CHANNELS= -DLANCZOS_A= -DBIT_PRECISION="
#endif
#define MIN(a,b) ((a)<(b)?(a):(b))
#define MAX(a,b) ((a)>(b)?(a):(b))
#define INTEGER_BITS 10
typedef ap_uint<8> byte_el_t;
typedef ap_fixed<INTEGER_BITS+BIT_PRECISION, INTEGER_BITS> num_el_t;
typedef ap_fixed<8+BIT_PRECISION, 8> kernel_t;
typedef ap_uint<8*NUM_CHANNELS> byte_t;
typedef ap_uint<(INTEGER_BITS+BIT_PRECISION)*NUM_CHANNELS> num_t;
struct cyclic_buffer_t {
    struct Slice {
        byte_t taps[2 * LANCZOS_A];
        byte_t &operator[](int i) { return taps[i]; }
    };
};
