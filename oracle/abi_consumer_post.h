/* abi_consumer_post.h -- epilogue of the compiled C-ABI consumer.
 * oracle/Makefile pipes ref_shim_pre.h, `sed -n 29,96p full_TB.h` (the reference's own lanczos_expected, read
 * where it lies) and this file into g++ and links the result against liblanczos_b200.so: one translation unit
 * in which the reference's static planar arrays (full_TB.h:20-21) are handed, as they are, to the reference's
 * function and to the binding INTEGRATION.md section 2 tells a maintainer to add.  Struct layout, linkage and
 * argument meaning of include/lanczos_b200.h are therefore checked by a C++ compiler, not by a ctypes mirror.
 * Test infrastructure: built into oracle/_ref/, run by tests/test_abi_consumer.py under -m gpu. */
#include "lanczos_b200.h"

static byte img_in[NUM_CHANNELS][IN_HEIGHT][IN_WIDTH];             /* full_TB.h:20 */
static byte img_out_ex[NUM_CHANNELS][OUT_HEIGHT][OUT_WIDTH];       /* full_TB.h:21 */
static byte img_out_b200[NUM_CHANNELS][OUT_HEIGHT][OUT_WIDTH];

/* INTEGRATION.md section 2, verbatim */
void lanczos_expected_b200(byte img_in[NUM_CHANNELS][IN_HEIGHT][IN_WIDTH],
                           byte img_out[NUM_CHANNELS][OUT_HEIGHT][OUT_WIDTH]) {
    lanczos_desc d = {};
    d.in_w = IN_WIDTH;   d.in_h = IN_HEIGHT;
    d.out_w = OUT_WIDTH; d.out_h = OUT_HEIGHT;
    d.channels = NUM_CHANNELS;
    d.a = LANCZOS_A;
    d.scale_n = SCALE_N; d.scale_d = SCALE_D;
    int rc = lanczos_b200_expected(&d, (const uint8_t *)img_in, (uint8_t *)img_out, /*device=*/0);
    if (rc != LANCZOS_OK) { printf("lanczos_b200: %s (%s)\n", lanczos_b200_strerror(rc), lanczos_b200_last_cuda_error()); exit(EXIT_FAILURE); }
}

int main(int argc, char **argv) {
    /* synthetic input in place of the PNG load of sim_tb (full_TB.h:127-138): xorshift64 noise, or dark noise 0..15 */
    const int dark = argc > 1 && !strcmp(argv[1], "dark");
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int c = 0; c < NUM_CHANNELS; c++)
        for (int y = 0; y < IN_HEIGHT; y++)
            for (int x = 0; x < IN_WIDTH; x++) {
                s ^= s << 13; s ^= s >> 7; s ^= s << 17;
                img_in[c][y][x] = (byte)(dark ? (s >> 24) & 15 : (s >> 24) & 255);
            }
    if (lanczos_b200_abi_version() != LANCZOS_B200_ABI_VERSION) { printf("ABI version mismatch\n"); return 2; }
    lanczos_expected(img_in, img_out_ex);            /* full_TB.h:141 */
    lanczos_expected_b200(img_in, img_out_b200);     /* the drop-in */
    const int differ = memcmp(img_out_ex, img_out_b200, sizeof(img_out_ex)) != 0;
    long n = 0;
    for (size_t i = 0; i < sizeof(img_out_ex); i++) n += ((const uint8_t *)img_out_ex)[i] != ((const uint8_t *)img_out_b200)[i];
    printf("abi_consumer %dx%d -> %dx%d c=%d a=%d %d/%d %s: %ld of %zu bytes differ\n", IN_WIDTH, IN_HEIGHT, OUT_WIDTH, OUT_HEIGHT,
           NUM_CHANNELS, LANCZOS_A, SCALE_N, SCALE_D, dark ? "dark" : "noise", n, sizeof(img_out_ex));
    return differ ? 1 : 0;
}
