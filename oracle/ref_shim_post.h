/* ref_shim_post.h -- C entry points around the reference's lanczos_expected()
 * (full_TB.h:79-96), planar byte[C][H][W] in and out like its static arrays
 * (full_TB.h:20-21). */
extern "C" int ref_lanczos_expected(const uint8_t *in_planar, uint8_t *out_planar) {
    memset(out_planar, 0, (size_t)NUM_CHANNELS * OUT_HEIGHT * OUT_WIDTH); /* zero-initialised global */
    lanczos_expected((byte (*)[IN_HEIGHT][IN_WIDTH])in_planar,
                     (byte (*)[OUT_HEIGHT][OUT_WIDTH])out_planar);
    return 0;
}
extern "C" void ref_config(int *cfg) {
    cfg[0] = IN_WIDTH; cfg[1] = IN_HEIGHT; cfg[2] = OUT_WIDTH; cfg[3] = OUT_HEIGHT;
    cfg[4] = SCALE_N; cfg[5] = SCALE_D; cfg[6] = LANCZOS_A; cfg[7] = NUM_CHANNELS;
}
