/* ref_hls_post.h -- C entry points around the reference's compute / compute_ / clamp_to_byte (worker.cpp:45-130).
 * Values cross the boundary as raw integers: bytes, kernel_t raw (units of 2^-BIT_PRECISION), num_el_t raw (the same). */
extern "C" void ref_hls_config(int *cfg) { cfg[0] = NUM_CHANNELS; cfg[1] = LANCZOS_A; cfg[2] = BIT_PRECISION; }

static kernel_t kernel_from_raw(int32_t raw) {
    kernel_t k;
    k.set_bits((uint64_t)(int64_t)raw);
    return k;
}

/* first pass (worker.cpp:45-78): taps[2a][C] bytes, kern[2a] raw -> out[C] num_el_t raw */
extern "C" void ref_hls_compute(const uint8_t *taps, const int32_t *kern_raw, int32_t *out_raw) {
    cyclic_buffer_t::Slice s;
    kernel_t kern[2 * LANCZOS_A];
    for (int i = 0; i < 2 * LANCZOS_A; i++) {
        byte_el_t px[NUM_CHANNELS];
        for (int c = 0; c < NUM_CHANNELS; c++) px[c] = byte_el_t((int)taps[i * NUM_CHANNELS + c]);
        s.taps[i] = pack_blob(px);
        kern[i] = kernel_from_raw(kern_raw[i]);
    }
    num_t r = compute(s, kern);
    num_el_t out[NUM_CHANNELS];
    unpack_blob(r, out);
    for (int c = 0; c < NUM_CHANNELS; c++) out_raw[c] = (int32_t)out[c].raw;
}

/* second pass (worker.cpp:81-115): in[2a][C] num_el_t raw, kern[2a] raw -> out[C] num_el_t raw */
extern "C" void ref_hls_compute2(const int32_t *in_raw, const int32_t *kern_raw, int32_t *out_raw) {
    num_t in[2 * LANCZOS_A];
    kernel_t kern[2 * LANCZOS_A];
    for (int i = 0; i < 2 * LANCZOS_A; i++) {
        num_el_t px[NUM_CHANNELS];
        for (int c = 0; c < NUM_CHANNELS; c++) px[c].set_bits((uint64_t)(int64_t)in_raw[i * NUM_CHANNELS + c]);
        in[i] = pack_blob(px);
        kern[i] = kernel_from_raw(kern_raw[i]);
    }
    num_t r = compute_(in, kern);
    num_el_t out[NUM_CHANNELS];
    unpack_blob(r, out);
    for (int c = 0; c < NUM_CHANNELS; c++) out_raw[c] = (int32_t)out[c].raw;
}

/* worker.cpp:118-130: num_el_t raw [C] -> bytes [C] */
extern "C" void ref_hls_clamp_to_byte(const int32_t *raw, uint8_t *out) {
    num_el_t px[NUM_CHANNELS];
    for (int c = 0; c < NUM_CHANNELS; c++) px[c].set_bits((uint64_t)(int64_t)raw[c]);
    byte_t b = clamp_to_byte(pack_blob(px));
    byte_el_t o[NUM_CHANNELS];
    unpack_blob(b, o);
    for (int c = 0; c < NUM_CHANNELS; c++) out[c] = (uint8_t)o[c].v;
}
