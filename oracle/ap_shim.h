/* ap_shim.h -- minimal integer-backed stand-ins for Xilinx ap_uint<W> / ap_fixed<W,I> (TEST INFRASTRUCTURE ONLY).
 *
 * Purpose: compile the reference's fixed-point MAC, de-ring clamp and byte conversion (worker.cpp:10-130:
 * unpack_blob, pack_blob, compute, compute_, clamp_to_byte) AS THEY ARE, read where they lie in the reference
 * tree, without the Xilinx headers (ap_fixed.h / ap_int.h are neither vendored nor installed).  oracle/Makefile
 * pipes ref_hls_pre.h, `sed -n 10,130p worker.cpp` and ref_hls_post.h into g++ (outputs under oracle/_ref/).
 * tests/test_hls_mode.py then pins oracle/hls_oracle.c's per-sample arithmetic to that compiled reference code.
 * What stays unpinned is only the LUT content (kernel.cpp:12-18 uses hls::sinpi).
 *
 * Semantics implemented (Vivado HLS ap_fixed.h defaults, which the reference uses: lanczos.h:79-82 give no Q/O modes):
 *   ap_fixed<W,I>   signed, value = raw * 2^-(W-I), raw kept sign-extended in an int64
 *   quantisation    AP_TRN: extra fraction bits are dropped, i.e. truncation toward minus infinity
 *   overflow        AP_WRAP: extra integer bits are dropped (two's complement wrap)
 *   a * b, a + b    exact (the result type of the Xilinx operators is wide enough); modelled by `fxv`, an exact
 *                   (raw, fraction bits) pair, quantised only when it is assigned to a declared type
 *   a += b          a = a + b, quantised to a's type
 *   ap_uint<W>(ap_fixed)  integer part like a C cast (toward zero), then wrapped to W bits
 *   x(hi, lo)       bit range of the raw value, readable and assignable
 * Widths up to 63 bits (num_t = ap_uint<(10+BP)*NUM_CHANNELS>: 54 bits for 3 channels at BP = 8). */
#ifndef AP_SHIM_H
#define AP_SHIM_H
#include <stdint.h>

struct fxv {              /* exact fixed-point value: raw * 2^-f */
    int64_t raw;
    int f;
};
static inline fxv fx_align(fxv a, int f) { fxv r = {a.raw << (f - a.f), f}; return r; }   /* f >= a.f */
static inline fxv operator*(fxv a, fxv b) { fxv r = {a.raw * b.raw, a.f + b.f}; return r; }
static inline fxv operator+(fxv a, fxv b) {
    const int f = a.f > b.f ? a.f : b.f;
    fxv r = {fx_align(a, f).raw + fx_align(b, f).raw, f};
    return r;
}
static inline int fx_cmp(fxv a, fxv b) {
    const int f = a.f > b.f ? a.f : b.f;
    const int64_t x = fx_align(a, f).raw, y = fx_align(b, f).raw;
    return x < y ? -1 : (x > y ? 1 : 0);
}
static inline bool operator<(fxv a, fxv b) { return fx_cmp(a, b) < 0; }
static inline bool operator>(fxv a, fxv b) { return fx_cmp(a, b) > 0; }
static inline bool operator==(fxv a, fxv b) { return fx_cmp(a, b) == 0; }

template <class T>
struct ap_range_ref {     /* x(hi, lo) */
    T *obj;
    int hi, lo;
    uint64_t get() const { return (obj->bits() >> lo) & ((hi - lo + 1) >= 64 ? ~0ull : ((1ull << (hi - lo + 1)) - 1)); }
    operator uint64_t() const { return get(); }
    void set(uint64_t v) {
        const uint64_t m = ((hi - lo + 1) >= 64 ? ~0ull : ((1ull << (hi - lo + 1)) - 1)) << lo;
        obj->set_bits((obj->bits() & ~m) | ((v << lo) & m));
    }
    template <class U> ap_range_ref &operator=(const ap_range_ref<U> &o) { set(o.get()); return *this; }
    ap_range_ref &operator=(const ap_range_ref &o) { set(o.get()); return *this; }
    ap_range_ref &operator=(uint64_t v) { set(v); return *this; }
};

template <int W>
struct ap_uint {
    uint64_t v;
    static uint64_t mask() { return W >= 64 ? ~0ull : ((1ull << W) - 1); }
    ap_uint() : v(0) {}
    ap_uint(int x) : v((uint64_t)(int64_t)x & mask()) {}
    ap_uint(unsigned x) : v((uint64_t)x & mask()) {}
    ap_uint(uint64_t x) : v(x & mask()) {}
    ap_uint(fxv x) {      /* C-like conversion: toward zero, then wrap */
        int64_t q = x.raw >> x.f;
        if (x.raw < 0 && (x.raw & ((1ll << x.f) - 1)) != 0) q += 1;
        v = (uint64_t)q & mask();
    }
    operator fxv() const { fxv r = {(int64_t)v, 0}; return r; }
    uint64_t bits() const { return v; }
    void set_bits(uint64_t b) { v = b & mask(); }
    ap_range_ref<ap_uint> operator()(int hi, int lo) { ap_range_ref<ap_uint> r = {this, hi, lo}; return r; }
    uint64_t to_uint64() const { return v; }
};

template <int W, int I>
struct ap_fixed {
    int64_t raw;          /* sign-extended W-bit two's complement */
    static int64_t wrap(int64_t x) {
        const uint64_t m = W >= 64 ? ~0ull : ((1ull << W) - 1);
        uint64_t u = (uint64_t)x & m;
        if (W < 64 && (u >> (W - 1)) & 1) u |= ~m;
        return (int64_t)u;
    }
    static int64_t quantise(fxv x) {      /* AP_TRN (floor), AP_WRAP */
        const int f = W - I;
        const int64_t q = x.f >= f ? (x.raw >> (x.f - f)) : (x.raw << (f - x.f));
        return wrap(q);
    }
    ap_fixed() : raw(0) {}
    ap_fixed(int x) : raw(quantise(fxv{(int64_t)x, 0})) {}
    ap_fixed(fxv x) : raw(quantise(x)) {}
    template <int W2> ap_fixed(ap_uint<W2> x) : raw(quantise((fxv)x)) {}
    template <int W2, int I2> ap_fixed(ap_fixed<W2, I2> x) : raw(quantise((fxv)x)) {}
    operator fxv() const { fxv r = {raw, W - I}; return r; }
    ap_fixed &operator+=(fxv x) { raw = quantise((fxv)(*this) + x); return *this; }
    uint64_t bits() const { return (uint64_t)raw & (W >= 64 ? ~0ull : ((1ull << W) - 1)); }
    void set_bits(uint64_t b) { raw = wrap((int64_t)b); }
    ap_range_ref<ap_fixed> operator()(int hi, int lo) { ap_range_ref<ap_fixed> r = {this, hi, lo}; return r; }
};

/* mixed operators the reference's expressions need (kernel_t * byte_el_t, kernel_t * num_el_t, comparisons, MIN/MAX of
 * two byte_el_t assigned to a num_el_t) */
template <int W, int I, int W2> static inline fxv operator*(ap_fixed<W, I> a, ap_uint<W2> b) { return (fxv)a * (fxv)b; }
template <int W, int I, int W2, int I2> static inline fxv operator*(ap_fixed<W, I> a, ap_fixed<W2, I2> b) { return (fxv)a * (fxv)b; }
template <int W, int I, int W2, int I2> static inline bool operator<(ap_fixed<W, I> a, ap_fixed<W2, I2> b) { return (fxv)a < (fxv)b; }
template <int W, int I, int W2, int I2> static inline bool operator>(ap_fixed<W, I> a, ap_fixed<W2, I2> b) { return (fxv)a > (fxv)b; }
template <int W, int W2> static inline bool operator<(ap_uint<W> a, ap_uint<W2> b) { return a.v < b.v; }
template <int W, int W2> static inline bool operator>(ap_uint<W> a, ap_uint<W2> b) { return a.v > b.v; }
#endif
