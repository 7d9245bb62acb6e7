// plan.cpp -- descriptor validation, rational ratio, weight tables (host).
// See plan.h for the reference interfaces this replaces.
//
// Compiled with -ffp-contract=off: the double weights must be the very values the
// reference computes at full_TB.h:60 (`lanczos_kernel(x - i)` with x = (double)xx/SCALE).
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <numeric>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace lzb {

// full_TB.h:39-44
static double sinc_ref(double x) {
    if (x == 0) return 1;
    return std::sin(x) / x;
}

// full_TB.h:51-53 (LANCZOS_A is an int macro: M_PI*x/a divides by (double)a)
double ref_kernel(double x, int a) { return sinc_ref(M_PI * x) * sinc_ref(M_PI * x / a); }

uint32_t half2_bits(double x, bool away) {
    const double ax = std::fabs(x);
    uint32_t h = 0;
    if (ax >= 65504.0) h = away ? 0x7C00u : 0x7BFFu;
    else if (ax > 0) {
        int e;
        std::frexp(ax, &e);                    // ax = m * 2^e, m in [0.5, 1)
        const int ex = std::max(e - 1, -14);   // fp16 exponent (subnormals share -14)
        const double q = std::ldexp(ax, 10 - ex);              // in units of the fp16 spacing at this exponent
        const double m = away ? std::ceil(q) : std::floor(q);
        // normals: (ex + 15) << 10 | (m - 1024), a carry into the exponent included; subnormals: m itself
        h = (ax >= std::ldexp(1.0, -14)) ? (uint32_t)(((ex + 15) << 10) + ((int)m - 1024)) : (uint32_t)m;
    }
    if (x < 0) h |= 0x8000u;
    return h | (h << 16);
}

int resolve_desc(const lanczos_desc *in, lanczos_desc *out) {
    if (!in || !out) return LANCZOS_ERR_NULL;
    lanczos_desc d = *in;
    if (d.in_w < 1 || d.in_h < 1 || d.out_w < 1 || d.out_h < 1) return LANCZOS_ERR_DIMS;
    // 2^20 per side keeps every byte offset inside a row and every coordinate product in int32
    if (d.in_w > (1 << 20) || d.in_h > (1 << 20) || d.out_w > (1 << 20) || d.out_h > (1 << 20))
        return LANCZOS_ERR_DIMS;
    if (d.channels < 1 || d.channels > 4) return LANCZOS_ERR_CHANNELS;
    if (d.a < 1 || d.a > kMaxTaps / 2) return LANCZOS_ERR_TAPS;
    if (d.reserved != 0) return LANCZOS_ERR_DIMS;
    if (d.scale_n == 0 && d.scale_d == 0) {
        // lanczos.h:110 SCALE_GCD = gcd(OUT_WIDTH, IN_WIDTH)
        const int g = std::gcd(d.out_w, d.in_w);
        d.scale_n = d.out_w / g;
        d.scale_d = d.in_w / g;
    } else {
        if (d.scale_n < 1 || d.scale_d < 1) return LANCZOS_ERR_RATIO;
        const int g = std::gcd(d.scale_n, d.scale_d);
        d.scale_n /= g;
        d.scale_d /= g;
    }
    if (d.scale_n < d.scale_d) return LANCZOS_ERR_RATIO;  // upscale only (worker.cpp:140)
    if (d.scale_n > (1 << 16)) return LANCZOS_ERR_RATIO;
    // the reference keeps the horizontal result in the first in_h rows of the output plane
    if (d.out_h < d.in_h) return LANCZOS_ERR_DIMS;
    const int64_t in_row = (int64_t)d.in_w * d.channels, out_row = (int64_t)d.out_w * d.channels;
    if (d.in_pitch == 0) d.in_pitch = in_row;
    if (d.out_pitch == 0) d.out_pitch = out_row;
    if (d.in_pitch < in_row || d.out_pitch < out_row) return LANCZOS_ERR_DIMS;
    *out = d;
    return LANCZOS_OK;
}

static int build_axis(AxisTables &t, int out_len, int in_len, int a, int n, int dd,
                      const std::vector<float> &phase_w, const std::vector<double> &phase_wd) {
    const int taps = 2 * a;
    const double scale = (double)n / dd;  // lanczos.h:112
    t.out_len = out_len;
    t.in_len = in_len;
    t.i0.resize(out_len);
    t.wd.assign((size_t)out_len * taps, 0.0);
    t.wf.assign((size_t)out_len * taps, 0.f);
    t.aligned_exact = true;
    t.uniform_phase = true;
    t.fast_err = 0;
    const double u = std::ldexp(1.0, -24);  // relative half-ulp of fp32
    for (int xx = 0; xx < out_len; xx++) {
        const double x = (double)xx / scale;  // full_TB.h:57,70
        const double fx = std::floor(x);
        const int64_t q = (int64_t)xx * dd / n;
        if ((double)q != fx) return LANCZOS_ERR_RATIO_FLOAT;
        const int phase = (int)(((int64_t)xx * dd) % n);
        if (phase == 0 && x != fx) t.aligned_exact = false;
        const int first = (int)q - a + 1;
        t.i0[xx] = first;
        double werr_coord = 0, werr_phase = 0, abs_sum = 0, round_err = 0;
        for (int k = 0; k < taps; k++) {
            const int i = first + k;
            const double w = ref_kernel(x - i, a);  // full_TB.h:60
            const float wf = (float)w;
            t.wd[(size_t)xx * taps + k] = w;
            if (w != phase_wd[(size_t)phase * taps + k]) t.uniform_phase = false;
            t.wf[(size_t)xx * taps + k] = wf;
            werr_coord += 255.0 * std::fabs((double)wf - w);
            werr_phase += 255.0 * std::fabs((double)phase_w[(size_t)phase * taps + k] - w);
            // taps are accumulated in ascending order (both kernels): the k-th FMA rounds a partial sum
            // bounded by the prefix of absolute products (+1 covers the -guard start value)
            abs_sum += 255.0 * std::fabs(w);
            round_err += u * (abs_sum + 1.0);
        }
        t.fast_err = std::max(t.fast_err, std::max(werr_coord, werr_phase) + round_err);
        // Tighter bound for the second-generation kernels (phase-table weights, FMA chains from -guard):
        //  * bytes are >= 0, so a partial sum lies in [255 * sum of negative weights, 255 * sum of positive ones];
        //  * one FMA rounds by at most half an ulp of its result: 2^(floor(log2 |result|) - 24);
        //  * the last FMA only matters when the sum is below 256 (anything above is clamped either way);
        //  * the weight error sum_k b_k (wf_k - w_k) is bounded by 255 * max(sum of positive, sum of negative errors).
        for (int order = 0; order < 2; order++) {
            double lo = 0, hi = 0, rerr = 0, dpos = 0, dneg = 0;
            for (int i = 0; i < taps; i++) {
                const int k = order == 0 ? i : ((i & 1) ? taps - 1 - i / 2 : i / 2);
                const double wf = (double)phase_w[(size_t)phase * taps + k];
                if (wf > 0) hi += 255.0 * wf; else lo += 255.0 * wf;
                double m = std::max(-lo, hi) + std::ldexp(1.0, -10);     // + the -guard start value
                if (i == taps - 1) m = std::min(m, 255.999);
                rerr += std::ldexp(1.0, (int)std::floor(std::log2(m)) - 24);
                const double dw = wf - t.wd[(size_t)xx * taps + k];
                if (dw > 0) dpos += dw; else dneg -= dw;
            }
            const double e = rerr + 255.0 * std::max(dpos, dneg);
            if (order == 0) t.err_asc = std::max(t.err_asc, e); else t.err_outer = std::max(t.err_outer, e);
        }
    }
    return LANCZOS_OK;
}

int build_plan(const lanczos_desc *desc, Plan *out) {
    Plan p;
    int rc = resolve_desc(desc, &p.d);
    if (rc != LANCZOS_OK) return rc;
    const int a = p.d.a, n = p.d.scale_n, dd = p.d.scale_d;
    p.taps = 2 * a;
    // polyphase table: phase ph of the period-N pattern is first reached at output xx with
    // (xx*D) mod N == ph; use that coordinate's weights (all coordinates of a phase agree to ~1e-13)
    p.phase_w.assign((size_t)n * p.taps, 0.f);
    p.phase_wd.assign((size_t)n * p.taps, 0.0);
    {
        const double scale = (double)n / dd;
        std::vector<char> seen(n, 0);
        for (int xx = 0; xx < n; xx++) {
            const int ph = (int)(((int64_t)xx * dd) % n);
            if (seen[ph]) continue;
            seen[ph] = 1;
            const double x = (double)xx / scale;
            const int first = (int)((int64_t)xx * dd / n) - a + 1;
            for (int k = 0; k < p.taps; k++) {
                const double w = ref_kernel(x - (first + k), a);
                p.phase_wd[(size_t)ph * p.taps + k] = w;
                p.phase_w[(size_t)ph * p.taps + k] = (float)w;
            }
        }
    }
    // Phase 0 (coordinate exactly on an input sample): the reference's weights are 1 at the centre
    // tap and ~1e-17 sin(k*pi) residues elsewhere (full_TB.h:43 has no |x|<a window).  Its double
    // sum can only fall below v (and truncate to v-1) if the negative residues outweigh half the
    // spacing of doubles below v, which is >= v*2^-54.  Sufficient for "output == v":
    //     sum_k K_k * b_k <= v,   K_k = |w_k| * 2^54 / 0.99 for residues w_k < 0 (else 0).
    p.align_k.assign(p.taps, 0.f);
    for (int k = 0; k < p.taps; k++) {
        const double w = p.phase_wd[k];
        if (k != a - 1 && w < 0) {
            const double K = -w * std::ldexp(1.0, 54) / 0.99;
            if (K > 1e-7) p.align_k[k] = (float)(K * (1.0 + 1e-6));
        }
    }
    // Sharper phase-0 test of the slow paths (lanczos_v6.cu phase0_doubt2), a = 3: with H = 2^ceil(log2 v) and
    // K_k = |w_k| * 2^54 the reference returns v if  K0 b0 - K1 b1 <= H  and  (K4 b4 <= H  or  K3 b3 - K4 b4 >= 2H).
    // Negative residues get the 1/0.99 margin (rounding of the reference's own products) and are rounded away
    // from zero; positive ones get 0.99 * (1 - 2^-10) (the second factor covers the rounding of the first of
    // two chained fp16 FMAs) and are rounded toward zero.  A residue with the wrong sign disables its term.
    if (a == 3) {
        const double *w0 = p.phase_wd.data();       // phase 0
        const double S = std::ldexp(1.0, 54 + 12);  // the kernel's fp16 values carry 2^-12
        auto neg_k = [&](double w) { return half2_bits(w * S / 0.99, true); };
        auto pos_k = [&](double w) { return half2_bits(w * S * 0.99 * (1.0 - 1.0 / 1024.0), false); };
        // the argument holds for this sign pattern only (taps 0 and 4 pull down, 1 and 3 lift, tap 5 is far
        // below half an ulp of 1); anything else: every sample stays in doubt (-inf constants)
        const bool pattern = w0[0] < 0 && w0[4] < 0 && w0[1] > 0 && w0[3] > 0 && w0[2] == 1.0 && std::fabs(w0[5]) < 1e-25;
        p.p0_half2[0] = pattern ? neg_k(w0[0]) : 0xFC00FC00u;
        p.p0_half2[1] = pattern ? pos_k(w0[1]) : 0u;
        p.p0_half2[2] = pattern ? pos_k(w0[3]) : 0u;
        p.p0_half2[3] = pattern ? neg_k(w0[4]) : 0xFC00FC00u;
    }
    rc = build_axis(p.x, p.d.out_w, p.d.in_w, a, n, dd, p.phase_w, p.phase_wd);
    if (rc != LANCZOS_OK) return rc;
    rc = build_axis(p.y, p.d.out_h, p.d.in_h, a, n, dd, p.phase_w, p.phase_wd);
    if (rc != LANCZOS_OK) return rc;
    // guard band: 1.25x the rigorous fp32 error bound, never below 2^-14
    const double e = std::max(p.x.fast_err, p.y.fast_err);
    p.guard = (float)std::max(1.25 * e, std::ldexp(1.0, -14));
    p.guard_asc = (float)(1.0625 * std::max(p.x.err_asc, p.y.err_asc));
    p.guard_outer = (float)(1.0625 * std::max(p.x.err_outer, p.y.err_outer));

    // in-place aliasing of the vertical pass (full_TB.h:67-77): going bottom-up, row xx reads
    // rows first..last; any row i > xx has already been overwritten with final output.
    p.alias_rows = 0;
    p.alias_top_row = -1;
    p.alias_in_rows = 0;
    if (!(p.d.flags & LANCZOS_FLAG_NO_ALIAS)) {
        auto last_row = [&](int xx) { return std::min(p.d.in_h - 1, p.y.i0[xx] + p.taps - 1); };
        for (int xx = 0; xx < p.d.out_h; xx++)
            if (last_row(xx) > xx) p.alias_rows = xx + 1;
        for (int xx = 0; xx < p.alias_rows; xx++) p.alias_top_row = std::max(p.alias_top_row, last_row(xx));
        for (int xx = 0; xx <= p.alias_top_row; xx++) p.alias_in_rows = std::max(p.alias_in_rows, last_row(xx) + 1);
    }
    *out = std::move(p);
    return LANCZOS_OK;
}

void band_rows(const Plan &p, int out_row0, int out_rows, int *in_row0, int *in_rows) {
    int lo = std::min(p.d.in_h - 1, std::max(0, p.y.i0[out_row0]));
    int hi = std::min(p.d.in_h - 1, p.y.i0[out_row0 + out_rows - 1] + p.taps - 1);
    if (out_row0 < p.alias_rows) {  // the in-place emulation starts from row alias_top_row
        lo = 0;
        hi = std::max(hi, p.alias_in_rows - 1);
    }
    if (hi < lo) hi = lo;  // band entirely below the last input row: nothing is read
    *in_row0 = lo;
    *in_rows = hi - lo + 1;
}

}  // namespace lzb
