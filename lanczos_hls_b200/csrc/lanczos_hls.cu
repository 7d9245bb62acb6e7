// lanczos_hls.cu -- the reference's fixed-point HLS arithmetic ("HLS mode") on sm_100a.
//
// PARITY UNPINNED against the reference itself (its LUT comes from Xilinx hls::sinpi, which is not
// available; see oracle/hls_oracle.c).  Bit-exact against the integer restatement in oracle/.
//
// Replaces (reference LanczosUpscaler/): kernel.cpp:40-67 (LUT indexed by |out*D - in*N|),
// worker.cpp:45-115 compute/compute_ (integer MAC, per-tap floor in the second pass, de-ring clamp
// to the two central taps), worker.cpp:118-130 clamp_to_byte, the zero / replicate borders of
// worker.cpp:170-198, :239-275 and cyclic_buffer.h:30-42, and the vertical-then-horizontal order
// of lanczos.cpp:68-98.  Integer scales only (SCALE_D = 1).
#include "kernels.cuh"

namespace lzb {
namespace {

constexpr int HT_W = 64, HT_H = 32, HT_THREADS = 256;
constexpr int HT_COLS = HT_W + 10;  // input columns a tile can touch (n >= 1, a <= 4)

struct HlsParams {
    const uint8_t *in; uint8_t *out;
    long long in_pitch, out_pitch, in_frame_stride, out_frame_stride;
    int in_w, in_h, out_w, out_h, a, n, bp;
    int lut[128];
};

template <int C>
__global__ void __launch_bounds__(HT_THREADS) lanczos_hls_kernel(const __grid_constant__ HlsParams p) {
    __shared__ int s_mid[HT_H][HT_COLS * C];   // vertical results, bp fraction bits
    const int tid = threadIdx.x, taps = 2 * p.a, n = p.n, a = p.a, bp = p.bp;
    const int x0 = blockIdx.x * HT_W, y0 = blockIdx.y * HT_H;
    const int tw = min(HT_W, p.out_w - x0), th = min(HT_H, p.out_h - y0);
    const uint8_t *in = p.in + (long long)blockIdx.z * p.in_frame_stride;
    uint8_t *out = p.out + (long long)blockIdx.z * p.out_frame_stride;
    const int cb0 = x0 / n - a + 1;                         // first (nominal) input column of the tile
    const int ncols = (x0 + tw - 1) / n + a - cb0 + 1;

    // vertical pass (ColWorkers, worker.cpp:138-155): zero rows above, last row replicated below
    for (int idx = tid; idx < th * ncols * C; idx += HT_THREADS) {
        const int ly = idx / (ncols * C), r = idx - ly * (ncols * C);
        const int lc = r / C, c = r - lc * C;
        const int col = cb0 + lc;
        int res = 0;
        if (col >= 0) {                                     // left of the image: zeros (worker.cpp:256-265)
            const int cc = min(col, p.in_w - 1);            // right of it: replicate (worker.cpp:244)
            const int y = y0 + ly, base = y / n;
            int acc = 0, c0 = 0, c1 = 0;
            for (int j = 0; j < taps; j++) {
                const int row = base - a + 1 + j;
                int v = 0;
                if (row >= 0) v = in[(long long)min(row, p.in_h - 1) * p.in_pitch + (long long)cc * C + c];
                acc += p.lut[abs(y - row * n)] * v;         // exact: bp fraction bits (worker.cpp:58)
                if (j == a - 1) c0 = v << bp;
                if (j == a) c1 = v << bp;
            }
            res = max(min(c0, c1), min(acc, max(c0, c1)));  // de-ring clamp (worker.cpp:66-74)
        }
        s_mid[ly][lc * C + c] = res;
    }
    __syncthreads();

    // horizontal pass (RowWorkers, worker.cpp:225-236) + clamp_to_byte (worker.cpp:118-130)
    for (int idx = tid; idx < th * tw * C; idx += HT_THREADS) {
        const int ly = idx / (tw * C), r = idx - ly * (tw * C);
        const int lx = r / C, c = r - lx * C;
        const int x = x0 + lx, base = x / n;
        int acc = 0, c0 = 0, c1 = 0;
        for (int j = 0; j < taps; j++) {
            const int col = base - a + 1 + j;
            const int v = s_mid[ly][(col - cb0) * C + c];
            acc += (int)(((long long)p.lut[abs(x - col * n)] * v) >> bp);   // per-tap floor (worker.cpp:95)
            if (j == a - 1) c0 = v;
            if (j == a) c1 = v;
        }
        acc = max(min(c0, c1), min(acc, max(c0, c1)));
        out[(long long)(y0 + ly) * p.out_pitch + (long long)x * C + c] = (uint8_t)(acc >> bp);
    }
}

}  // namespace

int launch_hls(const uint8_t *in, uint8_t *out, long long in_pitch, long long out_pitch, long long in_fs,
               long long out_fs, int n_frames, int in_w, int in_h, int out_w, int out_h, int channels, int a, int n,
               int bp, const int *lut, cudaStream_t s) {
    HlsParams p{};
    p.in = in; p.out = out; p.in_pitch = in_pitch; p.out_pitch = out_pitch;
    p.in_frame_stride = in_fs; p.out_frame_stride = out_fs;
    p.in_w = in_w; p.in_h = in_h; p.out_w = out_w; p.out_h = out_h; p.a = a; p.n = n; p.bp = bp;
    for (int i = 0; i <= a * n; i++) p.lut[i] = lut[i];
    dim3 grid((out_w + HT_W - 1) / HT_W, (out_h + HT_H - 1) / HT_H, n_frames);
    switch (channels) {
        case 1: lanczos_hls_kernel<1><<<grid, HT_THREADS, 0, s>>>(p); break;
        case 2: lanczos_hls_kernel<2><<<grid, HT_THREADS, 0, s>>>(p); break;
        case 3: lanczos_hls_kernel<3><<<grid, HT_THREADS, 0, s>>>(p); break;
        default: lanczos_hls_kernel<4><<<grid, HT_THREADS, 0, s>>>(p); break;
    }
    return (int)cudaGetLastError();
}

}  // namespace lzb
