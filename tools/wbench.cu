// wbench.cu -- what can the STORE side of the upscaler reach on this part?  (round 2, after the ablation of
// tools/ablate.sh showed the V pass alone to be bound by its stores, not by its arithmetic.)
// Config 2 writes 4x the bytes it reads (24.9 MB out, 6.2 MB in per frame), as 240-byte pieces per warp and row.
// Cases, all writing the 64-frame output of config 2 (1.59 GB), times by CUDA events:
//   memset            cudaMemsetAsync (the driver's own write-only stream)
//   linear            grid-stride kernel, 16 B per thread, consecutive addresses (the pattern of a copy)
//   strips W/L        one warp per strip of W bytes x 720 rows of one frame, L bytes per lane and store (our kernel:
//                     W = 240, L = 8), rows written top to bottom, 16 one-warp CTAs per SM like lanczos_v6_kernel
//   strips + read     the same while reading the input volume (1/4 of the bytes) with 16-byte loads
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/wbench tools/wbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int OUT_W_B = 3840 * 3, OUT_H = 2160, FRAMES = 64;
constexpr long long OUT_FRAME = (long long)OUT_W_B * OUT_H;

__global__ void linear_kernel(uint4 *out, long long n16) {
    const uint4 v = make_uint4(blockIdx.x, threadIdx.x, 3, 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) out[i] = v;
}

// one warp per CTA; dynamic smem keeps 16 CTAs per SM
template <int L>
__global__ void __launch_bounds__(32) strips_kernel(uint8_t *out, const uint8_t *in, int strip_b, int rows_per_seg, int read, unsigned *sink) {
    extern __shared__ uint8_t sm[];
    const int lane = threadIdx.x, strip = blockIdx.x, seg = blockIdx.y, frame = blockIdx.z;
    const int lanes = strip_b / L;
    uint8_t *o = out + frame * OUT_FRAME + (long long)seg * rows_per_seg * OUT_W_B + (long long)strip * strip_b + lane * L;
    unsigned acc = 0;
    const uint4 *ip = reinterpret_cast<const uint4 *>(in + (frame * (OUT_FRAME / 4)) + ((long long)seg * gridDim.x + strip) * (rows_per_seg / 4) * (strip_b));
    for (int y = 0; y < rows_per_seg; y++) {
        if (read && (y & 3) == 0) {   // one input row piece per 4 output rows (1/4 of the bytes), 16 B per lane
            if (lane * 16 < strip_b) { const uint4 v = ip[(y / 4) * (strip_b / 16) + lane]; acc += v.x ^ v.y ^ v.z ^ v.w; }
        }
        if (lane < lanes) {
            if (L == 8) *reinterpret_cast<uint2 *>(o) = make_uint2(y + acc, lane);
            else if (L == 16) *reinterpret_cast<uint4 *>(o) = make_uint4(y + acc, lane, 1, 2);
            else *reinterpret_cast<uint32_t *>(o) = y + acc;
        }
        o += OUT_W_B;
    }
    if (acc == 0x12345u) *sink = acc + sm[0];
}

template <class F>
float time_ms(F f, int iters = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < iters; i++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    const long long out_bytes = OUT_FRAME * FRAMES, in_bytes = out_bytes / 4;
    uint8_t *out, *in; unsigned *sink;
    CK(cudaMalloc(&out, out_bytes)); CK(cudaMalloc(&in, in_bytes + (1 << 20))); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(in, 1, in_bytes));
    auto report = [&](const char *name, float ms, double bytes) { printf("%-34s %.4f ms  %.0f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9); };
    report("memset", time_ms([&] { cudaMemsetAsync(out, 7, out_bytes); }), (double)out_bytes);
    report("linear 16 B/thread", time_ms([&] { linear_kernel<<<148 * 16, 256>>>((uint4 *)out, out_bytes / 16); }), (double)out_bytes);
    const size_t smem = 13 * 1024;
    CK(cudaFuncSetAttribute(strips_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(strips_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(strips_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int read = 0; read < 2; read++) {
        for (int segs : {3, 6}) {
            char name[128];
            snprintf(name, sizeof name, "strips 240 B, 8 B/lane, %d segs%s", segs, read ? " + read" : "");
            report(name, time_ms([&] { strips_kernel<8><<<dim3(OUT_W_B / 240, segs, FRAMES), 32, smem>>>(out, in, 240, OUT_H / segs, read, sink); }), (double)out_bytes * (read ? 1.25 : 1.0));
        }
        char name[128];
        snprintf(name, sizeof name, "strips 256 B, 8 B/lane, 3 segs%s", read ? " + read" : "");
        report(name, time_ms([&] { strips_kernel<8><<<dim3(OUT_W_B / 256, 3, FRAMES), 32, smem>>>(out, in, 256, OUT_H / 3, read, sink); }), (double)out_bytes * (read ? 1.25 : 1.0));
        snprintf(name, sizeof name, "strips 480 B, 16 B/lane, 3 segs%s", read ? " + read" : "");
        report(name, time_ms([&] { strips_kernel<16><<<dim3(OUT_W_B / 480, 3, FRAMES), 32, smem>>>(out, in, 480, OUT_H / 3, read, sink); }), (double)out_bytes * (read ? 1.25 : 1.0));
        snprintf(name, sizeof name, "strips 512 B, 16 B/lane, 3 segs%s", read ? " + read" : "");
        report(name, time_ms([&] { strips_kernel<16><<<dim3(OUT_W_B / 512, 3, FRAMES), 32, smem>>>(out, in, 512, OUT_H / 3, read, sink); }), (double)out_bytes * (read ? 1.25 : 1.0));
        snprintf(name, sizeof name, "strips 128 B, 4 B/lane, 3 segs%s", read ? " + read" : "");
        report(name, time_ms([&] { strips_kernel<4><<<dim3(OUT_W_B / 128, 3, FRAMES), 32, smem>>>(out, in, 128, OUT_H / 3, read, sink); }), (double)out_bytes * (read ? 1.25 : 1.0));
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
