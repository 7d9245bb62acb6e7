// kbench.cu -- C++ development harness around the C ABI (no Python, starts in a second on a fresh box).
//   tools/bin/kbench IN_W IN_H N D A C FRAMES CONTENT ITERS [FLAGS] [CMP_IMPL]
// CONTENT: smooth | noise | dark | mix (even frames smooth, odd frames noise).   Runs lanczos_b200_upscale_batch ITERS times (CUDA events), prints
// Gpix/s, GB/s of algorithmic traffic and an FNV-1a hash of the output.  If CMP_IMPL is given (any word, e.g.
// "generic") the same call is repeated with LANCZOS_FLAG_GENERIC_KERNEL (an independent implementation inside
// the library: per-coordinate weights, one thread per output pixel) and the two outputs are compared byte for
// byte on the device; with flags & 8 (tolerance mode) the comparison reports the exact-match fraction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/kbench tools/kbench.cu \
//        -Llanczos_hls_b200 -llanczos_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../lanczos_hls_b200'
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../include/lanczos_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// content 0: smooth + noise (SURVEY 8d ii), 1: uniform noise, 2: dark noise 0..15
__global__ void fill_kernel(uint8_t *img, long long n, int w, int h, int c, int content) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        const long long px = i / c;
        const int x = (int)(px % w), y = (int)((px / w) % h), f = (int)(px / ((long long)w * h));
        const uint32_t r = mix((uint32_t)i * 2654435761u + (uint32_t)(i >> 32) + 12345u);
        int v;
        if (content == 1 || (content == 3 && (f & 1))) v = r & 255;      // content 3: even frames smooth, odd frames noise
        else if (content == 2) v = r & 15;
        else {
            const float b = 128.f + 90.f * sinf(0.05f * x + ch + 0.3f * f) * cosf(0.037f * y);
            v = (int)(b + (float)((int)(r & 15) - 8));
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
        }
        img[i] = (uint8_t)v;
    }
}
__global__ void cmp_kernel(const uint8_t *a, const uint8_t *b, long long n, unsigned long long *ndiff, unsigned long long *first, int *maxd) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int d = abs((int)a[i] - (int)b[i]);
        if (d) {
            atomicAdd(ndiff, 1ull);
            atomicMin(first, (unsigned long long)i);
            atomicMax(maxd, d);
        }
    }
}

static uint64_t fnv(const uint8_t *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv) {
    if (argc < 10) { printf("usage: kbench IN_W IN_H N D A C FRAMES CONTENT ITERS [FLAGS] [CMP_IMPL]\n"); return 2; }
    const int iw = atoi(argv[1]), ih = atoi(argv[2]), n = atoi(argv[3]), d = atoi(argv[4]), a = atoi(argv[5]), c = atoi(argv[6]);
    const int frames = atoi(argv[7]);
    const std::string content = argv[8];
    const int iters = atoi(argv[9]);
    const unsigned flags = argc > 10 ? (unsigned)strtoul(argv[10], nullptr, 0) : 0u;
    const char *cmp_impl = argc > 11 ? argv[11] : nullptr;
    const int ow = (int)((long long)iw * n / d), oh = (int)((long long)ih * n / d);
    const int ct = content == "noise" ? 1 : (content == "dark" ? 2 : (content == "mix" ? 3 : 0));

    lanczos_desc desc{};
    desc.in_w = iw; desc.in_h = ih; desc.out_w = ow; desc.out_h = oh; desc.channels = c; desc.a = a;
    desc.scale_n = n; desc.scale_d = d; desc.flags = flags;
    const size_t in_frame = (size_t)iw * ih * c, out_frame = (size_t)ow * oh * c;
    uint8_t *d_in, *d_out, *d_out2 = nullptr;
    CK(cudaMalloc(&d_in, in_frame * frames));
    CK(cudaMalloc(&d_out, out_frame * frames));
    fill_kernel<<<2048, 256>>>(d_in, (long long)in_frame * frames, iw, ih, c, ct);
    CK(cudaMemset(d_out, 0xAB, out_frame * frames));
    CK(cudaDeviceSynchronize());

    lanczos_b200_enable_stats(1);
    int rc = lanczos_b200_upscale_batch(&desc, d_in, d_out, frames, 0, 0, 0, nullptr);
    if (rc != 0) { printf("upscale failed: %d %s (%s)\n", rc, lanczos_b200_strerror(rc), lanczos_b200_last_cuda_error()); return 1; }
    CK(cudaDeviceSynchronize());
    lanczos_stats st{};
    lanczos_b200_get_stats(&st);
    lanczos_b200_enable_stats(0);

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) lanczos_b200_upscale_batch(&desc, d_in, d_out, frames, 0, 0, 0, nullptr);
    CK(cudaDeviceSynchronize());
    float best = 1e30f, total = 0;
    for (int i = 0; i < iters; i++) {
        CK(cudaEventRecord(e0));
        rc = lanczos_b200_upscale_batch(&desc, d_in, d_out, frames, 0, 0, 0, nullptr);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best; total += ms;
    }
    // back-to-back launches (no host synchronisation in between): per-call time = max(host issue time, GPU time)
    {
        CK(cudaDeviceSynchronize());
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; i++) lanczos_b200_upscale_batch(&desc, d_in, d_out, frames, 0, 0, 0, nullptr);
        CK(cudaEventRecord(e1));
        clock_gettime(CLOCK_MONOTONIC, &t1);
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double host_us = ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 1e3 / iters;
        printf("   back to back: %.2f us per call on the GPU timeline, %.2f us of host time per call\n", ms * 1e3 / iters, host_us);
    }
    // the same frames as single-frame launches (one call per frame, buffers rotate through the whole batch so the
    // working set stays larger than L2), on 1 stream and round-robin over 4 streams
    if (frames >= 8) {
        cudaStream_t st[4];
        for (int i = 0; i < 4; i++) CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
        for (int ns = 1; ns <= 4; ns *= 2) {
            CK(cudaDeviceSynchronize());
            const int calls = iters * frames;
            CK(cudaEventRecord(e0, st[0]));
            for (int i = 1; i < ns; i++) CK(cudaStreamWaitEvent(st[i], e0, 0));
            for (int i = 0; i < calls; i++) {
                const int f = i % frames;
                lanczos_b200_upscale(&desc, d_in + in_frame * f, d_out + out_frame * f, 0, st[i % ns]);
            }
            cudaEvent_t ej[4];
            for (int i = 1; i < ns; i++) { CK(cudaEventCreate(&ej[i])); CK(cudaEventRecord(ej[i], st[i])); CK(cudaStreamWaitEvent(st[0], ej[i], 0)); }
            CK(cudaEventRecord(e1, st[0]));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("   single-frame launches over %d frames, %d stream(s): %.2f us per frame = %.1f Gpix/s\n", frames, ns, ms * 1e3 / calls,
                   (double)ow * oh / (ms * 1e-3 / calls) / 1e9);
        }
    }
    const double opx = (double)ow * oh * frames;
    const double bytes = (double)(in_frame + out_frame) * frames;
    const double avg = total / iters;
    std::vector<uint8_t> host(out_frame);
    CK(cudaMemcpy(host.data(), d_out + out_frame * (frames - 1), out_frame, cudaMemcpyDeviceToHost));
    const uint64_t hsh = fnv(host.data(), out_frame);
    printf("%dx%d %d/%d a=%d c=%d frames=%d %s flags=%u kernel_id=%d launches=%lld strict=%lld | avg %.4f ms best %.4f ms | %.1f Gpix/s  %.1f GB/s (%.3f of 6552.6) | fnv(last frame) %016llx\n",
           iw, ih, n, d, a, c, frames, content.c_str(), flags, st.kernel_id, (long long)st.kernel_launches, (long long)st.strict_samples,
           avg, best, opx / (avg * 1e-3) / 1e9, bytes / (avg * 1e-3) / 1e9, bytes / (avg * 1e-3) / 1e9 / 6552.6, (unsigned long long)hsh);

    if (cmp_impl) {
        CK(cudaMalloc(&d_out2, out_frame * frames));
        CK(cudaMemset(d_out2, 0xCD, out_frame * frames));
        lanczos_desc desc2 = desc;
        desc2.flags = (desc.flags & LANCZOS_FLAG_NO_ALIAS) | LANCZOS_FLAG_GENERIC_KERNEL;
        rc = lanczos_b200_upscale_batch(&desc2, d_in, d_out2, frames, 0, 0, 0, nullptr);
        if (rc != 0) { printf("cmp upscale failed: %d\n", rc); return 1; }
        CK(cudaDeviceSynchronize());
        lanczos_b200_get_stats(&st);
        unsigned long long *d_cnt; int *d_max;
        CK(cudaMalloc(&d_cnt, 16)); CK(cudaMalloc(&d_max, 4));
        const unsigned long long init[2] = {0ull, ~0ull};
        CK(cudaMemcpy(d_cnt, init, 16, cudaMemcpyHostToDevice)); CK(cudaMemset(d_max, 0, 4));
        cmp_kernel<<<2048, 256>>>(d_out, d_out2, (long long)out_frame * frames, d_cnt, d_cnt + 1, d_max);
        CK(cudaDeviceSynchronize());
        unsigned long long res[2]; int maxd;
        CK(cudaMemcpy(res, d_cnt, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&maxd, d_max, 4, cudaMemcpyDeviceToHost));
        float ms = 0;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 3; i++) lanczos_b200_upscale_batch(&desc2, d_in, d_out2, frames, 0, 0, 0, nullptr);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("   vs %s (kernel_id=%d, %.4f ms): %llu bytes differ of %.0f (exact fraction %.9f), max |diff| %d",
               cmp_impl, st.kernel_id, ms / 3, res[0], (double)out_frame * frames, 1.0 - (double)res[0] / ((double)out_frame * frames), maxd);
        if (res[0]) {
            const unsigned long long i = res[1];
            const unsigned long long f = i / out_frame, r = (i % out_frame) / ((size_t)ow * c), b = i % ((size_t)ow * c);
            printf("  first at frame %llu row %llu byte %llu", f, r, b);
        }
        printf("\n");
        return res[0] ? 3 : 0;
    }
    return 0;
}
