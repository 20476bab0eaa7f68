// FP64 pipe micro-benchmark for B200 (sm_100a): what is the real ceiling for the
// normal-equation assembly kernel?  Measures, with CUDA events,
//   (1) DFMA issue rate (vector FP64 pipe),
//   (2) DMMA rate for every f64 mma.sync shape (m8n8k4, m16n8k4, m16n8k8, m16n8k16),
//   (3) DMMA + DMUL interleaved (do they share one pipe?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dmma884(double* out, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int KK, int NMUL>
__global__ void k_dmma16(double* out, double a0, double b0) {
    double c[NACC][4];
    double a[8], b[4];
    double m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = a0 + i; m[i] = 1.0 + 1e-9 * (threadIdx.x + i); }
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = b0 + i;
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (KK == 4) dmma1684(c[i], a, b[0]);
            if (KK == 8) dmma1688(c[i], a, b);
            if (KK == 16) dmma16816(c[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < NMUL; i++) m[i % 8] = m[i % 8] * a0;   // independent DMUL chains
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
#pragma unroll
    for (int i = 0; i < 8; i++) s += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    // a configuration the device cannot launch (too many registers for the block size) must not be reported as a rate
    // (round 1 printed 239328 TFLOP/s for the 32-warp x 4-CTA rows: the launch had failed and the events timed nothing)
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) { cudaGetLastError(); return -1.0f; }
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    if (cudaGetLastError() != cudaSuccess) return -1.0f;
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
    for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM (one CTA of wps warps per SM x 4 CTAs)
        int threads = wps * 32;
        int blocks = sms * 4;
        double nthreads = (double)threads * blocks;
        float ms = time_it([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("warps/CTA %2d x4 CTA/SM  DFMA           %8.2f TFLOP/s\n", wps, 2.0 * 16 * ITERS * nthreads / ms / 1e9);
        double nwarps = nthreads / 32;
        ms = time_it([&] { k_dmma884<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("warps/CTA %2d x4 CTA/SM  DMMA m8n8k4    %8.2f TFLOP/s\n", wps, 2.0 * 8 * 8 * 4 * 8 * ITERS * nwarps / ms / 1e9);
        ms = time_it([&] { k_dmma16<8, 4, 0><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("warps/CTA %2d x4 CTA/SM  DMMA m16n8k4   %8.2f TFLOP/s\n", wps, 2.0 * 16 * 8 * 4 * 8 * ITERS * nwarps / ms / 1e9);
        ms = time_it([&] { k_dmma16<8, 8, 0><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("warps/CTA %2d x4 CTA/SM  DMMA m16n8k8   %8.2f TFLOP/s\n", wps, 2.0 * 16 * 8 * 8 * 8 * ITERS * nwarps / ms / 1e9);
        ms = time_it([&] { k_dmma16<8, 16, 0><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("warps/CTA %2d x4 CTA/SM  DMMA m16n8k16  %8.2f TFLOP/s\n", wps, 2.0 * 16 * 8 * 16 * 8 * ITERS * nwarps / ms / 1e9);
        ms = time_it([&] { k_dmma16<8, 4, 8><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("warps/CTA %2d x4 CTA/SM  DMMA m16n8k4 + 8 DMUL per 8 DMMA: DMMA-only rate %8.2f TFLOP/s (time %.3f ms)\n", wps,
               2.0 * 16 * 8 * 4 * 8 * ITERS * nwarps / ms / 1e9, ms);
        ms = time_it([&] { k_dmma16<8, 4, 32><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("warps/CTA %2d x4 CTA/SM  DMMA m16n8k4 + 32 DMUL per 8 DMMA: DMMA-only rate %8.2f TFLOP/s (time %.3f ms)\n", wps,
               2.0 * 16 * 8 * 4 * 8 * ITERS * nwarps / ms / 1e9, ms);
    }
    // single-accumulator latency (dependent chain), 1 warp per SMSP
    {
        float ms = time_it([&] { k_dmma16<1, 4, 0><<<sms, 128>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("dependent DMMA m16n8k4 chain: %.1f ns per mma (~%.0f cycles at 1.9 GHz)\n", ms * 1e6 / ITERS, ms * 1e6 / ITERS * 1.9);
        ms = time_it([&] { k_dmma16<1, 8, 0><<<sms, 128>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("dependent DMMA m16n8k8 chain: %.1f ns per mma\n", ms * 1e6 / ITERS);
        ms = time_it([&] { k_dmma884<1><<<sms, 128>>>(out, 1.0000001, 1e-9); });
        if (ms < 0.0f) { printf("(launch failed: configuration not launchable on this device)\n"); }
        if (ms > 0.0f) printf("dependent DMMA m8n8k4 chain: %.1f ns per mma\n", ms * 1e6 / ITERS);
    }
    cudaFree(out);
    return 0;
}
