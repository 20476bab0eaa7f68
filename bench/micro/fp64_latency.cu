// Dependent-issue latency of the FP64 instructions the one-CTA solve kernel is made of (B200, sm_100a):
// one warp, a chain of N dependent operations, clock64() around it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 4096;

template <int OP>
__global__ void k_chain(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[1024];
    sm[threadIdx.x] = 0.0;
    __syncthreads();
    double x = a + threadIdx.x, y = b;
    float xf = (float)a + threadIdx.x, yf = (float)b;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = fma(x, y, y);                       // DFMA
        if (OP == 1) x = x + y;                               // DADD
        if (OP == 2) x = x * y;                               // DMUL
        if (OP == 3) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x));      // MUFU.RCP64H
        if (OP == 4) x = __shfl_xor_sync(0xffffffffu, x, 1);  // 64-bit shuffle (2 SHFL)
        if (OP == 5) xf = fmaf(xf, yf, yf);                   // FFMA
        if (OP == 6) x = rsqrt(x) + y;                        // rsqrt sequence
        if (OP == 7) x = 1.0 / x + y;                         // IEEE division sequence
        if (OP == 8) x = sqrt(x) + y;
        if (OP == 9) { __syncthreads(); x += y; }               // block barrier (+ one DADD)
        if (OP == 10) x = (double)rsqrtf((float)x) + y;      // fp32 rsqrt seed with conversions
        if (OP == 11) { sm[threadIdx.x] = x; __syncthreads(); x = sm[(threadIdx.x + 33) % blockDim.x] + y; __syncthreads(); }   // STS, barrier, LDS from another warp, DADD, barrier
        if (OP == 12) { sm[threadIdx.x] = x; __syncwarp(); x = sm[threadIdx.x ^ 1] + y; __syncwarp(); }          // STS, LDS within the warp, DADD
        if (OP == 13) { x = sm[((int)threadIdx.x + i) & 1023] + x; }                                                  // dependent LDS + DADD (address independent)
        if (OP == 14) { const int a = __double2loint(x) & 1023; x = sm[a] + y; }                                       // pointer-chase LDS (address depends on the value)
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x + xf;
}

template <int OP>
void run(const char* name, int threads) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    k_chain<OP><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999);
    k_chain<OP><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %4d threads/CTA: %7.1f cycles per dependent op\n", name, threads, (double)c / N);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int th : {32, 256, 1024}) {
        run<0>("DFMA", th); run<1>("DADD", th); run<2>("DMUL", th); run<3>("MUFU.RCP64H", th); run<4>("SHFL 64-bit", th);
        run<5>("FFMA", th); run<6>("rsqrt(x)+y", th); run<7>("1/x+y (IEEE)", th); run<8>("sqrt(x)+y", th); run<9>("__syncthreads + DADD", th); run<10>("cvt+rsqrtf+cvt+DADD", th); run<11>("STS+BAR+LDS(other warp)+DADD+BAR", th); run<12>("STS+LDS(same warp)+DADD", th); run<13>("LDS+DADD", th); run<14>("LDS pointer chase+DADD", th);
    }
    return 0;
}
