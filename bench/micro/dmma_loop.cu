// What limits the assembly kernel's DMMA rate?  Reproduces its inner loop shape (11 row tiles x 2 column tiles,
// A fragment = product of two shared-memory operands) in isolation, variant by variant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_loop dmma_loop.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;   // "steps"
constexpr int RT = 10, CT = 2;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// V: 0 fixed register operands; 1 A operand from shared memory (1 LDS per row tile); 2 A = product of two LDS;
//    3 as 2 plus B operands from shared memory per step; 4 as 3 with software pipelining (loads 2 tiles ahead,
//    product 1 tile ahead)
template <int V>
__global__ void __launch_bounds__(128) k_loop(double* out, double a0, double b0) {
    __shared__ double sm[4][40 * 36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ws = sm[warp];
    for (int i = lane; i < 40 * 36; i += 32) ws[i] = 1.0 + 1e-9 * i;
    __syncwarp();
    int off1[RT], off2[RT];
#pragma unroll
    for (int r = 0; r < RT; r++) { off1[r] = ((r * 3 + (lane >> 2)) % 17) * 36; off2[r] = (17 + (r + (lane >> 2)) % 10) * 36; }
    const int colo[2] = {(28 + (lane >> 2)) * 36, (30 + (lane >> 2) % 4) * 36};
    double c[RT * CT][2];
#pragma unroll
    for (int i = 0; i < RT * CT; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double bc[CT] = {b0, b0 + 1};
    const double* w0 = ws + (lane & 3);
    for (int it = 0; it < ITERS; it++) {
        const double* w = w0 + (it & 7) * 4;
        if (V >= 3) { bc[0] = w[colo[0]]; bc[1] = w[colo[1]]; }
        if (V <= 3) {
#pragma unroll
            for (int r = 0; r < RT; r++) {
                double av;
                if (V == 0) av = a0;
                else if (V == 1) av = w[off1[r]];
                else av = w[off1[r]] * w[off2[r]];
#pragma unroll
                for (int q = 0; q < CT; q++) dmma884(c[r * CT + q][0], c[r * CT + q][1], av, bc[q]);
            }
        } else {
            double av = w[off1[0]] * w[off2[0]];
            double n0a = w[off1[1]], n0b = w[off2[1]], n1a = w[off1[2]], n1b = w[off2[2]];
#pragma unroll
            for (int r = 0; r < RT; r++) {
                const double cur = av;
                av = n0a * n0b; n0a = n1a; n0b = n1b;
                const int nr = (r + 3) % RT;
                n1a = w[off1[nr]]; n1b = w[off2[nr]];
#pragma unroll
                for (int q = 0; q < CT; q++) dmma884(c[r * CT + q][0], c[r * CT + q][1], cur, bc[q]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < RT * CT; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(double* out, int sms, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * ctas_per_sm;
    k_loop<V><<<blocks, 128>>>(out, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k_loop<V><<<blocks, 128>>>(out, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * 256 * RT * CT * (double)ITERS * blocks * 4;
    printf("variant %d, %d CTAs(4 warps)/SM: %7.2f TFLOP/s DMMA  (%.3f ms)  err=%s\n", V, ctas_per_sm, flops / best / 1e9, best,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 128);
    for (int c = 1; c <= 4; c++) {
        run<0>(out, sms, c); run<1>(out, sms, c); run<2>(out, sms, c); run<3>(out, sms, c); run<4>(out, sms, c);
    }
    return 0;
}
