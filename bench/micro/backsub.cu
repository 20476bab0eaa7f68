// Why does the 64-step back substitution of k_micro_solve_fast take ~650 cycles per step?  Isolated loop:
// 2 working warps, the other warps of the CTA waiting at the block barrier (V=0) or not present (V=1).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o backsub backsub.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NPAD = 64, LD = NPAD + 3;

template <int V>
__global__ void k(double* out, long long* cyc, int np) {
    extern __shared__ double Lr[];
    const int tid = threadIdx.x;
    for (int e = tid; e < (NPAD + 3) * LD; e += blockDim.x) Lr[e] = 1e-3 * ((e * 7) % 13);
    __syncthreads();
    const long long t0 = clock64();
    double v0 = 0, v1 = 0;
    if (tid < 64) {
        const int lane = tid & 31;
        const double* zrow = Lr + (np + (tid >> 5)) * LD;
        double acc0 = lane < np ? zrow[lane] : 0.0, acc1 = lane + 32 < np ? zrow[lane + 32] : 0.0;
        double r0, r1;
        {
            const double* row = Lr + (np - 1) * LD;
            r0 = lane < np - 1 ? row[lane] : 0.0;
            r1 = lane + 32 < np - 1 ? row[lane + 32] : 0.0;
        }
        for (int kk = np - 1; kk >= 0; kk--) {
            const double c0 = r0, c1 = r1;
            if (kk > 0) {
                const double* row = Lr + (kk - 1) * LD;
                r0 = lane < kk - 1 ? row[lane] : 0.0;
                r1 = lane + 32 < kk - 1 ? row[lane + 32] : 0.0;
            }
            const double xk = __shfl_sync(0xffffffffu, (kk & 32) ? acc1 : acc0, kk & 31);
            acc0 = fma(-c0, xk, acc0);
            acc1 = fma(-c1, xk, acc1);
        }
        v0 = acc0; v1 = acc1;
    }
    if (V == 0) __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cyc[0] = t1 - t0;
    out[tid] = v0 + v1;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    const size_t sm = (NPAD + 3) * LD * 8;
    for (int th : {64, 256, 1024}) {
        long long c;
        k<0><<<1, th, sm>>>(out, cyc, 64); k<0><<<1, th, sm>>>(out, cyc, 64);
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %4d, barrier after: %lld cycles (%.1f per step)\n", th, c, c / 64.0);
        k<1><<<1, th, sm>>>(out, cyc, 64); k<1><<<1, th, sm>>>(out, cyc, 64);
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %4d, no barrier   : %lld cycles (%.1f per step)\n", th, c, c / 64.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
