// Where do the cycles of k_assemble_r4's k-step go?  The k-step (8 fragment loads, 13 DMULs, 15 DMMA.8x8x4) is rebuilt
// here feature by feature and timed at 1..4 warps per SM sub-partition; printed: FP64-pipe cycles per k-step and sub-
// partition (240 = the 15 DMMAs alone).
//   mode 0  15 DMMAs, loop-invariant register operands
//   mode 1  + 13 DMULs whose results are NOT used by the DMMAs (independent chains)
//   mode 2  + 13 DMULs whose results ARE the DMMA operands (register inputs, one iteration-dependent factor)
//   mode 3  fragment loads from shared memory feeding the DMMAs directly (no DMULs)
//   mode 4  the real k-step: loads -> DMULs -> DMMAs (ptxas places the DMULs)
//   mode 5  as 4, the k-step loop not unrolled
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o kstep_mix kstep_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int kStride = 36, kSm = 10;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

#define DMMA15(u0, u1, u2, u3, v0, v1, v2, v3, a10, a11, a12, a13, a14, bf, fy) \
    dmma884(acc[0][0], acc[0][1], u0, v0); dmma884(acc[1][0], acc[1][1], u0, v1); dmma884(acc[2][0], acc[2][1], u0, v2); \
    dmma884(acc[3][0], acc[3][1], u0, v3); dmma884(acc[4][0], acc[4][1], u1, v1); dmma884(acc[5][0], acc[5][1], u1, v2); \
    dmma884(acc[6][0], acc[6][1], u1, v3); dmma884(acc[7][0], acc[7][1], u2, v2); dmma884(acc[8][0], acc[8][1], u2, v3); \
    dmma884(acc[9][0], acc[9][1], u3, v3); dmma884(acc[10][0], acc[10][1], a10, bf); dmma884(acc[11][0], acc[11][1], a11, bf); \
    dmma884(acc[12][0], acc[12][1], a12, v3); dmma884(acc[13][0], acc[13][1], a13, fy); dmma884(acc[14][0], acc[14][1], a14, fy);

template <int MODE>
__global__ void __launch_bounds__(128) k_mix(double* out, double a0, int zero) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ws = sm + warp * (32 * kStride + 32 * kSm);
    for (int i = lane; i < 32 * kStride + 32 * kSm; i += 32) ws[i] = 1.0 + 1e-9 * i;
    __syncwarp();
    const int g = lane >> 2;
    const double* fm = ws + (lane & 3);
    const int o_rr = g * kStride, o_bf = (8 + g) * kStride, o_llg = (16 + g) * kStride, o_rb = (24 + (g & 3)) * kStride;
    const int o_fy = (g < 4 ? 28 + g : 15) * kStride;
    const double2* smp = reinterpret_cast<const double2*>(ws + 32 * kStride + (lane & 3) * kSm);
    const bool hi = (g & 4) != 0, t12 = g < 4, t12l = (g & 2) != 0, t12r = (g & 1) != 0;
    double acc[15][2];
#pragma unroll
    for (int t = 0; t < 15; t++) { acc[t][0] = lane; acc[t][1] = t; }
    double r[15];
#pragma unroll
    for (int i = 0; i < 15; i++) r[i] = a0 + 1e-7 * (i + lane);
    double m[13];
#pragma unroll
    for (int i = 0; i < 13; i++) m[i] = 1.0 + 1e-9 * (i + lane);
    double pu[13];
#pragma unroll
    for (int i = 0; i < 13; i++) pu[i] = r[i];
    for (int it = 0; it < ITERS; it += 8) {
#pragma unroll(MODE == 5 ? 1 : 8)
        for (int step = 0; step < 8; step++) {
            if (MODE == 0) {
                DMMA15(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14]);
            } else if (MODE == 1) {
                DMMA15(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14]);
#pragma unroll
                for (int i = 0; i < 13; i++) m[i] *= a0;
            } else if (MODE == 2) {
                const double s = __hiloint2double(__double2hiint(a0), __double2loint(a0) ^ ((it + step) & zero));
                double u[13];
#pragma unroll
                for (int i = 0; i < 13; i++) u[i] = r[i] * s;
                DMMA15(u[0], u[1], u[2], u[3], u[4], u[5], u[6], u[7], u[8], u[9], u[10], u[11], u[12], r[13], r[14]);
            } else {
                const double* w = fm + step * 4;
                const double2* sp = smp + step * (4 * kSm / 2);
                const double rrg = w[o_rr], bf = w[o_bf], llg = w[o_llg], rb = w[o_rb], fy = w[o_fy];
                const double2 r89 = sp[0], s01 = sp[1], s23 = sp[2];
                const double c12 = t12 ? (t12l ? s23.y : s23.x) : 0.0, d12 = t12r ? r89.y : r89.x;
                const double e13 = hi ? s01.y : s01.x, e14 = hi ? s23.y : s23.x;
                if (MODE == 3) {
                    DMMA15(rrg, s01.x, s01.y, s23.x, bf, llg, s23.y, r89.x, r89.y, c12, d12, e13, e14, rb, fy);
                } else {
                    const double u0 = rrg * s01.x, u1 = rrg * s01.y, u2 = rrg * s23.x, u3 = rrg * s23.y;
                    const double v0 = bf * s01.x, v1 = bf * s01.y, v2 = bf * s23.x, v3 = bf * s23.y;
                    const double a10 = llg * r89.x, a11 = llg * r89.y, a12 = c12 * d12, a13 = e13 * rb, a14 = e14 * rb;
                    DMMA15(u0, u1, u2, u3, v0, v1, v2, v3, a10, a11, a12, a13, a14, bf, fy);
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < 15; t++) s += acc[t][0] + acc[t][1];
#pragma unroll
    for (int i = 0; i < 13; i++) s += m[i] + pu[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(double* out, int sms, int ctas_per_sm, double ghz) {
    const size_t smem = 4 * (32 * kStride + 32 * kSm) * sizeof(double);
    cudaFuncSetAttribute(k_mix<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * ctas_per_sm;
    k_mix<MODE><<<blocks, 128, smem>>>(out, 1.0000001, 0);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode %d: launch failed\n", MODE); return; }
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k_mix<MODE><<<blocks, 128, smem>>>(out, 1.0000001, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // one warp per sub-partition and CTA: ctas_per_sm warps share a sub-partition
    const double cyc = best * 1e-3 * ghz * 1e9 / ((double)ITERS * ctas_per_sm);
    printf("mode %d, %d warps per sub-partition: %7.1f pipe cycles per k-step (240 = DMMAs alone), %.3f ms\n", MODE, ctas_per_sm, cyc, best);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const double ghz = prop.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz (nominal; cycles are computed with it)\n", prop.name, sms, ghz);
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 128);
    for (int c = 1; c <= 4; c++) {
        run<0>(out, sms, c, ghz); run<1>(out, sms, c, ghz); run<2>(out, sms, c, ghz);
        run<3>(out, sms, c, ghz); run<4>(out, sms, c, ghz); run<5>(out, sms, c, ghz);
    }
    return 0;
}
