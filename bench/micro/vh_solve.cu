// Stand-alone check and timing of vh_lstsq (csrc/vh_lstsq.cuh) on a corpus of normal-equation systems written by
// bench/micro/vh_corpus.py: numerical rank and solution against numpy.linalg.lstsq, microseconds per solve for the
// 256-thread (8 lanes per column pair) and 1024-thread (32 lanes) configurations.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o vh_solve vh_solve.cu ; run: ./vh_solve corpus.bin
#include "../../bsde-tensor-networks_b200/csrc/vh_lstsq.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace bsde;

struct AMat { const double* a; int n; __device__ double operator()(int i, int j) const { return a[i * n + j]; } };

template <int LP, int UMAX, int NT>
__global__ void __launch_bounds__(NT, 1) k_vh(const double* Ag, const double* bg, int n, double* x, int* info, long long* clk) {
    extern __shared__ double sm[];
    const bool a_global = n > 100;               // larger systems: the matrix is read from global memory (as the product kernel reads its moments)
    double* As = sm;
    double* bs = As + (a_global ? 0 : n * n);
    double* xs = bs + n;
    double* base = xs + n; if ((size_t)(base - sm) & 1) base++;
    VhScratch S = vh_carve(base, n);
#ifdef VH_PROFILE
    __shared__ long long s_clk[16];
    if (threadIdx.x < 16) s_clk[threadIdx.x] = 0;
    S.clk = s_clk;
#endif
    if (!a_global) for (int e = threadIdx.x; e < n * n; e += blockDim.x) As[e] = Ag[e];
    for (int e = threadIdx.x; e < n; e += blockDim.x) bs[e] = bg[e];
    __syncthreads();
    const long long t0 = clock64();
    __shared__ int s_sweeps;
    if (threadIdx.x == 0) s_sweeps = -1;
    const int rank = vh_lstsq<NT, LP, UMAX>(AMat{a_global ? Ag : As, n}, n, bs, xs, S, &s_sweeps);
    __syncthreads();
    const long long t1 = clock64();
    for (int e = threadIdx.x; e < n; e += blockDim.x) x[e] = xs[e];
    if (threadIdx.x == 0) { info[0] = rank; info[1] = s_sweeps; clk[0] = t1 - t0;
#ifdef VH_PROFILE
        for (int k = 0; k < 16; k++) clk[1 + k] = s_clk[k];
#endif
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
    FILE* f = fopen(argc > 1 ? argv[1] : "vh_corpus.bin", "rb");
    if (!f) { printf("no corpus\n"); return 1; }
    int count = 0;
    if (fread(&count, 4, 1, f) != 1) return 1;
    double *dA, *db, *dx; int* dinfo; long long* dclk;
    CK(cudaMalloc(&dA, 144 * 144 * 8)); CK(cudaMalloc(&db, 144 * 8)); CK(cudaMalloc(&dx, 144 * 8)); CK(cudaMalloc(&dinfo, 16)); CK(cudaMalloc(&dclk, 256));
    CK(cudaFuncSetAttribute(k_vh<8, 4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_vh<32, 3, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_vh<16, 2, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_vh<16, 4, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double worst[4] = {0, 0, 0, 0}; int mism[4] = {0, 0, 0, 0}; double us_sum[4] = {0, 0, 0, 0}; int nsys = 0, nsys_big = 0; double us_big[2] = {0, 0};
    for (int s = 0; s < count; s++) {
        int n, rank_ref;
        if (fread(&n, 4, 1, f) != 1 || fread(&rank_ref, 4, 1, f) != 1) break;
        std::vector<double> A(n * n), b(n), xref(n), x(n);
        if (fread(A.data(), 8, n * n, f) != (size_t)(n * n) || fread(b.data(), 8, n, f) != (size_t)n || fread(xref.data(), 8, n, f) != (size_t)n) break;
        CK(cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice));
        const size_t smem = ((n > 100 ? 0 : (size_t)n * n) + 2 * n + vh_scratch_doubles(n) + 2) * 8;
        printf("sys %3d n=%3d rank_ref=%3d |", s, n, rank_ref);
        for (int cfg = 0; cfg < 4; cfg++) {
            if (cfg == 0 && n > 64) { printf(" (256 thr: n > 64 skipped) |"); continue; }
            if (cfg == 2 && n > 64) { printf(" (512: skipped) |"); continue; }
            if (cfg == 3 && (n <= 64 || n > 128)) continue;       // 1024 threads x 16 lanes per pair: 64 < n <= 128 only
            auto launch = [&]() {
                if (cfg == 0) k_vh<8, 4, 256><<<1, 256, smem>>>(dA, db, n, dx, dinfo, dclk);
                else if (cfg == 1) k_vh<32, 3, 1024><<<1, 1024, smem>>>(dA, db, n, dx, dinfo, dclk);
                else if (cfg == 2) k_vh<16, 2, 512><<<1, 512, smem>>>(dA, db, n, dx, dinfo, dclk);
                else k_vh<16, 4, 1024><<<1, 1024, smem>>>(dA, db, n, dx, dinfo, dclk);
            };
            launch(); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            for (int it = 0; it < 10; it++) launch();
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            int info[2]; long long clk; long long pc[17];
            CK(cudaMemcpy(x.data(), dx, n * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(info, dinfo, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(pc, dclk, 136, cudaMemcpyDeviceToHost)); clk = pc[0];
            // error in the A-seminorm (fit space) relative to the reference's
            std::vector<double> dxv(n); for (int i = 0; i < n; i++) dxv[i] = x[i] - xref[i];
            double num = 0, den = 0, nx = 0, nr = 0;
            for (int i = 0; i < n; i++) { double a1 = 0, a2 = 0; for (int j = 0; j < n; j++) { a1 += A[i * n + j] * dxv[j]; a2 += A[i * n + j] * xref[j]; } num += dxv[i] * a1; den += xref[i] * a2; nx += dxv[i] * dxv[i]; nr += xref[i] * xref[i]; }
            const double fiterr = sqrt(fabs(num) / fmax(den, 1e-300)), xerr = sqrt(nx / fmax(nr, 1e-300));
            printf(" rank %3d sweeps %2d %7.1f us (%lld cyc) fit-err %.1e x-err %.1e |", info[0], info[1], ms * 100.0, clk, fiterr, xerr);
#ifdef VH_PROFILE
            printf(" [chol %lld idx %lld load+dot %lld shfl %lld rot %lld apply %lld barrier %lld | chol: pre %lld argmax %lld rsqrt+dot %lld shfl %lld finish %lld barrier %lld] ", pc[1], pc[2], pc[3], pc[4], pc[5], pc[6], pc[7], pc[9], pc[10], pc[11], pc[12], pc[13], pc[14]);
#endif
            if (info[0] != rank_ref) mism[cfg]++;
            if (fiterr > worst[cfg]) worst[cfg] = fiterr;
            us_sum[cfg] += ms * 100.0;
            if (n > 64 && n <= 128 && (cfg == 1 || cfg == 3)) us_big[cfg == 3] += ms * 100.0;
        }
        if (n > 64 && n <= 128) nsys_big++;
        printf("\n");
        nsys++;
    }
    printf("SUMMARY systems %d | 256thr: mean %.1f us rank mismatches %d worst fit-err %.1e | 1024thr: mean %.1f us mism %d worst %.1e | 512thr: mean %.1f us mism %d worst %.1e\n",
           nsys, us_sum[0] / nsys, mism[0], worst[0], us_sum[1] / nsys, mism[1], worst[1], us_sum[2] / nsys, mism[2], worst[2]);
    printf("SUMMARY 64 < n <= 128: systems %d | 1024 thr x 32 lanes (round-robin): mean %.1f us | 1024 thr x 16 lanes (odd-even): mean %.1f us mism %d worst fit-err %.1e\n",
           nsys_big, us_big[0] / (nsys_big ? nsys_big : 1), us_big[1] / (nsys_big ? nsys_big : 1), mism[3], worst[3]);
    return 0;
}
