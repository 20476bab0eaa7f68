// Latency study of the in-CTA Cholesky (n = 64, augmented with 2 right-hand-side rows) used by k_micro_solve.
// Variants: V0 current (T threads, 1/piv by division, one __syncthreads per column)
//           V1 rsqrt instead of division
//           V2 V1 + two columns per barrier (rank-2 trailing update)
//           V3 one warp, no block barrier
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chol64 chol64.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <cmath>

constexpr int N = 64, LD = N + 1, ROWS = N + 2;

template <int T, int V>
__global__ void __launch_bounds__(T) k_chol(const double* Ain, double* Aout, long long* cycles) {
    __shared__ double A[ROWS * LD];
    __shared__ double dinv[N];
    const int tid = threadIdx.x;
    for (int e = tid; e < ROWS * LD; e += T) A[e] = Ain[e];
    __syncthreads();
    const long long t0 = clock64();
    constexpr int G = (T >= 1024) ? 32 : (T >= 256 ? 16 : 8);       // T = G * (T / G)
    const int ty = tid / G, tx = tid % G;
    constexpr int RG = T / G;
    if (V == 0 || V == 1) {
        for (int k = 0; k < N; k++) {
            const double piv = A[k * LD + k];
            double ipiv;
            if (V == 0) ipiv = 1.0 / piv;
            else { const double r = rsqrt(piv); ipiv = r * r; }
            if (tid == 0) dinv[k] = sqrt(ipiv);
            for (int i = k + 1 + ty; i < ROWS; i += RG) {
                const double lik = A[i * LD + k] * ipiv;
                const int jmax = (i < N) ? i : N - 1;
                for (int j = k + 1 + tx; j <= jmax; j += G) A[i * LD + j] = fma(-lik, A[j * LD + k], A[i * LD + j]);
            }
            __syncthreads();
        }
    } else if (V == 2) {
        for (int k = 0; k < N; k += 2) {
            // column k, then column k+1 of the 2-wide panel (every thread redundantly computes the 2x2 pivot block)
            const double p0 = A[k * LD + k];
            const double r0 = rsqrt(p0), i0 = r0 * r0;
            const double a10 = A[(k + 1) * LD + k];
            const double p1 = fma(-a10 * i0, a10, A[(k + 1) * LD + k + 1]);
            const double r1 = rsqrt(p1), i1 = r1 * r1;
            if (tid == 0) { dinv[k] = r0; dinv[k + 1] = r1; }
            // rank-2 update of the trailing part: for i > k+1: u_i = A_ik (col k), v_i = A_i,k+1 - A_ik*i0*a10 (updated col k+1)
            for (int i = k + 2 + ty; i < ROWS; i += RG) {
                const double ui = A[i * LD + k];
                const double vi = fma(-ui * i0, a10, A[i * LD + k + 1]);
                const double li0 = ui * i0, li1 = vi * i1;
                const int jmax = (i < N) ? i : N - 1;
                for (int j = k + 2 + tx; j <= jmax; j += G) {
                    const double uj = A[j * LD + k];
                    const double vj = fma(-uj * i0, a10, A[j * LD + k + 1]);
                    A[i * LD + j] = fma(-li1, vj, fma(-li0, uj, A[i * LD + j]));
                }
            }
            __syncthreads();
            // write back the updated column k+1 (needed by the back substitution), one thread per row
            for (int i = k + 2 + tid; i < ROWS; i += T) A[i * LD + k + 1] = fma(-A[i * LD + k] * i0, a10, A[i * LD + k + 1]);
            if (tid == 0) A[(k + 1) * LD + k + 1] = p1;
            __syncthreads();
        }
    } else {   // one warp
        if (tid < 32) {
            for (int k = 0; k < N; k++) {
                const double piv = A[k * LD + k];
                const double r = rsqrt(piv), ipiv = r * r;
                if (tid == 0) dinv[k] = r;
                for (int i = k + 1; i < ROWS; i++) {
                    const double lik = A[i * LD + k] * ipiv;
                    const int jmax = (i < N) ? i : N - 1;
                    for (int j = k + 1 + tid; j <= jmax; j += 32) A[i * LD + j] = fma(-lik, A[j * LD + k], A[i * LD + j]);
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) *cycles = t1 - t0;
    for (int e = tid; e < ROWS * LD; e += T) Aout[e] = A[e];
    if (tid < N) Aout[ROWS * LD + tid] = dinv[tid];
}

template <int T, int V>
void run(const double* dA, double* dOut, long long* dC, const std::vector<double>& ref, const char* name) {
    k_chol<T, V><<<1, T>>>(dA, dOut, dC);
    k_chol<T, V><<<1, T>>>(dA, dOut, dC);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dC, sizeof c, cudaMemcpyDeviceToHost);
    std::vector<double> out(ROWS * LD + N);
    cudaMemcpy(out.data(), dOut, out.size() * sizeof(double), cudaMemcpyDeviceToHost);
    // check: L_ik = out[i][k] * dinv[k] against the reference factor
    double err = 0;
    for (int i = 0; i < N; i++)
        for (int k = 0; k <= i; k++) {
            const double l = (i == k) ? 1.0 / out[ROWS * LD + k] : out[i * LD + k] * out[ROWS * LD + k];
            err = fmax(err, fabs(l - ref[i * N + k]));
        }
    printf("%-28s threads %4d: %7lld cycles (%.1f us at 1.965 GHz)  max |L - L_ref| = %.2e  %s\n", name, T, c, c / 1965.0, err,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    std::vector<double> M(200 * N), A(ROWS * LD, 0.0), ref(N * N, 0.0);
    srand(1);
    for (auto& v : M) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            double s = 0;
            for (int q = 0; q < 200; q++) s += M[q * N + i] * M[q * N + j];
            A[i * LD + j] = s;
        }
    for (int j = 0; j < N; j++) { A[N * LD + j] = 1.0; A[(N + 1) * LD + j] = (j & 1) ? 1.0 : -1.0; }
    // reference factor on the host
    std::vector<double> W(N * N);
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) W[i * N + j] = A[i * LD + j];
    for (int k = 0; k < N; k++) {
        ref[k * N + k] = sqrt(W[k * N + k]);
        for (int i = k + 1; i < N; i++) ref[i * N + k] = W[i * N + k] / ref[k * N + k];
        for (int i = k + 1; i < N; i++) for (int j = k + 1; j <= i; j++) W[i * N + j] -= ref[i * N + k] * ref[j * N + k];
    }
    double *dA, *dOut; long long* dC;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dOut, (ROWS * LD + N) * 8); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    run<1024, 0>(dA, dOut, dC, ref, "division");
    run<1024, 1>(dA, dOut, dC, ref, "rsqrt");
    run<256, 1>(dA, dOut, dC, ref, "rsqrt");
    run<64, 1>(dA, dOut, dC, ref, "rsqrt");
    run<1024, 2>(dA, dOut, dC, ref, "rsqrt, 2 columns/barrier");
    run<256, 2>(dA, dOut, dC, ref, "rsqrt, 2 columns/barrier");
    run<64, 2>(dA, dOut, dC, ref, "rsqrt, 2 columns/barrier");
    run<32, 3>(dA, dOut, dC, ref, "one warp");
    return 0;
}
