// K1 / K8 — forward path simulation, terminal condition, regression targets and small streaming helpers.
//
//   k_paths      Euler-Maruyama with b = 0 and diagonal sigma          (stochastic/path.py:26-32, bsde.py:41-46, 86-90)
//                x_n = (x_{n-1} + 0*dt) + sqrt(dt) * ((sigma0 * s(x_{n-1})) * xi_n),  s = x (BlackScholes) or 1 (HJB):
//                the reference's dense (N,d,d) sigma only adds exact zeros, so this is bit-identical to it on the
//                same increments (replay mode).  Production mode draws xi from Philox4x32-10, a counter-based
//                generator keyed by the seed and indexed by (global sample, asset, step pair) — no state, any
//                sample shard on any GPU regenerates exactly its own stream.
//   k_terminal   g(x): ||x||^2 (bsde.py:51-52) / log(1/2 + 1/2 ||x||^2) (bsde.py:95-96)
//   k_target     h(x,t,y,z)*dt + y_next [- sqrt(dt) z.xi]              (src/scripts/pipeline.py:129-133,
//                                                                       src/scripts/pipeline_explicit.py:97-102)
//   k_transpose, k_basis_eval, block-deterministic mean.
// All HBM-bound: one thread per (asset, sample) or per sample, sample-contiguous accesses.
#include "common.cuh"
#include "kernels.cuh"

namespace bsde {

namespace {

constexpr int kThreads = 256;

// ---- Philox4x32-10 (Salmon et al., SC'11) ---------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// Two standard normals from one Philox block: u1 in (0,1], u2 in [0,1) with 53 bits each, Box-Muller.
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint64_t sample, uint32_t asset, uint32_t pair,
                                                   double& n0, double& n1) {
    uint32_t c[4] = {(uint32_t)sample, (uint32_t)(sample >> 32), asset, pair};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)c[1] << 32) | c[0];
    const uint64_t b = ((uint64_t)c[3] << 32) | c[2];
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    n0 = rad * cs;
    n1 = rad * sn;
}

struct PathArgs {
    int model;           // 0 BlackScholes, 1 HJB
    double sigma0;
    double sqrt_dt;
    int d, steps;
    int64_t N;
    const double* x0;    // [d] device: one starting point for every sample; null: slab 0 of x holds per-sample starting points
    uint64_t seed;
    int64_t sample_offset;
    int replay;
    double* xi;          // [steps+1][d][N] or null
    double* x;           // [steps+1][d][N]
};

__global__ void __launch_bounds__(kThreads) k_paths(PathArgs a) {
    const int64_t total = (int64_t)a.d * a.N;
    const int64_t slab = total;     // one time step
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(t / a.N);
        const int64_t s = t - (int64_t)j * a.N;
        double x = a.x0 ? a.x0[j] : a.x[t];
        a.x[t] = x;
        if (!a.replay && a.xi) a.xi[t] = 0.0;     // xi[:,0] is never used (path.py:34)
        double nrm0 = 0.0, nrm1 = 0.0;
        for (int n = 1; n <= a.steps; n++) {
            double xi;
            if (a.replay) {
                xi = a.xi[(int64_t)n * slab + t];
            } else {
                if (n & 1) philox_normal_pair(a.seed, (uint64_t)(a.sample_offset + s), (uint32_t)j, (uint32_t)(n >> 1), nrm0, nrm1);
                xi = (n & 1) ? nrm0 : nrm1;
                if (a.xi) a.xi[(int64_t)n * slab + t] = xi;
            }
            const double sig = (a.model == 0) ? __dmul_rn(a.sigma0, __dmul_rn(x, 1.0)) : a.sigma0;
            // (x + b*dt) + sqrt(dt) * (sigma * xi), b = 0 — separate roundings as numpy does (no FMA contraction)
            x = __dadd_rn(__dadd_rn(x, 0.0), __dmul_rn(a.sqrt_dt, __dmul_rn(sig, xi)));
            a.x[(int64_t)n * slab + t] = x;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
k_terminal(int model, int d, int64_t N, const double* __restrict__ x, double* __restrict__ out) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        double ss = 0.0;
        for (int j = 0; j < d; j++) {
            const double v = x[(int64_t)j * N + s];
            ss = __dadd_rn(ss, __dmul_rn(v, v));
        }
        if (model == 0) {
            const double nr = sqrt(ss);          // numpy: linalg.norm(x, axis=1) ** 2
            out[s] = __dmul_rn(nr, nr);
        } else {
            out[s] = log(__dadd_rn(0.5, __dmul_rn(0.5, ss)));
        }
    }
}

struct TargetArgs {
    int model;
    double r, sigma0, dt, sqrt_dt;
    int d;
    int64_t N;
    const double* x; const double* y; const double* y_next; const double* grad; const double* xi;
    int apply_sigma;
    double* out;
};

__global__ void __launch_bounds__(kThreads) k_target(TargetArgs a) {
    const int64_t N = a.N;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0, zxi = 0.0;
        for (int j = 0; j < a.d; j++) {
            const int64_t o = (int64_t)j * N + s;
            const double xv = a.x[o];
            double z = a.grad[o];
            if (a.apply_sigma) z = __dmul_rn((a.model == 0) ? __dmul_rn(a.sigma0, __dmul_rn(xv, 1.0)) : a.sigma0, z);
            acc = __dadd_rn(acc, (a.model == 0) ? __dmul_rn(xv, z) : __dmul_rn(z, z));
            if (a.xi) zxi = __dadd_rn(zxi, __dmul_rn(z, a.xi[o]));
        }
        const double h = (a.model == 0) ? __dmul_rn(-a.r, __dadd_rn(a.y[s], -acc)) : __dmul_rn(-0.5, acc);
        double t = __dadd_rn(__dmul_rn(h, a.dt), a.y_next[s]);
        if (a.xi) t = __dadd_rn(t, -__dmul_rn(a.sqrt_dt, zxi));
        a.out[s] = t;
    }
}

__global__ void k_transpose(const double* __restrict__ src, int64_t rows, int64_t cols, double* __restrict__ dst) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[c * rows + r] = tile[threadIdx.x][i];
    }
}

template <int PB>
__global__ void __launch_bounds__(kThreads) k_basis_eval(Inputs in, int deriv, double* __restrict__ out) {
    const int64_t N = in.N;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        double phi[PB];
        basis_values<PB>(in, 0, s, deriv, phi);
#pragma unroll
        for (int m = 0; m < PB; m++)
            if (m < in.p) out[(int64_t)m * N + s] = phi[m];
    }
}

// per-block partial sums in a fixed order; combined by launch_sum_partials
__global__ void __launch_bounds__(kThreads) k_block_sums(const double* __restrict__ v, int64_t N, double* __restrict__ partial) {
    __shared__ double scratch[kThreads / 32];
    double loc = 0.0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) loc += v[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loc += __shfl_down_sync(0xffffffffu, loc, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = loc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < kThreads / 32; k++) t += scratch[k];
        partial[blockIdx.x] = t;
    }
}

inline int grid_for(int64_t n) {
    int64_t b = (n + kThreads - 1) / kThreads;
    const int64_t cap = 148 * 8;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace

int launch_paths(int model, double sigma0, double sqrt_dt, int d, int64_t N, int steps, const double* x0_dev,
                 uint64_t seed, int64_t sample_offset, int replay, double* xi, double* x, cudaStream_t st) {
    PathArgs a{model, sigma0, sqrt_dt, d, steps, N, x0_dev, seed, sample_offset, replay, xi, x};
    k_paths<<<grid_for((int64_t)d * N), kThreads, 0, st>>>(a);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_terminal(int model, int d, int64_t N, const double* x, double* out, cudaStream_t st) {
    k_terminal<<<grid_for(N), kThreads, 0, st>>>(model, d, N, x, out);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_target(int model, double r, double sigma0, double dt, int d, int64_t N, const double* x, const double* y,
                  const double* y_next, const double* grad, const double* xi, int apply_sigma, double* out,
                  cudaStream_t st) {
    TargetArgs a{model, r, sigma0, dt, sqrt(dt), d, N, x, y, y_next, grad, xi, apply_sigma, out};
    k_target<<<grid_for(N), kThreads, 0, st>>>(a);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_transpose(const double* src, int64_t rows, int64_t cols, double* dst, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    // row tiles on grid.x (2^31-1 limit) because the typical call is (N, d) with N in the millions
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    if (grid.y > 65535) return fail_msg(2, "transpose: more than %d columns not supported", 65535 * 32);
    k_transpose<<<grid, dim3(32, 8), 0, st>>>(src, rows, cols, dst);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_basis_eval(const Inputs& in, int deriv, double* out, cudaStream_t st) {
    if (in.p > kMaxP) return fail_msg(2, "basis_eval: p=%d exceeds %d", in.p, kMaxP);
    if (in.p <= 4) k_basis_eval<4><<<grid_for(in.N), kThreads, 0, st>>>(in, deriv, out);
    else k_basis_eval<8><<<grid_for(in.N), kThreads, 0, st>>>(in, deriv, out);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int mean_grid(int64_t N) { return grid_for(N); }

int launch_block_sums(const double* v, int64_t N, double* partial, cudaStream_t st) {
    k_block_sums<<<grid_for(N), kThreads, 0, st>>>(v, N, partial);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace bsde
