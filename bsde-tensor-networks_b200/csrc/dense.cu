// K5 — the small dense algebra of one micro-step, entirely inside one CTA (no host round trip):
//   * expansion of the moment tensors (assembly.cu) into the normal matrix A (n x n) and right-hand side b
//   * solver(A, b, method)                       reference: core/optimisation/solve.py:6-19
//       'lstsq' = numpy.linalg.lstsq(rcond=None): minimum-norm solution, singular values <= eps*n*s_max dropped.
//         Fast path: equilibrated Cholesky with the forward substitution fused into the factorisation, taken only
//         when the pivots prove the matrix is far from the truncation threshold (then lstsq == exact solve).
//         Otherwise: one-sided Jacobi SVD of A with LAPACK gelsd's truncation rule (rank-deficient systems such
//         as the step-0 fit where every sample is identical).
//       'lu' / 'solve' / 'qr' = Gaussian elimination with partial pivoting.
//   * Householder QR (LAPACK dgeqr2/dorg2r sign conventions) of the updated core and the push of the
//     triangular factor into the neighbouring core            reference: als.py:70-77 and :87-94
//   * right/left orthonormalisation of a whole train            reference: tensor_train.py:34-87
#include "common.cuh"
#include "kernels.cuh"
#include "vh_lstsq.cuh"
#include <cfloat>

namespace bsde {

namespace {

constexpr int kDenseThreads = 1024;   // 32 warps: one Jacobi pair per warp for n = 64; the Cholesky path uses 8 of them

__device__ __forceinline__ int pair_index(int a, int b, int r) {   // a <= b < r, row-major upper triangle
    return a * r - (a * (a - 1)) / 2 + (b - a);
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Householder QR of M (m x k, row-major with leading dimension ld, m >= k) executed by ONE warp.
// On exit M holds Q (m x k, explicit) and Rm (k x k, row-major, ld k) the upper-triangular factor.
// Sign convention = LAPACK dlarfg: beta = -sign(alpha) * norm, so diag(R) may be negative exactly as numpy's qr.
__device__ void warp_householder_qr(double* M, int m, int k, int ld, double* Rm, double* tau, double* vbuf) {
    const int lane = threadIdx.x & 31;
    for (int c = 0; c < k; c++) {
        double ss = 0.0;
        for (int i = c + 1 + lane; i < m; i += 32) ss = fma(M[i * ld + c], M[i * ld + c], ss);
        ss = warp_sum(ss);
        const double alpha = M[c * ld + c];
        double t = 0.0, beta = alpha;
        if (ss != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + ss), alpha);
            t = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
            for (int i = c + 1 + lane; i < m; i += 32) M[i * ld + c] *= sc;
        }
        __syncwarp();
        if (lane == 0) { tau[c] = t; M[c * ld + c] = beta; }
        __syncwarp();
        // apply H = I - t v v^T (v_c = 1) to the trailing columns
        for (int cc = c + 1; cc < k; cc++) {
            double w = 0.0;
            for (int i = c + 1 + lane; i < m; i += 32) w = fma(M[i * ld + c], M[i * ld + cc], w);
            w = warp_sum(w) + M[c * ld + cc];
            w *= t;
            for (int i = c + 1 + lane; i < m; i += 32) M[i * ld + cc] = fma(-w, M[i * ld + c], M[i * ld + cc]);
            __syncwarp();
            if (lane == 0) M[c * ld + cc] -= w;
            __syncwarp();
        }
    }
    // R
    for (int e = lane; e < k * k; e += 32) {
        const int r = e / k, c = e % k;
        Rm[e] = (c >= r) ? M[r * ld + c] : 0.0;
    }
    __syncwarp();
    // Q = H_0 ... H_{k-1} [I_k; 0]  (dorg2r: backwards accumulation)
    for (int c = k - 1; c >= 0; c--) {
        const double t = tau[c];
        for (int i = c + 1 + lane; i < m; i += 32) vbuf[i] = M[i * ld + c];
        __syncwarp();
        // columns cc > c currently hold Q columns restricted to rows >= c+1 (row c of them is zero before H_c)
        for (int cc = c + 1; cc < k; cc++) {
            double w = 0.0;
            for (int i = c + 1 + lane; i < m; i += 32) w = fma(vbuf[i], M[i * ld + cc], w);
            w = warp_sum(w) * t;      // row c of column cc is 0 before applying H_c
            for (int i = c + 1 + lane; i < m; i += 32) M[i * ld + cc] = fma(-w, vbuf[i], M[i * ld + cc]);
            if (lane == 0) M[c * ld + cc] = -w;
            __syncwarp();
        }
        // column c itself: H_c e_c = e_c - t v
        for (int i = c + 1 + lane; i < m; i += 32) M[i * ld + c] = -t * vbuf[i];
        if (lane == 0) M[c * ld + c] = 1.0 - t;
        for (int i = lane; i < c; i += 32) M[i * ld + c] = 0.0;
        __syncwarp();
    }
}

// 1 / d to ~1 ulp without the IEEE division sequence (approximate reciprocal + two Newton steps): FP64 has ~75 cycles of
// dependent-issue latency here, so the length of the dependent chain is what a one-warp factorisation costs.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(fma(-d, r, 1.0), r, r);
    r = fma(fma(-d, r, 1.0), r, r);
    return r;
}

// Reciprocal from the 20-bit hardware approximation and ONE third-order step: r (1 + e + e^2), e = 1 - d r; |e| <= 2^-23, so
// the truncation error e^3 is below 2^-69 and the result is the correctly rounded reciprocal up to the final rounding.
__device__ __forceinline__ double fast_rcp3(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    return fma(r, fma(e, e, e), r);
}

// Householder QR of M (m x k row-major, ld; m <= 64, k <= KM) by one warp with the matrix held in registers: lane l owns
// rows l and l + 32.  Same reflectors and sign convention as warp_householder_qr (LAPACK dgeqr2 / dorg2r), but every
// column costs two batched warp reductions instead of one per trailing column, and no shared-memory round trips.
template <int KM>
__device__ void warp_householder_qr_regs(double* M, int m, int k, int ld, double* Rm) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool h0 = lane < m, h1 = lane + 32 < m;
    const bool small = m <= 16;          // lanes 16.. hold zeros: their partial sums are zero and every lane below 16 is
                                         // complete after the 8, 4, 2, 1 steps (results are only read from lanes < m)
    double x0[KM], x1[KM], tau[KM];
#pragma unroll
    for (int c = 0; c < KM; c++) {
        x0[c] = (h0 && c < k) ? M[lane * ld + c] : 0.0;
        x1[c] = (h1 && c < k) ? M[(lane + 32) * ld + c] : 0.0;
        tau[c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < KM; c++) {
        if (c >= k) break;
        double ss = ((lane > c) ? x0[c] * x0[c] : 0.0) + x1[c] * x1[c];
        if (small) {                                          // rows live in lanes 0..15 only: four butterfly steps
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(full, ss, o);
        } else {
            ss = warp_sum(ss);
        }
        const double alpha = __shfl_sync(full, x0[c], c);
        double t = 0.0, beta = alpha, scale = 0.0;
        if (ss != 0.0) {
            const double s2 = fma(alpha, alpha, ss);
            const double rs = rsqrt(s2);
            beta = -copysign(s2 * rs, alpha);
            t = fma(fabs(alpha), rs, 1.0);                   // (beta - alpha) / beta
            scale = fast_rcp(alpha - beta);
        }
        tau[c] = t;
        if (lane > c) x0[c] *= scale;
        x1[c] *= scale;
        if (lane == c) x0[c] = beta;
        const double v0 = (lane > c) ? x0[c] : (lane == c ? 1.0 : 0.0), v1 = x1[c];
        double w[KM];
#pragma unroll
        for (int cc = 0; cc < KM; cc++) w[cc] = (cc > c && cc < k) ? fma(v0, x0[cc], v1 * x1[cc]) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            if (o == 16 && small) continue;
#pragma unroll
            for (int cc = 0; cc < KM; cc++)
                if (cc > c) w[cc] += __shfl_xor_sync(full, w[cc], o);
        }
#pragma unroll
        for (int cc = 0; cc < KM; cc++)
            if (cc > c && cc < k) {
                const double tw = t * w[cc];
                x0[cc] = fma(-tw, v0, x0[cc]);
                x1[cc] = fma(-tw, v1, x1[cc]);
            }
    }
    // R: row c lives in lane c
#pragma unroll
    for (int c = 0; c < KM; c++)
        if (c < k && lane < k) Rm[lane * k + c] = (c >= lane) ? x0[c] : 0.0;
    // Q = H_0 ... H_{k-1} [I_k; 0]
    double q0[KM], q1[KM];
#pragma unroll
    for (int c = 0; c < KM; c++) { q0[c] = (lane == c && c < k) ? 1.0 : 0.0; q1[c] = 0.0; }
#pragma unroll
    for (int c = KM - 1; c >= 0; c--) {
        if (c >= k) continue;
        const double v0 = (lane > c) ? x0[c] : (lane == c ? 1.0 : 0.0), v1 = x1[c];
        double w[KM];
#pragma unroll
        for (int cc = 0; cc < KM; cc++) w[cc] = (cc >= c && cc < k) ? fma(v0, q0[cc], v1 * q1[cc]) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            if (o == 16 && small) continue;
#pragma unroll
            for (int cc = 0; cc < KM; cc++)
                if (cc >= c) w[cc] += __shfl_xor_sync(full, w[cc], o);
        }
#pragma unroll
        for (int cc = 0; cc < KM; cc++)
            if (cc >= c && cc < k) {
                const double tw = tau[c] * w[cc];
                q0[cc] = fma(-tw, v0, q0[cc]);
                q1[cc] = fma(-tw, v1, q1[cc]);
            }
    }
#pragma unroll
    for (int c = 0; c < KM; c++)
        if (c < k) {
            if (h0) M[lane * ld + c] = q0[c];
            if (h1) M[(lane + 32) * ld + c] = q1[c];
        }
    __syncwarp();
}

// Slot e of the reduced moments: the sum, in fixed order, of the a.moment_rows (<= kRedGroups) group rows k_moment_reduce
// left behind.  All loads are issued before the first addition.
__device__ __forceinline__ double load_moment(const MicroSolveArgs& a, int e) {
    double t[kRedGroups];
#pragma unroll
    for (int g = 0; g < kRedGroups; g++) t[g] = g < a.moment_rows ? a.moments[(size_t)g * a.n_slots + e] : 0.0;
    double acc = t[0];
#pragma unroll
    for (int g = 1; g < kRedGroups; g++) acc += t[g];
    return acc;
}

// Diagonal entry `row` of the system matrix from the moment value v: the optional diagonal vector / ridge and SALSA's
// regulariser, added in the reference's order  A += reg (= eta^2 Gamma-term + eta^2 Theta-term);  A += eta I  (als.py:143-174).
// Gamma term: index (a, i, k) of (r_l, p, r_r) -> gamma[a]^2.  Theta term: the reference flat-reshapes a (r_r, p, r_l)^2
// tensor, so the diagonal entry e carries theta[e / (p * r_l)]^2.
__device__ __forceinline__ double regularised_diagonal(const MicroSolveArgs& a, int row, double v) {
    if (a.reg_diag) v += a.reg_diag[row];
    v += a.ridge;
    if (a.eta_ptr) {
        const double eta = *a.eta_ptr;
        if (a.gamma && a.theta) {
            double reg = 0.0;
            if (a.reg_left) { const double g = a.gamma[row / (a.p * a.rr)]; reg += eta * eta * g * g; }
            if (a.reg_right) { const double t = a.theta[row / (a.p * a.rl)]; reg += eta * eta * t * t; }
            v += reg;
        }
        v += eta;
    }
    return v;
}
__device__ __forceinline__ bool has_diagonal_terms(const MicroSolveArgs& a) {
    return a.reg_diag != nullptr || a.ridge != 0.0 || a.eta_ptr != nullptr;
}

// 2^-k with 2^k <= m < 2^(k+1): multiplying the system by it is exact and brings its largest diagonal entry into [1, 2).
// The normal equations of a long train can sit anywhere in the double range (interfaces are products of d factors: 1e-160
// at d = 100 for Black-Scholes samples), and the acceptance tests and the Jacobi phase square their entries.
__device__ __forceinline__ double pow2_normaliser(double m) {
    if (!(m > 0.0) || !isfinite(m)) return 1.0;
    int ex = (__double2hiint(m) >> 20) & 0x7ff;
    if (ex == 0) return 0x1p+1000;                       // subnormal diagonal: lift it well into the normal range
    ex = 2046 - ex;                                      // biased exponent of 2^-(ex - 1023)
    if (ex > 2045) ex = 2045;
    if (ex < 1) ex = 1;
    return __hiloint2double(ex << 20, 0);
}

// Entry (i, j) of the system matrix as vh_lstsq reads it: an explicit matrix (bsde_solve), or the staged moments through
// the expansion map with the regularised diagonal held in `diag` (n doubles; null = the plain moments).
struct SystemMatrix {
    const double* direct;     // n x n row-major or null
    const double* stage;
    const uint16_t* map;
    const double* diag;       // already normalised
    int n;
    double scale;             // power-of-two normalisation applied to direct / staged entries (pow2_normaliser)
    __device__ __forceinline__ double operator()(int i, int j) const {
        if (direct) return direct[(size_t)i * n + j] * scale;
        if (i == j && diag) return diag[i];
        return stage[map[i * n + j]] * scale;
    }
};

struct SolveScratch {
    double* A;      // (n+1) x (n+1) augmented, row-major, ld = n+1
    double* G;      // n x n column-major (Jacobi)        — aliases A
    double* V;      // n x n column-major (Jacobi)
    double* x;      // n
    double* sc;     // n  equilibration
    double* bvec;   // n  copy of b
    double* red;    // n  scratch
};

__device__ void expand_normal_equations(const MicroSolveArgs& a, double* A, int ld, double* bvec, double* stage) {
    const int n = a.rl * a.p * a.rr;
    if (a.A_direct) {    // bsde_solve: an explicit system instead of moments (rl = n, p = rr = 1)
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) A[(e / n) * ld + (e % n)] = a.A_direct[e];
        for (int e = threadIdx.x; e < n; e += blockDim.x) bvec[e] = a.b_direct[e];
        return;
    }
    // stage the moments in shared memory with coalesced loads (the gather below would otherwise pay one L2 round trip
    // per entry)
    // (a.moments holds the summed partial-row slots of the assembly kernel; a.n_slots of them)
    for (int e = threadIdx.x; e < a.n_slots; e += blockDim.x) stage[e] = load_moment(a, e);
    __syncthreads();
    // a.expand_map[e] (built once per core shape on the host, assembly.cu) = index of the moment that entry e of the
    // row-major A (then of b) is equal to: A[(a,m,b),(a',m',b')] = M1[beta(m,m')][lambda(a,a')][rho(b,b')]
    const uint16_t* map = a.expand_map;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int row = e / n, col = e - row * n;
        double v = stage[map[e]];
        if (row == col) v = regularised_diagonal(a, row, v);
        A[row * ld + col] = v;
    }
    for (int row = threadIdx.x; row < n; row += blockDim.x) bvec[row] = stage[map[n * n + row]];
}

// Named barrier for the first `count` threads of the CTA (the fast path runs on 8 of the 32 warps).

constexpr int kCholThreads = 1024;  // the factorisation uses the whole CTA: <= 4 trailing entries per thread and column
constexpr double kGateSafety = 100.0;

// Fast path of 'lstsq', executed by threads 0..255: equilibrated Cholesky of A with two right-hand sides carried
// through the factorisation as extra rows — b (row n) and a fixed +-1 probe vector e (row n+1).  The probe gives
// sigma_min(A) <= |e| / |A^-1 e|; with sigma_max(A) <= trace(A) the test
//        (|e| / |A^-1 e|) / trace(A)  >  kGateSafety * eps * n
// certifies that numpy.linalg.lstsq(rcond=None) (cutoff eps*n*sigma_max) would not have truncated any singular
// value, i.e. that the exact solve IS the lstsq answer.  Anything else (non-positive pivot, failed test) returns
// false and the caller takes the SVD path, which reproduces the truncation rule literally.
__device__ bool cholesky_solve(SolveScratch& S, int n, double* xout, double* s_red) {
    const int ld = n + 3;                                     // odd pitch; columns n, n+1 of the upper part are used below
    double* A = S.A;
    const int tid = threadIdx.x;
    // s_red: [0] trace, [1] bad flag, [2] |z|^2
    if (tid < 32) {
        double tr = 0.0;
        int bad = 0;
        for (int i = tid; i < n; i += 32) {
            const double d = A[i * ld + i];
            if (!(d > 0.0) || !isfinite(d)) bad = 1;
            tr += d;
        }
        tr = warp_sum(tr);
        bad = __any_sync(0xffffffffu, bad);
        if (tid == 0) { s_red[0] = tr; s_red[1] = bad ? 1.0 : 0.0; }
    }
    __syncthreads();
    if (s_red[1] != 0.0) return false;
    for (int i = tid; i < n; i += kCholThreads) S.sc[i] = rsqrt(A[i * ld + i]);
    __syncthreads();
    // scale the lower triangle; rows n, n+1 = equilibrated right-hand sides
    for (int e = tid; e < (n + 2) * n; e += kCholThreads) {
        const int i = e / n, j = e - i * n;
        if (i == n) A[i * ld + j] = S.bvec[j] * S.sc[j];
        else if (i == n + 1) A[i * ld + j] = ((((unsigned)j * 2654435761u) >> 7) & 1u ? 1.0 : -1.0) * S.sc[j];
        else if (j <= i) A[i * ld + j] *= S.sc[i] * S.sc[j];
    }
    __syncthreads();
    // Right-looking factorisation, two columns per trailing update (rank-2), ONE block barrier per column pair.
    // FP64 has ~75 cycles of dependent-issue latency on this part (bench/micro/chol64.cu), so the cost is the number of
    // dependent FP64 operations and barriers per pivot:
    //   * every thread recomputes the 2x2 pivot block redundantly instead of waiting for a broadcast;
    //   * reciprocal pivots come from rcp.approx + 2 Newton steps (5 dependent operations instead of rsqrt + square);
    //     1/sqrt(pivot), needed only by the back substitution, is computed off the critical path;
    //   * columns are left unscaled (stored column k = L_ik * sqrt(pivot_k)); the updated second column of a pair is
    //     parked in the unused upper triangle (A[k+1][i]) so that no second barrier is needed to publish it.
    const int ty = tid >> 5, tx = tid & 31;
    double* dinv = S.red;                                     // 1 / sqrt(pivot_k)
    int k = 0;
    for (; k + 1 < n; k += 2) {
        const double p0 = A[k * ld + k];
        if (!(p0 > 0.0)) return false;                        // uniform: every thread reads the same values
        const double i0 = fast_rcp(p0);
        const double a10 = A[(k + 1) * ld + k];
        const double p1 = fma(-a10 * i0, a10, A[(k + 1) * ld + k + 1]);
        if (!(p1 > 0.0)) return false;
        const double i1 = fast_rcp(p1);
        for (int i = k + 2 + ty; i <= n + 1; i += 32) {
            const double ui = A[i * ld + k];
            const double li0 = ui * i0;
            const double vi = fma(-li0, a10, A[i * ld + k + 1]);
            const double li1 = vi * i1;
            if (tx == 0) A[(k + 1) * ld + i] = vi;            // final column k+1, row i -> upper triangle
            const int jmax = (i < n) ? i : n - 1;
            for (int j = k + 2 + tx; j <= jmax; j += 32) {
                const double uj = A[j * ld + k];
                const double vj = fma(-uj * i0, a10, A[j * ld + k + 1]);
                A[i * ld + j] = fma(-li1, vj, fma(-li0, uj, A[i * ld + j]));
            }
        }
        if (tid == 0) { dinv[k] = rsqrt(p0); dinv[k + 1] = rsqrt(p1); }
        __syncthreads();
    }
    if (k < n) {                                              // odd n: last column on its own (stays in the lower part)
        const double piv = A[k * ld + k];
        if (!(piv > 0.0)) return false;
        const double ipiv = fast_rcp(piv);
        if (tid == 0) dinv[k] = rsqrt(piv);
        for (int i = k + 1 + ty; i <= n + 1; i += 32) {
            const double lik = A[i * ld + k] * ipiv;
            const int jmax = (i < n) ? i : n - 1;
            for (int j = k + 1 + tx; j <= jmax; j += 32) A[i * ld + j] = fma(-lik, A[j * ld + k], A[i * ld + j]);
        }
        __syncthreads();
    }
    const int npair = n & ~1;                                 // columns below npair with odd index live in the upper part
    // back substitution L^T x = z for both right-hand sides: warp 0 -> b, warp 1 -> probe.
    // Stored column k holds L_ik * sqrt(pivot_k) (unscaled), the right-hand-side rows hold z_k * sqrt(pivot_k).
    if (tid < 64) {
        const int lane = tid & 31;
        const int zr = n + (tid >> 5);
        double* z = A + zr * ld;
        for (int i = lane; i < n; i += 32) {
            const double raw = ((i & 1) && i < npair) ? A[i * ld + zr] : z[i];
            z[i] = raw * dinv[i];                                              // z_k
        }
        __syncwarp();
        for (int kk = n - 1; kk >= 0; kk--) {
            const double xk = z[kk] * dinv[kk];                                  // x_k = z_k / L_kk
            __syncwarp();
            if (lane == 0) z[kk] = xk;
            // L_ki = (stored column i at row k) * dinv_i, i < k
            for (int i = lane; i < kk; i += 32) {
                const double lki = ((i & 1) && i < npair) ? A[i * ld + kk] : A[kk * ld + i];
                z[i] = fma(-lki * dinv[i], xk, z[i]);
            }
            __syncwarp();
        }
        if (tid < 32) {
            for (int i = lane; i < n; i += 32) xout[i] = z[i] * S.sc[i];
        } else {
            double zz = 0.0;
            for (int i = lane; i < n; i += 32) { const double v = z[i] * S.sc[i]; zz = fma(v, v, zz); }
            zz = warp_sum(zz);
            if (lane == 0) s_red[2] = zz;
        }
    }
    __syncthreads();
    const double zz = s_red[2];
    if (!isfinite(zz) || !(zz > 0.0)) return false;
    const double sigma_min_bound = sqrt((double)n / zz);
    return sigma_min_bound / s_red[0] > kGateSafety * DBL_EPSILON * (double)n;
}

// One-sided (Hestenes) Jacobi SVD of the n x n matrix held column-major in G; V accumulates the right vectors.
// Then the minimum-norm least-squares solution with LAPACK gelsd's rule: drop s_i <= rcond * s_max, rcond = eps*n.
__device__ int jacobi_lstsq(SolveScratch& S, int n, const double* bvec, double* xout) {
    double* G = S.G;
    double* V = S.V;
    double* cn = S.sc;            // squared column norms at the start of the sweep (for the negligible-column test)
    __shared__ int s_rot;
    __shared__ double s_smax, s_nmax;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) V[e] = ((e / n) == (e % n)) ? 1.0 : 0.0;
    __syncthreads();
    const int m = (n + 1) & ~1;
    const double tol = DBL_EPSILON * sqrt((double)n);
    for (int sweep = 0; sweep < 40; sweep++) {
        for (int c = warp; c < n; c += nwarps) {
            const double* g = G + (size_t)c * n;
            double ss = 0.0;
            for (int i = lane; i < n; i += 32) ss = fma(g[i], g[i], ss);
            ss = warp_sum(ss);
            if (lane == 0) cn[c] = ss;
        }
        if (threadIdx.x == 0) s_rot = 0;
        __syncthreads();
        if (threadIdx.x < 32) {
            double mx = 0.0;
            for (int c = lane; c < n; c += 32) mx = fmax(mx, cn[c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) s_nmax = mx;
        }
        __syncthreads();
        // columns whose norm is far below the truncation cutoff (eps*n*s_max) carry only rounding noise: rotating them
        // never converges in a relative sense and they are dropped from the solution anyway
        // (threshold = 1e-2 of the truncation cutoff in norm: a column that small at the start of a sweep cannot carry a
        // singular direction that survives the truncation; 0.5 of the cutoff was tried and is NOT safe — unconverged
        // columns below it still mix directions above it, which moved the C1 pipeline's Y0 by 2e-6)
        const double negligible = 1e-4 * (DBL_EPSILON * n) * (DBL_EPSILON * n) * s_nmax;
        for (int step = 0; step < m - 1; step++) {
            for (int k = warp; k < m / 2; k += nwarps) {
                int p, q;
                if (k == 0) { p = m - 1; q = step; }
                else { p = (step + k) % (m - 1); q = (step - k + m - 1) % (m - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                if (q >= n) continue;
                if (cn[p] <= negligible || cn[q] <= negligible) continue;     // cn: norms at the start of the sweep
                double* gp = G + (size_t)p * n;
                double* gq = G + (size_t)q * n;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int i = lane; i < n; i += 32) {
                    const double x = gp[i], y = gq[i];
                    al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga);
                }
                al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
                if (fabs(ga) > tol * sqrt(al * be) && ga != 0.0) {
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    for (int i = lane; i < n; i += 32) {
                        const double x = gp[i], y = gq[i];
                        gp[i] = c * x - s * y;
                        gq[i] = s * x + c * y;
                    }
                    double* vp = V + (size_t)p * n;
                    double* vq = V + (size_t)q * n;
                    for (int i = lane; i < n; i += 32) {
                        const double x = vp[i], y = vq[i];
                        vp[i] = c * x - s * y;
                        vq[i] = s * x + c * y;
                    }
                    if (lane == 0) s_rot = 1;
                }
            }
            __syncthreads();
        }
        const int rot = s_rot;
        __syncthreads();
#ifdef BSDE_DEBUG_JACOBI
        if (threadIdx.x == 0 && (!rot || sweep == 39)) printf("jacobi n=%d sweeps=%d\n", n, sweep + 1);
#endif
        if (!rot) break;
    }
    // singular values, projections u_i . b
    double* sig = S.sc;
    double* proj = S.x;
    for (int c = warp; c < n; c += nwarps) {
        const double* g = G + (size_t)c * n;
        double ss = 0.0, pb = 0.0;
        for (int i = lane; i < n; i += 32) { ss = fma(g[i], g[i], ss); pb = fma(g[i], bvec[i], pb); }
        ss = warp_sum(ss); pb = warp_sum(pb);
        if (lane == 0) { sig[c] = sqrt(ss); proj[c] = pb; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double mx = 0.0;
        for (int c = 0; c < n; c++) mx = fmax(mx, sig[c]);
        s_smax = mx;
    }
    __syncthreads();
    const double cutoff = DBL_EPSILON * (double)n * s_smax;
    int rank = 0;
    for (int c = 0; c < n; c++) rank += (sig[c] > cutoff) ? 1 : 0;
    // x = sum_c v_c (g_c . b) / sig_c^2
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < n; c++)
            if (sig[c] > cutoff) acc = fma(V[(size_t)c * n + i], proj[c] / (sig[c] * sig[c]), acc);
        xout[i] = acc;
    }
    __syncthreads();
    return rank;
}

// Symmetric systems (every normal-equation matrix of the ALS / SALSA micro-steps): the truncated minimum-norm solution
// from the eigen-decomposition A = V diag(lam) V^T computed by the classical TWO-sided Jacobi method.
//
// jacobi_lstsq above orthogonalises the columns of A, i.e. it implicitly diagonalises A^T A = A^2.  The systems that
// reach the SVD path are the ill-conditioned ones (monomial features: cond(A) = 1e15 .. 1e40), and on A^2 the relative
// convergence test is hopeless: measured on a B200, 40 sweeps (the cap) and 5.7 ms per 64 x 64 solve in the d = 100
// backward solves, against 35 us for the Cholesky path.  Rotating A itself (A <- J^T A J) converges quadratically in
// 6 - 10 sweeps whatever the grading.
//
// Parallel ordering: a round of the round-robin tournament has m/2 disjoint pairs (p_k, q_k).  The m/2 rotations are
// computed by m/2 threads; then thread (k, l) owns the 2 x 2 block rows {p_k, q_k} x columns {p_l, q_l} and applies
// J_k^T from the left and J_l from the right in registers — the (m/2)^2 blocks tile the matrix, so the whole two-sided
// update of a round is ONE phase between two barriers (1024 blocks = one per thread at n = 64).  Rows p_k, q_k of Vt
// (eigenvectors as rows) are rotated in the same phase.
// Rotation test: |a_pq| > max(eps * sqrt|a_pp a_qq|, floor), floor = 1e-3 eps^2 n max|a_ii|: the relative part keeps the
// accuracy of small eigenvalues of graded positive definite matrices (Demmel-Veselic), the absolute floor — a factor
// 1e-3 eps below the truncation cutoff eps n lam_max — makes the iteration terminate on blocks that hold nothing but
// rounding noise.  Singular values of a symmetric matrix are |lam_i|; cutoff and minimum-norm sum as in jacobi_lstsq.
__device__ __forceinline__ void rr_pair(int k, int step, int m, int& p, int& q) {
    if (k == 0) { p = m - 1; q = step; }
    else { p = (step + k) % (m - 1); q = (step - k + m - 1) % (m - 1); }
    if (p > q) { const int t = p; p = q; q = t; }
}

__device__ int jacobi_eig_lstsq(SolveScratch& S, int n, const double* bvec, double* xout, int* sweeps_out) {
    double* A = S.G;              // n x n, ld = n, symmetric
    double* Vt = S.V;             // row i = eigenvector i
    double* rc = S.sc;            // cosines of the round's rotations (m/2 <= n entries)
    double* rs = S.red;           // sines
    int* pidx = reinterpret_cast<int*>(S.x);      // the round's pairs (p, q; q = -1: padding of an odd n)
    __shared__ int s_rot;
    __shared__ double s_dmax, s_smax;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
    for (int e = tid; e < n * n; e += nthreads) Vt[e] = ((e / n) == (e % n)) ? 1.0 : 0.0;
    const int m = (n + 1) & ~1, hp = m >> 1;
    int* qidx = pidx + hp;
    int sweep = 0;
    for (; sweep < 30; sweep++) {
        if (tid < 32) {
            double mx = 0.0;
            for (int i = lane; i < n; i += 32) mx = fmax(mx, fabs(A[i * n + i]));
            mx = warp_max(mx);
            if (lane == 0) { s_dmax = mx; s_rot = 0; }
        }
        __syncthreads();
        const double floor_abs = 1e-3 * DBL_EPSILON * DBL_EPSILON * (double)n * s_dmax;
        for (int step = 0; step < m - 1; step++) {
            // ---- the round's rotations: one thread per pair (reciprocal / rsqrt based, ~60 dependent FP64 operations) ----
            if (tid < hp) {
                int p, q;
                rr_pair(tid, step, m, p, q);
                double c = 1.0, s = 0.0;
                if (q < n) {
                    const double app = A[p * n + p], aqq = A[q * n + q], apq = A[p * n + q];
                    const double prod = fabs(app * aqq);
                    if (fabs(apq) > fmax(DBL_EPSILON * (prod * rsqrt(fmax(prod, DBL_MIN))), floor_abs)) {
                        const double zeta = (aqq - app) * (0.5 * fast_rcp(apq));
                        const double az = fabs(zeta);
                        if (az < 1e150) {
                            const double h = fma(zeta, zeta, 1.0);
                            const double t = copysign(fast_rcp(az + h * rsqrt(h)), zeta);
                            c = rsqrt(fma(t, t, 1.0));
                            s = c * t;
                        }
                        if (s != 0.0) s_rot = 1;
                    }
                }
                rc[tid] = c; rs[tid] = s;
                pidx[tid] = p; qidx[tid] = q < n ? q : -1;
            }
            __syncthreads();
            // ---- thread (k, l): block rows {p_k, q_k} x columns {p_l, q_l}  <-  J_k^T (block) J_l ----
            for (int k = warp; k < hp; k += nwarps) {
                const double ck = rc[k], sk = rs[k];
                const int rp = pidx[k] * n, rq = qidx[k] * n;
                const bool hasq = rq >= 0;
                for (int l = lane; l < hp; l += 32) {
                    const double cl = rc[l], sl = rs[l];
                    if (sk == 0.0 && sl == 0.0) continue;
                    const int cp = pidx[l], cq = qidx[l];
                    const bool hascq = cq >= 0;
                    const double a00 = A[rp + cp];
                    const double a01 = hascq ? A[rp + cq] : 0.0;
                    const double a10 = hasq ? A[rq + cp] : 0.0;
                    const double a11 = (hasq && hascq) ? A[rq + cq] : 0.0;
                    const double b00 = fma(ck, a00, -sk * a10), b01 = fma(ck, a01, -sk * a11);
                    const double b10 = fma(sk, a00, ck * a10), b11 = fma(sk, a01, ck * a11);
                    double n00 = fma(cl, b00, -sl * b01), n01 = fma(sl, b00, cl * b01);
                    double n10 = fma(cl, b10, -sl * b11), n11 = fma(sl, b10, cl * b11);
                    if (k == l && sk != 0.0) {                // the annihilated pair: exact zeros, diagonal by the stable update
                        const double t = sk * fast_rcp(ck);
                        n00 = fma(-t, a01, a00); n11 = fma(t, a01, a11); n01 = 0.0; n10 = 0.0;
                    }
                    A[rp + cp] = n00;
                    if (hascq) A[rp + cq] = n01;
                    if (hasq) A[rq + cp] = n10;
                    if (hasq && hascq) A[rq + cq] = n11;
                }
                if (sk != 0.0) {                              // rows p_k, q_k of Vt (a rotated pair always has a real q)
                    for (int j = lane; j < n; j += 32) {
                        const double v0 = Vt[rp + j], v1 = Vt[rq + j];
                        Vt[rp + j] = fma(ck, v0, -sk * v1);
                        Vt[rq + j] = fma(sk, v0, ck * v1);
                    }
                }
            }
            __syncthreads();
        }
        const int rot = s_rot;
        __syncthreads();
        if (!rot) break;
    }
    if (sweeps_out && tid == 0) *sweeps_out = sweep + 1;
    double* lam = S.sc;
    double* proj = S.x;
    for (int c = warp; c < n; c += nwarps) {
        const double* v = Vt + (size_t)c * n;
        double pb = 0.0;
        for (int i = lane; i < n; i += 32) pb = fma(v[i], bvec[i], pb);
        pb = warp_sum(pb);
        if (lane == 0) { lam[c] = A[c * n + c]; proj[c] = pb; }
    }
    __syncthreads();
    if (tid == 0) {
        double mx = 0.0;
        for (int c = 0; c < n; c++) mx = fmax(mx, fabs(lam[c]));
        s_smax = mx;
    }
    __syncthreads();
    const double cutoff = DBL_EPSILON * (double)n * s_smax;
    int rank = 0;
    for (int c = 0; c < n; c++) rank += (fabs(lam[c]) > cutoff) ? 1 : 0;
#ifdef BSDE_DEBUG_JACOBI
    if (tid == 0) {
        double mn = 1e300;
        for (int c = 0; c < n; c++) mn = fmin(mn, fabs(lam[c]));
        printf("two-sided jacobi n=%d sweeps=%d rank=%d log10(cond)=%.1f\n", n, sweep + 1, rank, log10(s_smax / fmax(mn, 1e-300)));
    }
#endif
    for (int i = tid; i < n; i += nthreads) {
        double acc = 0.0;
        for (int c = 0; c < n; c++)
            if (fabs(lam[c]) > cutoff) acc = fma(Vt[(size_t)c * n + i], proj[c] / lam[c], acc);
        xout[i] = acc;
    }
    __syncthreads();
    return rank;
}

// Gaussian elimination with partial pivoting on [A | b] (row-major, ld = n+3 as everywhere in the kernel, b in column n).
__device__ void lu_solve(SolveScratch& S, int n, double* xout) {
    const int ld = n + 3;
    double* A = S.A;
    __shared__ int s_piv;
    for (int i = threadIdx.x; i < n; i += blockDim.x) A[i * ld + n] = S.bvec[i];
    __syncthreads();
    for (int k = 0; k < n; k++) {
        if (threadIdx.x == 0) {
            int best = k; double bv = fabs(A[k * ld + k]);
            for (int i = k + 1; i < n; i++) { const double v = fabs(A[i * ld + k]); if (v > bv) { bv = v; best = i; } }
            s_piv = best;
        }
        __syncthreads();
        const int pr = s_piv;
        if (pr != k)
            for (int j = threadIdx.x; j <= n; j += blockDim.x) { const double t = A[k * ld + j]; A[k * ld + j] = A[pr * ld + j]; A[pr * ld + j] = t; }
        __syncthreads();
        const double inv = 1.0 / A[k * ld + k];
        __syncthreads();
        for (int i = k + 1 + (threadIdx.x >> 4); i < n; i += (blockDim.x >> 4)) {
            const double f = A[i * ld + k] * inv;
            for (int j = k + 1 + (threadIdx.x & 15); j <= n; j += 16) A[i * ld + j] = fma(-f, A[k * ld + j], A[i * ld + j]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int k = n - 1; k >= 0; k--) {
            const double xk = A[k * ld + n] / A[k * ld + k];
            __syncwarp();
            if (lane == 0) xout[k] = xk;
            for (int i = lane; i < k; i += 32) A[i * ld + n] = fma(-A[i * ld + k], xk, A[i * ld + n]);
            __syncwarp();
        }
    }
    __syncthreads();
}

// phase timing (option "profile"): thread 0 accumulates clock64() deltas per phase into a.clk[phase]
#define BSDE_PHASE(idx)                                                                              \
    do {                                                                                             \
        if (a.clk && threadIdx.x == 0) {                                                             \
            const long long _now = clock64();                                                        \
            atomicAdd((unsigned long long*)(a.clk + (idx)), (unsigned long long)(_now - t_phase));   \
            t_phase = _now;                                                                          \
        }                                                                                            \
    } while (0)

// Orthogonalisation step of the sweep (als.py:70-77 / :87-94): QR of the solved core, push of the triangular factor into
// the neighbouring core.  Mq: work matrix (n doubles), W / qrR / tau / vbuf: scratch as laid out by the kernels.
__device__ void orthogonalise_and_push(const MicroSolveArgs& a, int n, const double* x, double* Mq, double* W, double* qrR,
                                       double* tau, double* vbuf, long long& t_phase) {
    // ---- orthogonalisation step of the sweep (als.py:70-77 / :87-94) ----
    if (a.direction == 0) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) a.core_j[i] = x[i];
        return;
    }
    if (a.direction > 0) {
        const int mrows = a.rl * a.p, k = a.rr;            // left_unfold(X): (rl*p) x rr, row-major == x itself
        for (int i = threadIdx.x; i < n; i += blockDim.x) Mq[i] = x[i];
        __syncthreads();
        if (threadIdx.x < 32) {
            if (mrows <= 64 && k <= 4) warp_householder_qr_regs<4>(Mq, mrows, k, k, qrR);
            else if (mrows <= 64 && k <= 8) warp_householder_qr_regs<8>(Mq, mrows, k, k, qrR);
            else warp_householder_qr(Mq, mrows, k, k, qrR, tau, vbuf);
        }
        __syncthreads();
        BSDE_PHASE(2);
        for (int i = threadIdx.x; i < n; i += blockDim.x) a.core_j[i] = Mq[i];
        // core_{j+1} <- S @ right_unfold(core_{j+1})   (rr x (p*r_nb))
        const int cols = a.p * a.r_nb;
        for (int e = threadIdx.x; e < k * cols; e += blockDim.x) {
            const int r = e / cols, c = e % cols;
            double acc = 0.0;
            for (int t = r; t < k; t++) acc = fma(qrR[r * k + t], a.core_nb[t * cols + c], acc);
            W[e] = acc;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < k * cols; e += blockDim.x) a.core_nb[e] = W[e];
    } else {
        const int mrows = a.p * a.rr, k = a.rl;            // right_unfold(X)^T: (p*rr) x rl
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int r = e / k, c = e % k;                 // Mq[r][c] = X[c][r]
            Mq[e] = x[c * mrows + r];
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            if (mrows <= 64 && k <= 4) warp_householder_qr_regs<4>(Mq, mrows, k, k, qrR);
            else if (mrows <= 64 && k <= 8) warp_householder_qr_regs<8>(Mq, mrows, k, k, qrR);
            else warp_householder_qr(Mq, mrows, k, k, qrR, tau, vbuf);
        }
        __syncthreads();
        BSDE_PHASE(2);
        for (int e = threadIdx.x; e < n; e += blockDim.x) {   // core_j = Q^T
            const int c = e / mrows, r = e % mrows;
            a.core_j[e] = Mq[r * k + c];
        }
        // core_{j-1} <- left_unfold(core_{j-1}) @ S^T     ((r_nb*p) x rl)
        const int rows = a.r_nb * a.p;
        for (int e = threadIdx.x; e < rows * k; e += blockDim.x) {
            const int r = e / k, c = e % k;
            double acc = 0.0;
            for (int t = c; t < k; t++) acc = fma(a.core_nb[r * k + t], qrR[c * k + t], acc);
            W[e] = acc;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < rows * k; e += blockDim.x) a.core_nb[e] = W[e];
    }
    BSDE_PHASE(3);
}

// =====================================================================================================================
// Fast path of the ALS micro-step for n <= 64 unknowns: k_micro_solve_fast (256 threads, matrix in registers)
// =====================================================================================================================
// k_micro_solve runs 1024 threads (the Jacobi fallback wants one warp per column pair), i.e. 64 registers per thread, and
// its Cholesky moves every trailing entry through the shared memory of the one SM once per column pair: ~1400 cycles per
// pair at n = 64 although the dependent FP64 chain is short (DFMA latency is 8 cycles, bench/micro/fp64_latency.cu).
// Here the augmented lower triangle (rows n, n+1 = right-hand sides b and the +-1 probe, see cholesky_solve) lives in
// REGISTERS for the whole factorisation: thread (ty, tx) of a 16 x 16 grid owns the 2 x 2 blocks (R, C), R = ty + 16 a,
// C = tx + 16 b.  Per column pair t only the two pivot columns are published through (ping-pong) shared memory and the
// owner of the diagonal block publishes the reciprocal pivots; one block barrier per pair.  The factor is written out as
// the UNIT lower-triangular L~ of A = L~ D L~^T (row-major, Lr[i][k] = L_ik / L_kk), the right-hand-side rows as
// D^-1 L~^-1 rhs, so the back substitution is one shuffle + one FMA per step on register-resident accumulators.
// Same acceptance test as cholesky_solve; when it fails (or a pivot is not positive) the kernel only sets *fast_flag = 1
// and k_micro_solve, launched right behind it, redoes the micro-step on its SVD path.  *fast_flag = 0: done.
constexpr int kFastThreads = 256;
constexpr int kFastColPitch = 72;

struct FastTile {
    // Shared-memory byte addresses of this thread's rows / columns in the published pivot columns (parity 0;
    // + kFastParityBytes for parity 1) and in the factor, made opaque to the compiler (it otherwise rebuilds them from
    // threadIdx in every iteration: 450 instructions per column pair for 74 FP64 operations) and used through explicit
    // ld.shared / st.shared (generic accesses cost a far longer round trip).  Slots without a block point at dummies.
    unsigned crow[3];               // &cu[2 R]      (cw = + kFastColPitch doubles)
    unsigned ccol[2];               // &cu[2 C]
    unsigned lrow[3][2];            // row 2 R + r of Lr
    unsigned piv;                   // pivbuf
    unsigned dinv;                  // reciprocal pivots 1 / D_k, kept for the second-tier acceptance test
    int tx, ty;
};
constexpr unsigned kFastParityBytes = 2 * kFastColPitch * 8;     // between the two ping-pong column buffers
constexpr unsigned kFastCwBytes = kFastColPitch * 8;            // from column 2t to column 2t + 1

__device__ __forceinline__ unsigned smem_addr(const void* p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ double2 lds2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts2(unsigned addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts1(unsigned addr, double x) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(x) : "memory");
}

// Publishes column pair t of this thread's row blocks (column block TB) into the ping-pong buffers of parity t & 1;
// the owner of the diagonal block adds the reciprocal pivots.  Called by the threads with tx == t % 16 only.
template <int TB>
__device__ __forceinline__ void fast_publish(const FastTile& ft, const double (&e)[3][2][2][2], int t) {
    const unsigned par = (t & 1) * kFastParityBytes;
#pragma unroll
    for (int ai = TB; ai < 3; ai++) {             // row block 0 is finished when TB = 1
        sts2(ft.crow[ai] + par, e[ai][TB][0][0], e[ai][TB][1][0]);
        sts2(ft.crow[ai] + par + kFastCwBytes, e[ai][TB][0][1], e[ai][TB][1][1]);
    }
    if (ft.ty == (t & 15)) {
        const unsigned pv = ft.piv + (t & 1) * 32;
        const double p0 = e[TB][TB][0][0], a10 = e[TB][TB][1][0], w11 = e[TB][TB][1][1];
#ifdef BSDE_FAST_SERIAL_PIVOTS
        const double i0 = fast_rcp(p0);
        const double l10 = a10 * i0;
        const double p1 = fma(-l10, a10, w11);
        const double i1 = fast_rcp(p1);
#else
        // This thread's chain — pivot block -> two reciprocals -> shared memory — is the critical path of the whole
        // factorisation (dependent FP64 operations cost ~40 cycles each on an otherwise idle SM).  The second reciprocal
        // does not wait for the first: 1 / p1 = p0 / det with det = p0 w11 - a10^2 (same cancellation, hence the same
        // relative accuracy, as p1 = w11 - a10^2 / p0), and both reciprocals take one third-order correction step
        // (approximation error 2^-23 cubed) instead of two Newton steps: 8 dependent operations instead of 13.
        const double det = fma(p0, w11, -(a10 * a10));
        const double i0 = fast_rcp3(p0);
        const double idet = fast_rcp3(det);
        const double l10 = a10 * i0;
        const double i1 = p0 * idet;
        const double p1 = det;                   // sign of the second pivot (p0 > 0 is checked with it)
#endif
        sts2(pv, i0, a10);
        sts2(pv + 16, i1, (p0 > 0.0 && p1 > 0.0) ? 1.0 : 0.0);
        sts2(ft.dinv + 16 * t, i0, i1);
        sts1(ft.lrow[TB][1] + 16 * t, l10);
    }
}

// One column pair t of k_micro_solve_fast; its columns and pivots were published before the barrier at the top.
// TB = t / 16 = the column block the pair lives in, TBN = that of pair t + 1 (compile time, so that the register tile is
// never indexed dynamically).  The threads that own column pair t + 1 update those blocks FIRST and publish them (the
// owner of the next diagonal block also its reciprocal pivots), so that the dependent chain from one pivot to the next
// — update, two reciprocals, shared-memory round trip — overlaps with the bulk of the trailing update.
// Finished rows / columns and the blocks above the diagonal are updated like everything else — their registers are dead
// and the factor entries they write lie above the diagonal, which the back substitution never reads — so the body is
// branch-free apart from the publishing threads.  Returns true when a pivot is not positive (uniform over the CTA).
template <int TB, int TBN>
__device__ __forceinline__ bool fast_pair_step(const FastTile& ft, double (&e)[3][2][2][2], int t, bool last) {
    const unsigned par = (t & 1) * kFastParityBytes;
    const unsigned pv = ft.piv + (t & 1) * 32;
    const int t15 = t & 15;
    constexpr int A0 = TB, B0 = TB;              // row block 0 / column block 0 are finished when TB = 1
    __syncthreads();
    const double2 pa = lds2(pv), pb = lds2(pv + 16);
    if (pb.y == 0.0) return true;
    const double i0 = pa.x, a10 = pa.y, i1 = pb.x;
    double uj[2][2], vj[2][2], li0[3][2], li1[3][2];
#pragma unroll
    for (int bi = B0; bi < 2; bi++) {
        const double2 u = lds2(ft.ccol[bi] + par), w = lds2(ft.ccol[bi] + par + kFastCwBytes);
        uj[bi][0] = u.x; uj[bi][1] = u.y;
        vj[bi][0] = fma(-(u.x * i0), a10, w.x);
        vj[bi][1] = fma(-(u.y * i0), a10, w.y);
    }
#pragma unroll
    for (int ai = A0; ai < 3; ai++) {
        const double2 u = lds2(ft.crow[ai] + par), w = lds2(ft.crow[ai] + par + kFastCwBytes);
        li0[ai][0] = u.x * i0; li0[ai][1] = u.y * i0;
        li1[ai][0] = fma(-li0[ai][0], a10, w.x) * i1;
        li1[ai][1] = fma(-li0[ai][1], a10, w.y) * i1;
    }
    // the next pivot columns first
    const bool next_owner = !last && ft.tx == ((t + 1) & 15);
    if (next_owner) {
#pragma unroll
        for (int ai = TBN; ai < 3; ai++)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++)
                    e[ai][TBN][r][c] = fma(-li1[ai][r], vj[TBN][c], fma(-li0[ai][r], uj[TBN][c], e[ai][TBN][r][c]));
        fast_publish<TBN>(ft, e, t + 1);
    }
    if (ft.tx == t15) {
#pragma unroll
        for (int ai = A0; ai < 3; ai++)
#pragma unroll
            for (int r = 0; r < 2; r++) sts2(ft.lrow[ai][r] + 16 * t, li0[ai][r], li1[ai][r]);
    }
#pragma unroll
    for (int ai = A0; ai < 3; ai++)
#pragma unroll
        for (int bi = B0; bi < 2; bi++) {
            if (ai == 0 && bi == 1) continue;    // row pair < 16 <= column pair: above the diagonal for every thread
            if (bi == TBN && ai >= TBN && next_owner) continue;      // done above
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++)
                    e[ai][bi][r][c] = fma(-li1[ai][r], vj[bi][c], fma(-li0[ai][r], uj[bi][c], e[ai][bi][r][c]));
        }
    return false;
}

// Second-tier acceptance test of k_micro_solve_fast (see the call site): warp 0 estimates sigma_min(A) by three more
// inverse iterations with the factor A_s = L~ D L~^T (A_s = S A S, S = diag(sc)) starting from A^-1 e (pv), and
// sigma_max(A) by six power iterations on A read back from the staged moments.  Lane l carries rows l and l + 32.
// Returns (uniformly) whether sigma_min / sigma_max > kGate2 * eps * n.
constexpr double kGate2 = 1.25;
__device__ bool fast_second_tier(const MicroSolveArgs& a, int n, int np, int ld, const double* __restrict__ Lr,
                                 const double* __restrict__ sc, const double* __restrict__ dinv, double* pv,
                                 const double* __restrict__ stage, const double* __restrict__ dg, double* s_red) {
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned full = 0xffffffffu;
    if (tid < 32) {
        const bool h0 = lane < n, h1 = lane + 32 < n;
        const bool f0 = lane < np, f1 = lane + 32 < np;
        double w0 = h0 ? pv[lane] : 0.0, w1 = h1 ? pv[lane + 32] : 0.0;
        double grow = 0.0;
        for (int it = 0; it < 3; it++) {
            const double inv = rsqrt(warp_sum(fma(w0, w0, w1 * w1)));
            double z0 = h0 ? w0 * inv * sc[lane] : 0.0, z1 = h1 ? w1 * inv * sc[lane + 32] : 0.0;
#pragma unroll 4
            for (int k = 0; k < np; k++) {                       // L~ z = S w / |w|
                const double zk = __shfl_sync(full, (k & 32) ? z1 : z0, k & 31);
                const double l0 = (f0 && lane > k) ? Lr[lane * ld + k] : 0.0;
                const double l1 = (f1 && lane + 32 > k) ? Lr[(lane + 32) * ld + k] : 0.0;
                z0 = fma(-l0, zk, z0);
                z1 = fma(-l1, zk, z1);
            }
            z0 = f0 ? z0 * dinv[lane] : 0.0;
            z1 = f1 ? z1 * dinv[lane + 32] : 0.0;
#pragma unroll 4
            for (int kk = np - 1; kk >= 0; kk--) {               // L~^T y = D^-1 z
                const double yk = __shfl_sync(full, (kk & 32) ? z1 : z0, kk & 31);
                const double c0 = lane < kk ? Lr[kk * ld + lane] : 0.0;
                const double c1 = lane + 32 < kk ? Lr[kk * ld + lane + 32] : 0.0;
                z0 = fma(-c0, yk, z0);
                z1 = fma(-c1, yk, z1);
            }
            w0 = h0 ? z0 * sc[lane] : 0.0;
            w1 = h1 ? z1 * sc[lane + 32] : 0.0;
            grow = sqrt(warp_sum(fma(w0, w0, w1 * w1)));        // |A^-1 w| / |w|  ->  1 / sigma_min from below
        }
        // power iteration, start vector sqrt(diag A) = 1 / sc
        const uint16_t* map = a.expand_map;
        __syncwarp();
        if (h0) pv[lane] = 1.0 / sc[lane];
        if (h1) pv[lane + 32] = 1.0 / sc[lane + 32];
        __syncwarp();
        double rq = 0.0;
        for (int it = 0; it < 6; it++) {
            double u0 = 0.0, u1 = 0.0;
            // (dg: the diagonal with SALSA's regulariser; equal to the staged moment for ALS, then the correction is zero)
            if (h0) { const uint16_t* m0 = map + lane * n; for (int j = 0; j < n; j++) u0 = fma(stage[m0[j]], pv[j], u0); u0 = fma(dg[lane] - stage[m0[lane]], pv[lane], u0); }
            if (h1) { const uint16_t* m1 = map + (lane + 32) * n; for (int j = 0; j < n; j++) u1 = fma(stage[m1[j]], pv[j], u1); u1 = fma(dg[lane + 32] - stage[m1[lane + 32]], pv[lane + 32], u1); }
            const double v0 = h0 ? pv[lane] : 0.0, v1 = h1 ? pv[lane + 32] : 0.0;
            const double vv = warp_sum(fma(v0, v0, v1 * v1)), vu = warp_sum(fma(v0, u0, v1 * u1));
            const double uu = warp_sum(fma(u0, u0, u1 * u1));
            rq = vu / vv;
            const double inv = rsqrt(uu);
            __syncwarp();
            if (h0) pv[lane] = u0 * inv;
            if (h1) pv[lane + 32] = u1 * inv;
            __syncwarp();
        }
        if (lane == 0) {
            const double smin = 1.0 / grow;
            const bool ok = isfinite(grow) && grow > 0.0 && isfinite(rq) && rq > 0.0 &&
                            smin > kGate2 * DBL_EPSILON * (double)n * rq;
            s_red[5] = ok ? 1.0 : 0.0;
            s_red[6] = rq; s_red[7] = smin;
        }
    }
    __syncthreads();
    return s_red[5] != 0.0;
}

// doubles at the front of k_micro_solve_fast's shared memory: the factorisation's buffers or, once it has been rejected,
// vh_lstsq's scratch (whichever is larger; even, so that what follows stays 16-byte aligned)
__host__ __device__ inline size_t fast_front_doubles(int n) {
    const size_t np = (size_t)((n + 1) & ~1);
    const size_t fact = 4 * kFastColPitch + 8 + (np + 3) * (np + 4) + 3 * np;
    const size_t vh = vh_scratch_doubles(n);
    return ((fact > vh ? fact : vh) + 1) & ~(size_t)1;
}

__global__ void __launch_bounds__(kFastThreads, 1)
k_micro_solve_fast(MicroSolveArgs a) {
    extern __shared__ double sm[];
    long long t_phase = a.clk ? clock64() : 0;
    const int n = a.rl * a.p * a.rr;
    const int np = (n + 1) & ~1;                // odd n: padded with an identity row / column
    const int NP = np >> 1;                     // column pairs; row pair NP = the two right-hand-side rows
    const int ld = np + 4;                      // even: the factor is written in 16-byte pairs
    double* colbuf = sm;                        // [2 parities][2 columns][kFastColPitch]   (16-byte aligned)
    double* pivbuf = colbuf + 4 * kFastColPitch;   // [2 parities][4]
    double* Lr = pivbuf + 8;                    // (np + 3) x ld: factor rows, two right-hand-side rows, one dummy row
    double* sc = Lr + (size_t)(np + 3) * ld;    // np
    double* dinv = sc + np;                     // np: 1 / D_k
    double* pv = dinv + np;                     // np: power-iteration vector (second-tier test)
    // (everything above is dead when the factorisation is rejected: vh_lstsq's scratch overlays it from sm on)
    double* xs = sm + fast_front_doubles(n);    // np: the solution
    double* dg = xs + np;                       // np: diagonal of the system (moments + regulariser)
    double* bv = dg + np;                       // np: right-hand side (truncated path)
    double* W = bv + np;                        // kMaxR*kMaxP*kMaxR
    double* qrR = W + kMaxR * kMaxP * kMaxR;    // kMaxR^2
    double* tau = qrR + kMaxR * kMaxR;          // kMaxR
    double* vbuf = tau + kMaxR;                 // kMaxR*kMaxP
    double* stage = vbuf + kMaxR * kMaxP + 16;  // a.n_slots
    __shared__ double s_red[8];
    // Thread (ty, tx) of the 16 x 16 grid.  BSDE_FAST_COLMAJOR (default): tx = tid / 16, so the 16 threads that publish a
    // pivot column pair (tx == t % 16) are ONE half-warp — the other seven warps never enter the publishing branch and start
    // their share of the trailing update at once (with tx = tid % 16 every warp held two of them and walked through it).
#ifndef BSDE_FAST_COLMAJOR
#define BSDE_FAST_COLMAJOR 1
#endif
#if BSDE_FAST_COLMAJOR
    const int tid = threadIdx.x, ty = tid & 15, tx = tid >> 4;
#else
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
#endif

    // ---- the map entries of this thread's blocks (global loads issued before the moments are staged) ----
    const uint16_t* map = a.expand_map;
    int Rw[3], Cw[2];
    bool rowok[3], colok[2];
#pragma unroll
    for (int ai = 0; ai < 3; ai++) { Rw[ai] = ty + 16 * ai; rowok[ai] = Rw[ai] <= NP; }
#pragma unroll
    for (int bi = 0; bi < 2; bi++) { Cw[bi] = tx + 16 * bi; colok[bi] = Cw[bi] < NP; }
    int midx[3][2][2][2];
#pragma unroll
    for (int ai = 0; ai < 3; ai++)
#pragma unroll
        for (int bi = 0; bi < 2; bi++)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const int j = 2 * Cw[bi] + c;
                    int m = -1;                                     // -1: zero, -2: one (padding diagonal), -3: probe
                    if (rowok[ai] && colok[bi] && j < n) {
                        if (Rw[ai] < NP) {
                            const int i = 2 * Rw[ai] + r;
#if BSDE_FAST_COLMAJOR          // lanes run over the rows: read the mirror entry (the map is symmetric), 2-byte stride across the warp
                            if (i < n) { if (j <= i) m = map[j * n + i]; }
#else
                            if (i < n) { if (j <= i) m = map[i * n + j]; }
#endif
                        } else {
                            m = r == 0 ? (int)map[n * n + j] : -3;
                        }
                    } else if (rowok[ai] && colok[bi] && Rw[ai] < NP && 2 * Rw[ai] + r == j) {
                        m = -2;
                    }
                    midx[ai][bi][r][c] = m;
                }
    const int hint0 = a.hint ? *a.hint : 0;                    // (read with the other global loads, used after the staging)
    const int dmap = tid < n ? (int)map[tid * n + tid] : -1;
    const int bmap = tid < n ? (int)map[n * n + tid] : 0;      // (loaded here: a global round trip in front of the factorisation otherwise)
    // Everything read so far is older than the previous micro-step (the plan's expansion map never changes, the hint word was
    // written by this core's previous solve): under programmatic dependent launch those loads overlap the reduction.
    pdl_wait();
    pdl_release();
    // (four slots per thread and trip, every load issued before the first addition: the 897 slots of the C3 shape are one
    // round trip to L2 instead of the three of the compiler's unroll-by-two; same additions in the same order)
    for (int e0 = tid; e0 < a.n_slots; e0 += 4 * kFastThreads) {
        double t[4][kRedGroups];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u * kFastThreads;
#pragma unroll
            for (int g = 0; g < kRedGroups; g++)
                t[u][g] = (e < a.n_slots && g < a.moment_rows) ? a.moments[(size_t)g * a.n_slots + e] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u * kFastThreads;
            double acc = t[u][0];
#pragma unroll
            for (int g = 1; g < kRedGroups; g++) acc += t[u][g];
            if (e < a.n_slots) stage[e] = acc;
        }
    }
    __syncthreads();
    // ---- equilibration, trace, positivity of the diagonal ----
    if (tid < np) dg[tid] = tid < n ? regularised_diagonal(a, tid, stage[dmap]) : 0.0;
    __syncthreads();
    {   // normalise: largest diagonal entry into [1, 2) (exact; the solution does not change)
        double mx = 0.0;
        for (int i = tid & 31; i < n; i += 32) mx = fmax(mx, dg[i]);
        mx = warp_max(mx);
        const double nrm = pow2_normaliser(mx);
        __syncthreads();
        for (int e = tid; e < a.n_slots; e += kFastThreads) stage[e] *= nrm;
        if (tid < np) {
            const double d = tid < n ? dg[tid] * nrm : 1.0;
            sc[tid] = rsqrt(d);
            dg[tid] = d;
        }
    }
    __syncthreads();
    if (tid < n) bv[tid] = stage[bmap];
    if (tid < 32) {
        double tr = 0.0;
        int bad = 0;
        for (int i = tid; i < n; i += 32) {
            const double d = dg[i];
            if (!(d > 0.0) || !isfinite(d)) bad = 1;
            tr += d;
        }
        tr = warp_sum(tr);
        bad = __any_sync(0xffffffffu, bad);
        if (tid == 0) { s_red[0] = tr; s_red[1] = bad ? 1.0 : 0.0; }
    }
    double e[3][2][2][2];
#pragma unroll
    for (int ai = 0; ai < 3; ai++)
#pragma unroll
        for (int bi = 0; bi < 2; bi++)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const int m = midx[ai][bi][r][c];
                    const int j = min(2 * Cw[bi] + c, np - 1);
                    double v = 0.0;
                    if (m >= 0) v = (Rw[ai] < NP && 2 * Rw[ai] + r == 2 * Cw[bi] + c) ? dg[j] : stage[m];
                    else if (m == -2) v = 1.0;
                    else if (m == -3) v = (((unsigned)j * 2654435761u) >> 7) & 1u ? 1.0 : -1.0;
                    v *= sc[j];
                    if (Rw[ai] < NP) v *= sc[min(2 * Rw[ai] + r, np - 1)];
                    e[ai][bi][r][c] = v;
                }
    __syncthreads();
    // accepted = the exact LDL^T solve is numpy.linalg.lstsq's answer.  It stops being true on a non-positive diagonal
    // entry or pivot and when the acceptance tests say that lstsq would truncate; the micro-step is then redone by
    // vh_lstsq below, in this kernel (round 1 launched k_micro_solve behind every fast kernel for that: a 3 us no-op
    // launch per micro-step, and a 1.3 ms two-sided Jacobi when it was needed).
    bool accepted = s_red[1] == 0.0;
    // *a.hint != 0: the last solve of this core truncated (set below; cleared at the start of every fit).  The systems of
    // a core stay rank-deficient from sweep to sweep once the fit has converged, so the factorisation and its acceptance
    // tests (~35k cycles) are not tried again — vh_lstsq returns lstsq's answer whatever the conditioning, and clears the
    // hint again when it finds the system comfortably regular.  (uniform: one word read by every thread)
    if (hint0 != 0) accepted = false;
    BSDE_PHASE(0);
    long long t_sub = t_phase;
    if (a.stop_after == 1) return;
    if (accepted) {

    // ---- factorisation: column pairs 0..15 (column block 0), then 16.. (column block 1: row block 0 and column
    // block 0 are finished by then, so that half touches a third of the registers) ----
    FastTile ft;
    // dummy targets of slots without a block: entries 68.. of the published columns, row np + 2 of the factor
#pragma unroll
    for (int ai = 0; ai < 3; ai++) {
        ft.crow[ai] = smem_addr(colbuf + (rowok[ai] ? 2 * Rw[ai] : 68));
#pragma unroll
        for (int r = 0; r < 2; r++) ft.lrow[ai][r] = smem_addr(Lr + (size_t)(rowok[ai] ? 2 * Rw[ai] + r : np + 2) * ld);
    }
#pragma unroll
    for (int bi = 0; bi < 2; bi++) ft.ccol[bi] = smem_addr(colbuf + (colok[bi] ? 2 * Cw[bi] : 70));
    ft.piv = smem_addr(pivbuf);
    ft.dinv = smem_addr(dinv);
    ft.tx = tx; ft.ty = ty;
    bool failed = false;
#ifdef BSDE_FAST_TWICE      // measurement only: the factorisation runs twice, the second pass (warm instruction cache) is the one timed
    double e_keep[3][2][2][2];
#pragma unroll
    for (int ai = 0; ai < 3; ai++)
#pragma unroll
        for (int bi = 0; bi < 2; bi++)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 2; c++) e_keep[ai][bi][r][c] = e[ai][bi][r][c];
    for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) {
        __syncthreads();
#pragma unroll
        for (int ai = 0; ai < 3; ai++)
#pragma unroll
            for (int bi = 0; bi < 2; bi++)
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int c = 0; c < 2; c++) e[ai][bi][r][c] = e_keep[ai][bi][r][c];
        if (a.clk && tid == 0) t_sub = clock64();
    }
#endif
    if (tx == 0) fast_publish<0>(ft, e, 0);
    for (int t = 0; t < NP && t < 15 && !failed; t++) failed = fast_pair_step<0, 0>(ft, e, t, t == NP - 1);
    if (NP > 15 && !failed) failed = fast_pair_step<0, 1>(ft, e, 15, NP == 16);
    for (int t = 16; t < NP && !failed; t++) failed = fast_pair_step<1, 1>(ft, e, t, t == NP - 1);
#ifdef BSDE_FAST_TWICE
    }
#endif
#ifdef BSDE_DEBUG_JACOBI
    if (failed && tid == 0) printf("fast cholesky pivot failed n=%d\n", n);
#endif
    if (failed) accepted = false;                                // uniform
    }
    if (a.stop_after == 2) return;
    __syncthreads();
    if (a.clk && tid == 0) { const long long now = clock64(); atomicAdd((unsigned long long*)(a.clk + 5), (unsigned long long)(now - t_sub)); t_sub = now; }
    if (accepted) {

    // ---- back substitution L~^T x = D^-1 L~^-1 rhs: warp 0 -> b, warp 1 -> probe; lane l carries rows l and l + 32 ----
    if (tid < 64) {
        const int lane = tid & 31;
        const double* zrow = Lr + (np + (tid >> 5)) * ld;
        double acc0 = lane < np ? zrow[lane] : 0.0, acc1 = lane + 32 < np ? zrow[lane + 32] : 0.0;
        double r0 = 0.0, r1 = 0.0;                // row kk of L~ (columns lane, lane + 32), fetched one step ahead
        {
            const double* row = Lr + (np - 1) * ld;
            r0 = lane < np - 1 ? row[lane] : 0.0;
            r1 = lane + 32 < np - 1 ? row[lane + 32] : 0.0;
        }
        for (int kk = np - 1; kk >= 0; kk--) {
            const double c0 = r0, c1 = r1;
            if (kk > 0) {
                const double* row = Lr + (kk - 1) * ld;
                r0 = lane < kk - 1 ? row[lane] : 0.0;
                r1 = lane + 32 < kk - 1 ? row[lane + 32] : 0.0;
            }
            const double xk = __shfl_sync(0xffffffffu, (kk & 32) ? acc1 : acc0, kk & 31);
            acc0 = fma(-c0, xk, acc0);            // c = 0 for columns >= kk: finished rows keep their value
            acc1 = fma(-c1, xk, acc1);
        }
        const double v0 = lane < n ? acc0 * sc[lane] : 0.0, v1 = lane + 32 < n ? acc1 * sc[lane + 32] : 0.0;
        if (tid < 32) {
            if (lane < n) xs[lane] = v0;
            if (lane + 32 < n) xs[lane + 32] = v1;
        } else {
            const double zz = warp_sum(fma(v0, v0, v1 * v1));
            if (lane == 0) s_red[2] = zz;
            if (lane < np) pv[lane] = v0;                        // A^-1 e, the start of the inverse iteration below
            if (lane + 32 < np) pv[lane + 32] = v1;
        }
    }
    __syncthreads();
    if (a.clk && tid == 0) { const long long now = clock64(); atomicAdd((unsigned long long*)(a.clk + 6), (unsigned long long)(now - t_sub)); }
    {
        const double zz = s_red[2];
        bool ok = isfinite(zz) && zz > 0.0 && sqrt((double)n / zz) / s_red[0] > kGateSafety * DBL_EPSILON * (double)n;
        // Hopeless systems skip the second tier (~45k cycles on one warp): cond >= (trace / n) / sigma_min, so when the
        // bound on sigma_min is below eps / 4 of the trace lstsq truncates whatever the better estimates say (its threshold
        // is cond = 1 / (eps n)) — the systems of a converged fit with excess rank sit at 1e-17 .. 1e-20.
        const bool hopeless = !(sqrt((double)n / zz) / s_red[0] > 0.25 * DBL_EPSILON);
        if (!ok && isfinite(zz) && zz > 0.0 && a.tier2 && !hopeless) {
            // ---- second tier.  The first test is cheap but loose (trace >= sigma_max by up to n, one probe solve, safety
            // 100): at N >= 2^21 samples a tenth of the C3 micro-steps fail it although lstsq truncates nothing (measured:
            // cond 10^12.4 .. 10^13.6 against the truncation threshold 1/(eps n) = 10^13.85), and each of them costs an
            // eigen-decomposition.  Here both ends of the spectrum are estimated properly: sigma_max by six power
            // iterations on A (Rayleigh quotient, a lower bound that is tight for these matrices), sigma_min by three more
            // inverse iterations on the factor we already hold.  The exact solve is accepted when the estimated ratio
            // stays kGate2 = 1.25 above eps n (LAPACK's own sigma_min carries an absolute error of a few eps sigma_max, i.e.
            // a few percent of sigma_min ~ eps n sigma_max: closer to the threshold its decision is noise too).
            ok = fast_second_tier(a, n, np, ld, Lr, sc, dinv, pv, stage, dg, s_red);
        }
#ifdef BSDE_DEBUG_JACOBI
        if (!ok && tid == 0) printf("fast gate failed n=%d log10(trace/sigma_min_bound)=%.1f tier2 log10(cond)=%.1f\n", n, log10(s_red[0] / sqrt((double)n / zz)), log10(s_red[6] / s_red[7]));
#endif
        if (!ok) accepted = false;
    }
    }
    if (a.stop_after == 3) return;
    if (accepted) {
        if (tid == 64 && a.info) { atomicAdd(a.info, 1); atomicMin(a.info + 3, n); }   // not warp 0: it runs the QR next
    } else {
        // ---- truncated minimum-norm solve (vh_lstsq.cuh); its scratch overlays the factorisation's buffers ----
        __syncthreads();
        __shared__ int s_sweeps;
        const SystemMatrix Amat{nullptr, stage, map, has_diagonal_terms(a) ? dg : nullptr, n, 1.0};      // stage is normalised already
        VhScratch vhs = vh_carve(sm, n);
#ifdef VH_PROFILE      // measurement build: phase cycles of the truncated solve into slots 16.. of the clock buffer
        __shared__ long long s_vhclk[16];
        if (tid < 16) s_vhclk[tid] = 0;
        __syncthreads();
        vhs.clk = a.clk ? s_vhclk : nullptr;
        const long long vh_t0 = clock64();
#endif
        int rank = vh_lstsq<kFastThreads, 8, 4>(Amat, n, bv, xs, vhs, a.clk ? &s_sweeps : nullptr);
#ifdef VH_PROFILE
        if (a.clk && tid < 14) atomicAdd((unsigned long long*)(a.clk + 16 + tid), (unsigned long long)s_vhclk[tid]);
        if (a.clk && tid == 0) atomicAdd((unsigned long long*)(a.clk + 31), (unsigned long long)(clock64() - vh_t0));
#endif
        if (rank < 0) {                       // moments that are not a Gram matrix (overflow, NaN): no answer
            for (int i = tid; i < n; i += kFastThreads) xs[i] = nan("");
            rank = 0;
            __syncthreads();
        }
        if (tid == 0) {
            if (a.info) { atomicAdd(a.info + 1, 1); atomicMin(a.info + 3, rank); }
            if (a.clk) atomicAdd((unsigned long long*)(a.clk + 7), (unsigned long long)s_sweeps);
            // full rank with the margin over lstsq's cutoff that the second-tier test asks for (1.25; 1.5 here): the exact
            // solve would have been accepted, so the next solve of this core tries the factorisation again.  (Measured on
            // the 51 fits of the full solve: with a margin of 1000 the hint kept 80 regular systems per fit on the
            // truncated path, 6.75 s against 6.29 s without hints.)
            if (a.hint) *a.hint = (rank == n && vhs.red[4] > 1.5) ? 0 : 1;
        }
    }
    BSDE_PHASE(1);
    orthogonalise_and_push(a, n, xs, sm /* the factorisation's storage is free now */, W, qrR, tau, vbuf, t_phase);
}

// The eigen-decomposition / SVD fallbacks below keep a second n x n matrix (the vectors) in shared memory: up to this n
constexpr int kJacobiMaxN = 100;

// doubles of the first shared-memory region of k_micro_solve: the augmented system (+ the Jacobi vectors for small n),
// overlaid by vh_lstsq's scratch once the Cholesky path has been rejected
__host__ __device__ inline size_t dense_front_doubles(int n) {
    const size_t chol = (size_t)(n + 2) * (n + 3) + (n <= kJacobiMaxN ? (size_t)n * n : 0);
    const size_t vh = vh_scratch_doubles(n);
    return ((chol > vh ? chol : vh) + 1) & ~(size_t)1;
}

__global__ void __launch_bounds__(kDenseThreads)
k_micro_solve(MicroSolveArgs a) {
    extern __shared__ double sm[];
    pdl_wait();
    pdl_release();
    long long t_phase = a.clk ? clock64() : 0;
    const int n = a.rl * a.p * a.rr;
    const int ld = n + 3;
    // Cores whose system does not fit the shared memory of the CTA (n > ~140 unknowns) keep the matrix region and the
    // staged moments in the global workspace a.big_ws (L2-resident, a few MB): the same code on generic pointers — slower
    // per entry, but every core shape the assembly kernels accept can be solved.  The small vectors stay in shared memory.
    double* const front = a.big_ws ? a.big_ws : sm;
    double* const tail = a.big_ws ? sm : sm + dense_front_doubles(n);
    SolveScratch S;
    S.A = front;                                // (n+2) x (n+3): augmented normal equations / Jacobi G / QR work matrix
    S.G = front;
    S.V = front + (size_t)(n + 2) * (n + 3);    // n^2     : Jacobi vectors (n <= kJacobiMaxN only)
    S.x = tail;                                 // n
    S.sc = S.x + n;                             // n
    S.bvec = S.sc + n;                          // n
    S.red = S.bvec + n;                         // n       : Jacobi solution scratch
    double* W = S.red + n;                      // kMaxR*kMaxP*kMaxR : neighbour-core product
    double* qrR = W + kMaxR * kMaxP * kMaxR;    // kMaxR^2
    double* tau = qrR + kMaxR * kMaxR;          // kMaxR
    double* vbuf = tau + kMaxR;                 // kMaxR*kMaxP
    double* stage_buf = a.big_ws ? a.big_ws + dense_front_doubles(n)      // a.n_slots : the assembly kernel's summed slots
                                 : vbuf + kMaxR * kMaxP + 16;
    __shared__ int s_path, s_rank, s_ok, s_sym, s_sweeps;
    __shared__ double s_red[4], s_wmax[32], s_wdif[32];

    expand_normal_equations(a, S.A, ld, S.bvec, stage_buf);
    __syncthreads();
    BSDE_PHASE(0);
    if (a.dbg_A) {
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) a.dbg_A[e] = S.A[(e / n) * ld + (e % n)];
        for (int e = threadIdx.x; e < n; e += blockDim.x) a.dbg_A[n * n + e] = S.bvec[e];
    }
    __syncthreads();
    // normalise the system (largest diagonal entry into [1, 2), exact): see pow2_normaliser
    __shared__ double s_scale;
    if (threadIdx.x < 32) {
        double mx = 0.0;
        for (int i = threadIdx.x; i < n; i += 32) mx = fmax(mx, S.A[i * ld + i]);
        mx = warp_max(mx);
        if (threadIdx.x == 0) s_scale = pow2_normaliser(mx);
    }
    __syncthreads();
    const double nrm = s_scale;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) S.A[(e / n) * ld + (e % n)] *= nrm;
    for (int e = threadIdx.x; e < n; e += blockDim.x) S.bvec[e] *= nrm;
    __syncthreads();
    if (a.solver == SOLVER_LSTSQ) {
        {
            const bool ok = cholesky_solve(S, n, S.x, s_red);
            if (threadIdx.x == 0) { s_ok = ok ? 1 : 0; s_path = ok ? 0 : 1; s_rank = n; s_sym = 1; }
        }
        __syncthreads();
        if (!s_ok) {
            // Moment matrices are symmetric by construction.  An explicit system (bsde_solve) is tested: symmetric ones
            // take the same path, anything else the general one-sided SVD.  Tolerant of a rounded host matrix.
            if (a.A_direct) {
                double amax = 0.0, dmax = 0.0;
                for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                    const int i = e / n, j = e - i * n;
                    if (j > i) continue;
                    const double u = a.A_direct[i * n + j], v = a.A_direct[j * n + i];
                    amax = fmax(amax, fmax(fabs(u), fabs(v)));
                    dmax = fmax(dmax, fabs(u - v));
                }
                amax = warp_max(amax); dmax = warp_max(dmax);
                if ((threadIdx.x & 31) == 0) { s_wmax[threadIdx.x >> 5] = amax; s_wdif[threadIdx.x >> 5] = dmax; }
                __syncthreads();
                if (threadIdx.x == 0) {
                    double a2 = 0.0, d2 = 0.0;
                    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a2 = fmax(a2, s_wmax[w]); d2 = fmax(d2, s_wdif[w]); }
                    s_sym = (d2 <= 8.0 * DBL_EPSILON * a2) ? 1 : 0;
                }
                __syncthreads();
            }
            int rank = -1;
            double* xtmp = S.red;
            if (s_sym && !a.force_onesided) {
                // positive semi-definite systems (every normal-equation matrix): pivoted Cholesky + one-sided Jacobi on the
                // factor (vh_lstsq.cuh).  Its scratch overlays S.A, so the entries are read through the expansion again;
                // the regularised diagonal is parked in S.sc.
                const bool diag = !a.A_direct && has_diagonal_terms(a);
                if (diag)      // (S.A was scaled and factorised in place by the rejected Cholesky attempt)
                    for (int i = threadIdx.x; i < n; i += blockDim.x) S.sc[i] = regularised_diagonal(a, i, stage_buf[a.expand_map[i * n + i]]) * nrm;
                __syncthreads();
                const SystemMatrix Amat{a.A_direct, stage_buf, a.expand_map, diag ? S.sc : nullptr, n, nrm};
                // up to 128 unknowns: 16 lanes per column pair, i.e. 64 pairs at once — a whole round of the odd-even
                // ordering with one column per lane group resident in registers (one pass, one barrier per round);
                // beyond: 32 lanes per pair and the round-robin tournament in two passes per round
                // (bench/micro/vh_solve.cu times both on cores of 65..128 unknowns)
                // (UMAX = 16-byte row chunks per lane: 2 * 32 * UMAX rows)
                if (n <= 128) rank = vh_lstsq<kDenseThreads, 16, 4>(Amat, n, S.bvec, xtmp, vh_carve(front, n), a.clk ? &s_sweeps : nullptr);
                else if (n <= 192) rank = vh_lstsq<kDenseThreads, 32, 3>(Amat, n, S.bvec, xtmp, vh_carve(front, n), a.clk ? &s_sweeps : nullptr);
                else if (n <= 256) rank = vh_lstsq<kDenseThreads, 32, 4>(Amat, n, S.bvec, xtmp, vh_carve(front, n), a.clk ? &s_sweeps : nullptr);
                else rank = vh_lstsq<kDenseThreads, 32, 8>(Amat, n, S.bvec, xtmp, vh_carve(front, n), a.clk ? &s_sweeps : nullptr);
                if (rank >= 0 && a.clk && threadIdx.x == 0) atomicAdd((unsigned long long*)(a.clk + 7), (unsigned long long)s_sweeps);
            }
            if (rank < 0) {
                // symmetric but indefinite (or A/B option svd_onesided), or not symmetric at all: the round-1 paths, for
                // the sizes whose vectors fit beside the matrix
                if (n > kJacobiMaxN) {
                    for (int i = threadIdx.x; i < n; i += blockDim.x) xtmp[i] = nan("");
                    rank = 0;
                    __syncthreads();
                } else {
                    expand_normal_equations(a, S.G, n, S.bvec, stage_buf);      // ld = n for the Jacobi kernels
                    __syncthreads();
                    for (int e = threadIdx.x; e < n * n; e += blockDim.x) S.G[e] *= nrm;
                    for (int e = threadIdx.x; e < n; e += blockDim.x) S.bvec[e] *= nrm;
                    __syncthreads();
                    if (s_sym && !a.force_onesided) {
                        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {      // exact symmetry: mirror the lower triangle
                            const int i = e / n, j = e - i * n;
                            if (j > i) S.G[i * n + j] = S.G[j * n + i];
                        }
                        __syncthreads();
                        xtmp = S.A + (size_t)n * n;      // S.sc / S.red hold the rotations; G uses n*n of the (n+2) x (n+3) storage
                        rank = jacobi_eig_lstsq(S, n, S.bvec, xtmp, a.clk ? &s_sweeps : nullptr);
                    } else {
                        // jacobi_lstsq orthogonalises the COLUMNS of the matrix it finds as G[c * n + i]; the expansion
                        // wrote A row-major, which is the same thing only for a symmetric A: transpose a general one first
                        for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                            const int i = e / n, j = e - i * n;
                            if (j > i) { const double t = S.G[i * n + j]; S.G[i * n + j] = S.G[j * n + i]; S.G[j * n + i] = t; }
                        }
                        __syncthreads();
                        xtmp = S.red;
                        rank = jacobi_lstsq(S, n, S.bvec, xtmp);
                    }
                }
            }
            for (int i = threadIdx.x; i < n; i += blockDim.x) S.x[i] = xtmp[i];
            if (threadIdx.x == 0) s_rank = rank;
            __syncthreads();
        }
    } else {
        lu_solve(S, n, S.x);
        if (threadIdx.x == 0) { s_path = 2; s_rank = n; }
    }
    __syncthreads();
    BSDE_PHASE(1);
    if (threadIdx.x == 0 && a.info) {          // counters: [0] cholesky, [1] truncated (eigen / SVD), [2] lu, [3] min numerical rank
        atomicAdd(a.info + s_path, 1);
        atomicMin(a.info + 3, s_rank);
    }

    orthogonalise_and_push(a, n, S.x, S.A /* the factorisation storage is free now */, W, qrR, tau, vbuf, t_phase);
}

// Whole-train orthonormalisation by successive QRs (tensor_train.py:34-87), one CTA, cores in global memory.
struct OrthoArgs {
    double* cores;
    const int64_t* off;   // device
    const int* ranks;     // device
    const int* pvec;      // device, d basis sizes — or null: every core has p
    int p, d, start, stop, right;
};

__global__ void __launch_bounds__(kDenseThreads)
k_orthonormalize(OrthoArgs a) {
    extern __shared__ double sm[];
    double* Mq = sm;                               // up to kMaxR*kMaxP*kMaxR
    double* W = Mq + kMaxR * kMaxP * kMaxR;
    double* qrR = W + kMaxR * kMaxP * kMaxR;
    double* tau = qrR + kMaxR * kMaxR;
    double* vbuf = tau + kMaxR;
    if (a.right) {
        for (int k = a.stop; k > a.start; k--) {
            double* cur = a.cores + a.off[k];
            double* prev = a.cores + a.off[k - 1];
            const int rl = a.ranks[k], rr = a.ranks[k + 1], rp = a.ranks[k - 1];
            const int p = a.pvec ? a.pvec[k] : a.p, pprev = a.pvec ? a.pvec[k - 1] : a.p;
            const int mrows = p * rr, kk = rl, n = rl * p * rr;
            for (int e = threadIdx.x; e < n; e += blockDim.x) { const int r = e / kk, c = e % kk; Mq[e] = cur[c * mrows + r]; }
            __syncthreads();
            if (threadIdx.x < 32) warp_householder_qr(Mq, mrows, kk, kk, qrR, tau, vbuf);
            __syncthreads();
            for (int e = threadIdx.x; e < n; e += blockDim.x) { const int c = e / mrows, r = e % mrows; cur[e] = Mq[r * kk + c]; }
            const int rows = rp * pprev;
            for (int e = threadIdx.x; e < rows * kk; e += blockDim.x) {
                const int r = e / kk, c = e % kk;
                double acc = 0.0;
                for (int t = c; t < kk; t++) acc = fma(prev[r * kk + t], qrR[c * kk + t], acc);
                W[e] = acc;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < rows * kk; e += blockDim.x) prev[e] = W[e];
            __syncthreads();
        }
    } else {
        for (int k = a.start; k < a.stop; k++) {
            double* cur = a.cores + a.off[k];
            double* next = a.cores + a.off[k + 1];
            const int rl = a.ranks[k], rr = a.ranks[k + 1], rn = a.ranks[k + 2];
            const int p = a.pvec ? a.pvec[k] : a.p, pnext = a.pvec ? a.pvec[k + 1] : a.p;
            const int mrows = rl * p, kk = rr, n = rl * p * rr;
            for (int e = threadIdx.x; e < n; e += blockDim.x) Mq[e] = cur[e];
            __syncthreads();
            if (threadIdx.x < 32) warp_householder_qr(Mq, mrows, kk, kk, qrR, tau, vbuf);
            __syncthreads();
            for (int e = threadIdx.x; e < n; e += blockDim.x) cur[e] = Mq[e];
            const int cols = pnext * rn;
            for (int e = threadIdx.x; e < kk * cols; e += blockDim.x) {
                const int r = e / cols, c = e % cols;
                double acc = 0.0;
                for (int t = r; t < kk; t++) acc = fma(qrR[r * kk + t], next[t * cols + c], acc);
                W[e] = acc;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < kk * cols; e += blockDim.x) next[e] = W[e];
            __syncthreads();
        }
    }
}

}  // namespace

// The opt-in shared-memory limit is a per-device attribute of a kernel: set it once for every device a context launches on
// (the C ABI allows one context per GPU in one process).
static int set_smem_attributes() {
    static bool done[64] = {};
    int dev = 0;
    BSDE_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && done[dev]) return 0;
    BSDE_CUDA_OK(cudaFuncSetAttribute(k_micro_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    BSDE_CUDA_OK(cudaFuncSetAttribute(k_micro_solve_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return 0;
}

#define BSDE_TRY_RC(expr) do { const int _rc = (expr); if (_rc) return _rc; } while (0)

constexpr size_t kDenseSmemLimit = 220 * 1024;
// doubles of k_micro_solve's shared memory besides the matrix region and the staged moments (keep in sync with its layout)
static size_t dense_tail_doubles_n(int n) {
    return 4 * (size_t)n + kMaxR * kMaxP * kMaxR + kMaxR * kMaxR + kMaxR + kMaxR * kMaxP + 16 + 16;
}

size_t micro_solve_workspace_bytes(int n, int n_slots) {
    const size_t slots = (size_t)(n_slots > 0 ? n_slots : 0);
    const size_t all = (dense_front_doubles(n) + dense_tail_doubles_n(n) + slots) * sizeof(double);
    if (all <= kDenseSmemLimit) return 0;
    return (dense_front_doubles(n) + slots + 16) * sizeof(double);
}

int launch_micro_solve(const MicroSolveArgs& a, cudaStream_t st, int* extra_launches) {
    const int n = a.rl * a.p * a.rr;
    if (n > kMaxN) return fail_msg(2, "micro_solve: n = %d unknowns exceeds the in-CTA solver limit %d", n, kMaxN);
    if (a.direction > 0 && a.rl * a.p < a.rr) return fail_msg(2, "micro_solve: left QR needs r_j*p >= r_{j+1} (%d*%d < %d)", a.rl, a.p, a.rr);
    if (a.direction < 0 && a.p * a.rr < a.rl) return fail_msg(2, "micro_solve: right QR needs p*r_{j+1} >= r_j (%d*%d < %d)", a.p, a.rr, a.rl);
    // keep in sync with the layout at the top of k_micro_solve
    const size_t doubles = dense_front_doubles(n) + dense_tail_doubles_n(n) + (size_t)(a.n_slots > 0 ? a.n_slots : 0);
    size_t bytes = doubles * sizeof(double);
    BSDE_TRY_RC(set_smem_attributes());
    // plain micro-step of at most 64 unknowns (ALS and SALSA): the register-resident kernel, which also holds the
    // truncated path; everything else (larger cores, LU, explicit systems, the parity dump) goes to k_micro_solve
    const bool fast = a.fast_flag && a.solver == SOLVER_LSTSQ && n <= 64 && !a.A_direct && !a.dbg_A && a.n_slots > 0 && !a.force_onesided;
    if (fast) {
        const int np = (n + 1) & ~1;
        const size_t fd = fast_front_doubles(n) + 3 * (size_t)np + kMaxR * kMaxP * kMaxR + kMaxR * kMaxR + kMaxR +
                          kMaxR * kMaxP + 16 + (size_t)a.n_slots + 16;
        if (fd * sizeof(double) > 100 * 1024) return fail_msg(2, "micro_solve: %zu B of shared memory for the fast kernel", fd * sizeof(double));
        if (g_pdl_launch) BSDE_CUDA_OK(launch_chain(k_micro_solve_fast, dim3(1), dim3(kFastThreads), fd * sizeof(double), st, true, a));
        else k_micro_solve_fast<<<1, kFastThreads, fd * sizeof(double), st>>>(a);
        BSDE_CUDA_OK(cudaGetLastError());
    } else {
        if (bytes > kDenseSmemLimit) {      // matrix region and staged moments in the global workspace
            if (!a.big_ws) return fail_msg(2, "micro_solve: n = %d needs %zu B of shared memory and no workspace was provided", n, bytes);
            bytes = dense_tail_doubles_n(n) * sizeof(double);
        } else if (a.big_ws) {
            return fail_msg(2, "micro_solve: workspace given for a core of %d unknowns that fits shared memory", n);
        }
        if (g_pdl_launch) BSDE_CUDA_OK(launch_chain(k_micro_solve, dim3(1), dim3(kDenseThreads), bytes, st, true, a));
        else k_micro_solve<<<1, kDenseThreads, bytes, st>>>(a);
        BSDE_CUDA_OK(cudaGetLastError());
    }
    if (extra_launches) *extra_launches = 0;
    return 0;
}

static int launch_ortho(double* cores, const int64_t* off_dev, const int* ranks_dev, int p, int d, int start, int stop,
                        int right, cudaStream_t st, const int* pvec_dev = nullptr) {
    OrthoArgs a{cores, off_dev, ranks_dev, pvec_dev, p, d, start, stop, right};
    const size_t bytes = sizeof(double) * (2 * kMaxR * kMaxP * kMaxR + kMaxR * kMaxR + kMaxR + kMaxR * kMaxP + 16);
    k_orthonormalize<<<1, kDenseThreads, bytes, st>>>(a);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_orthonormalize_right(double* cores, const int64_t* off_dev, const int* ranks_dev, int p, int d, int start,
                                int stop, cudaStream_t st, const int* pvec_dev) {
    return launch_ortho(cores, off_dev, ranks_dev, p, d, start, stop, 1, st, pvec_dev);
}
int launch_orthonormalize_left(double* cores, const int64_t* off_dev, const int* ranks_dev, int p, int d, int start,
                               int stop, cudaStream_t st, const int* pvec_dev) {
    return launch_ortho(cores, off_dev, ranks_dev, p, d, start, stop, 0, st, pvec_dev);
}

}  // namespace bsde
