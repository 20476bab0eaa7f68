// SALSA's small dense steps, each one warp of one CTA, nothing leaves the device
// (reference: SALSA, core/optimisation/als.py:99-323; quirks kept as written — SURVEY.md App. A.3):
//
//   k_salsa_left   als.py:240-260   SVD of left_unfold(core_{j-1}) = U S Vt;  core_{j-1} <- U,
//                                   core_j <- (S * Vt) @ right_unfold(core_j)      [S scales the COLUMNS of Vt]
//                                   optional +1 rank (rank_adaptation, als.py:185-211):  S <- [S, 0.01 min_sv],
//                                   new column of U / new row of V = all-ones, twice Gram-Schmidt projected and
//                                   normalised;  core_j <- diag(S) @ [Vt @ R ; new row];  gamma = 1 / max(S, min_sv)
//   k_salsa_right  als.py:264-277   SVD of left_unfold(core_j);  core_j <- U S,  core_{j+1} <- Vt @ right_unfold(core_{j+1}),
//                                   theta = 1 / max(S, min_sv)
//   k_salsa_scalars als.py:168-172, 288-298   eta / omega / min_sv updates from the residual
//
// The SVD is a one-sided (Hestenes) Jacobi on the columns of the (r p) x r unfolding, singular values sorted in
// decreasing order as LAPACK returns them.
#include "common.cuh"
#include "kernels.cuh"
#include <cfloat>

namespace bsde {

namespace {

constexpr int kSalsaR = 8;             // SALSA ranks (after growth) are limited to 8, like the assembly kernel
constexpr int kMaxM = kSalsaR * kMaxP; // rows of an unfolding
constexpr int kMaxK = kSalsaR;         // columns

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp.  G (m x k, column-major, ld = m) holds the matrix on entry and U*diag(S) on exit (columns sorted by
// decreasing singular value); V (k x k, column-major) the right singular vectors; S the singular values.
__device__ void warp_jacobi_svd(double* G, double* V, double* S, int m, int k) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < k * k; e += 32) V[e] = ((e / k) == (e % k)) ? 1.0 : 0.0;
    __syncwarp();
    const double tol = DBL_EPSILON;
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotated = 0;
        for (int p = 0; p < k - 1; p++)
            for (int q = p + 1; q < k; q++) {
                double* gp = G + p * m;
                double* gq = G + q * m;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int i = lane; i < m; i += 32) {
                    const double x = gp[i], y = gq[i];
                    al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga);
                }
                al = wsum(al); be = wsum(be); ga = wsum(ga);
                if (ga != 0.0 && fabs(ga) > tol * sqrt(al * be)) {
                    rotated = 1;
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    for (int i = lane; i < m; i += 32) {
                        const double x = gp[i], y = gq[i];
                        gp[i] = c * x - s * y;
                        gq[i] = s * x + c * y;
                    }
                    for (int i = lane; i < k; i += 32) {
                        const double x = V[p * k + i], y = V[q * k + i];
                        V[p * k + i] = c * x - s * y;
                        V[q * k + i] = s * x + c * y;
                    }
                }
                __syncwarp();
            }
        if (!rotated) break;
    }
    for (int c = 0; c < k; c++) {
        double ss = 0.0;
        for (int i = lane; i < m; i += 32) ss = fma(G[c * m + i], G[c * m + i], ss);
        ss = wsum(ss);
        if (lane == 0) S[c] = sqrt(ss);
    }
    __syncwarp();
    // Selection sort, decreasing, STABLE for ties within rounding: SALSA's (S * Vt) @ R quirk makes the result depend
    // on the order of (numerically) equal singular values, and the systematic case — core_0 is exactly orthonormal at
    // the start of every sweep after the first, S = (1, .., 1) — keeps its column order in LAPACK as it does here.
    for (int a = 0; a < k - 1; a++) {
        int best = a;
        for (int b = a + 1; b < k; b++)
            if (S[b] > S[best] * (1.0 + 64.0 * DBL_EPSILON)) best = b;
        if (best != a) {
            for (int i = lane; i < m; i += 32) { const double t = G[a * m + i]; G[a * m + i] = G[best * m + i]; G[best * m + i] = t; }
            for (int i = lane; i < k; i += 32) { const double t = V[a * k + i]; V[a * k + i] = V[best * k + i]; V[best * k + i] = t; }
            __syncwarp();
            if (lane == 0) { const double t = S[a]; S[a] = S[best]; S[best] = t; }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ double inv_sv(double s, double min_sv) {
    // nan_to_num(1 / max(S, min_sv), posinf=1e-6, neginf=1e-6)   (als.py:260,277)
    const double v = 1.0 / fmax(s, min_sv);
    if (isnan(v)) return 0.0;
    if (isinf(v)) return 1e-6;
    return v;
}

struct SalsaLeftArgs {
    double* core_prev;     // (rp, p, k)   in; (rp, p, k') out
    double* core_cur;      // (k, p, rn)   in; (k', p, rn) out
    int rp, p, k, rn;
    const double* min_sv;  // device scalar
    int expand;            // host decision: expand_rank && k < max_rank of this bond
    int do_reg;
    double* gamma;         // k' values out (or untouched when !do_reg)
    int* rank_out;         // new bond rank k'
    int* adapted_out;      // set to 1 when the rank grew
};

__global__ void __launch_bounds__(32) k_salsa_left(SalsaLeftArgs a) {
    __shared__ double G[kMaxM * kMaxK], V[kMaxK * kMaxK], S[kMaxK + 1];
    __shared__ double Rm[kMaxK * kMaxP * kSalsaR];    // right_unfold(core_cur): k x cols
    __shared__ double W[(kMaxK + 1) * kMaxP * kSalsaR];
    __shared__ double comp[kMaxM], proj[kMaxK + 1];
    const int lane = threadIdx.x;
    const int m = a.rp * a.p, k = a.k, cols = a.p * a.rn;
    const double min_sv = *a.min_sv;
    for (int e = lane; e < m * k; e += 32) G[(e % k) * m + (e / k)] = a.core_prev[e];     // row-major (m,k) -> column-major
    for (int e = lane; e < k * cols; e += 32) Rm[e] = a.core_cur[e];
    __syncwarp();
    warp_jacobi_svd(G, V, S, m, k);
    // U = G / S (columns); Vt[i][q] = V[i*k + q]  (V column-major: column i = i-th right singular vector).
    // Columns whose singular value is zero to rounding (rank-deficient cores: the step-0 fit, where every sample is
    // identical) are replaced by an orthonormal completion, as LAPACK returns — dividing rounding noise by a
    // near-zero S would put NaNs or non-orthogonal garbage into core_{j-1}.
    const double s_floor = S[0] * (double)(m > k ? m : k) * DBL_EPSILON;
    for (int c = 0; c < k; c++) {
        if (S[c] > s_floor && S[c] > 0.0) {
            const double inv = 1.0 / S[c];
            for (int i = lane; i < m; i += 32) G[c * m + i] *= inv;
            __syncwarp();
            continue;
        }
        for (int trial = 0; trial < m; trial++) {             // e_trial minus its projection on columns [0, c), twice
            for (int i = lane; i < m; i += 32) comp[i] = (i == trial) ? 1.0 : 0.0;
            __syncwarp();
            for (int pass = 0; pass < 2; pass++) {
                for (int q = 0; q < c; q++) {
                    double t = 0.0;
                    for (int i = lane; i < m; i += 32) t += G[q * m + i] * comp[i];
                    t = wsum(t);
                    for (int i = lane; i < m; i += 32) comp[i] -= t * G[q * m + i];
                    __syncwarp();
                }
            }
            double nn = 0.0;
            for (int i = lane; i < m; i += 32) nn += comp[i] * comp[i];
            nn = sqrt(wsum(nn));
            if (nn > 0.5) {
                for (int i = lane; i < m; i += 32) G[c * m + i] = comp[i] / nn;
                break;
            }
            __syncwarp();
        }
        __syncwarp();
    }
    const bool adapt = a.expand && (S[k - 1] > min_sv);
    int kn = k;
    if (!adapt) {
        // W = (S * Vt) @ R : W[i][c] = sum_q Vt[i][q] * S[q] * R[q][c]
        for (int e = lane; e < k * cols; e += 32) {
            const int i = e / cols, c = e % cols;
            double acc = 0.0;
            for (int q = 0; q < k; q++) acc += (S[q] * V[i * k + q]) * Rm[q * cols + c];
            W[e] = acc;
        }
    } else {
        // Gm = Vt @ R  (k x cols), rows are "a" of V[a, b, c]
        for (int e = lane; e < k * cols; e += 32) {
            const int i = e / cols, c = e % cols;
            double acc = 0.0;
            for (int q = 0; q < k; q++) acc += V[i * k + q] * Rm[q * cols + c];
            W[e] = acc;
        }
        __syncwarp();
        // U_comp over (a,b) = rows of the unfolding: ones - U (U^T ones), twice
        for (int i = lane; i < m; i += 32) comp[i] = 1.0;
        __syncwarp();
        for (int pass = 0; pass < 2; pass++) {
            for (int c = 0; c < k; c++) {
                double t = 0.0;
                for (int i = lane; i < m; i += 32) t += G[c * m + i] * comp[i];
                t = wsum(t);
                if (lane == 0) proj[c] = t;
            }
            __syncwarp();
            for (int i = lane; i < m; i += 32) {
                double t = 0.0;
                for (int c = 0; c < k; c++) t += G[c * m + i] * proj[c];
                comp[i] -= t;
            }
            __syncwarp();
        }
        double un = 0.0;
        for (int i = lane; i < m; i += 32) un += comp[i] * comp[i];
        un = sqrt(wsum(un));
        // V_comp over (b,c) = columns of Gm: ones - Gm^T (Gm ones), twice; reuse Rm as scratch for V_comp
        double* vcomp = Rm;
        for (int c = lane; c < cols; c += 32) vcomp[c] = 1.0;
        __syncwarp();
        for (int pass = 0; pass < 2; pass++) {
            for (int i = 0; i < k; i++) {
                double t = 0.0;
                for (int c = lane; c < cols; c += 32) t += W[i * cols + c] * vcomp[c];
                t = wsum(t);
                if (lane == 0) proj[i] = t;
            }
            __syncwarp();
            for (int c = lane; c < cols; c += 32) {
                double t = 0.0;
                for (int i = 0; i < k; i++) t += W[i * cols + c] * proj[i];
                vcomp[c] -= t;
            }
            __syncwarp();
        }
        double vn = 0.0;
        for (int c = lane; c < cols; c += 32) vn += vcomp[c] * vcomp[c];
        vn = sqrt(wsum(vn));
        if (un != 0.0 && vn != 0.0) {       // the reference appends each side independently; a zero norm on one side
            kn = k + 1;                     // only would leave it with mismatched core shapes, so both are required
            for (int i = lane; i < m; i += 32) G[k * m + i] = comp[i] / un;
            for (int c = lane; c < cols; c += 32) W[k * cols + c] = vcomp[c] / vn;
            if (lane == 0) S[k] = 0.01 * min_sv;
            __syncwarp();
            // W = diag(S_ext) @ V_ext
            for (int e = lane; e < kn * cols; e += 32) W[e] *= S[e / cols];
        } else {
            for (int e = lane; e < k * cols; e += 32) W[e] *= S[e / cols];   // rank_adaptation's W without the new row
        }
    }
    __syncwarp();
    for (int e = lane; e < m * kn; e += 32) a.core_prev[e] = G[(e % kn) * m + (e / kn)];
    for (int e = lane; e < kn * cols; e += 32) a.core_cur[e] = W[e];
    if (a.do_reg)
        for (int c = lane; c < kn; c += 32) a.gamma[c] = inv_sv(S[c], min_sv);
    if (lane == 0) {
        *a.rank_out = kn;
        if (kn != k) *a.adapted_out = 1;
    }
}

struct SalsaRightArgs {
    double* core_cur;      // (rl, p, k)
    double* core_next;     // (k, p, rn)
    int rl, p, k, rn;
    const double* min_sv;
    int do_reg;
    double* theta;         // k values out
};

__global__ void __launch_bounds__(32) k_salsa_right(SalsaRightArgs a) {
    __shared__ double G[kMaxM * kMaxK], V[kMaxK * kMaxK], S[kMaxK];
    __shared__ double Rm[kMaxK * kMaxP * kSalsaR], Z[kMaxK * kMaxP * kSalsaR];
    const int lane = threadIdx.x;
    const int m = a.rl * a.p, k = a.k, cols = a.p * a.rn;
    const double min_sv = *a.min_sv;
    for (int e = lane; e < m * k; e += 32) G[(e % k) * m + (e / k)] = a.core_cur[e];
    for (int e = lane; e < k * cols; e += 32) Rm[e] = a.core_next[e];
    __syncwarp();
    warp_jacobi_svd(G, V, S, m, k);
    // core_cur <- U * S = G ; core_next <- Vt @ R
    for (int e = lane; e < m * k; e += 32) a.core_cur[e] = G[(e % k) * m + (e / k)];
    for (int e = lane; e < k * cols; e += 32) {
        const int i = e / cols, c = e % cols;
        double acc = 0.0;
        for (int q = 0; q < k; q++) acc += V[i * k + q] * Rm[q * cols + c];
        Z[e] = acc;
    }
    __syncwarp();
    for (int e = lane; e < k * cols; e += 32) a.core_next[e] = Z[e];
    if (a.do_reg)
        for (int c = lane; c < k; c += 32) a.theta[c] = inv_sv(S[c], min_sv);
}

// state = {eta, omega, min_sv}.  mode 0: eta = min(res / N, 1e4)  (als.py:168-172, first micro-step of the first sweep);
// mode 1: end-of-sweep update (als.py:288-298).
__global__ void k_salsa_scalars(double* state, const double* res_sum, double n_samples, double freq_omega, int mode) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double r = *res_sum / n_samples;
    if (mode == 0) {
        state[0] = fmin(r, 10000.0);
    } else {
        const double eta_old = state[0];
        const double eta = r;
        const double rel = fabs(eta_old - eta) / eta_old;
        double omega = fmin(state[1] / freq_omega, fmax(eta, sqrt(eta)));
        omega = fmax(omega, rel);
        state[0] = eta;
        state[1] = omega;
        state[2] = fmax(fmin(0.2 * omega, 0.2 * rel), 0.0);
    }
}

}  // namespace

int launch_salsa_left(double* core_prev, double* core_cur, int rp, int p, int k, int rn, const double* min_sv, int expand,
                      int do_reg, double* gamma, int* rank_out, int* adapted_out, cudaStream_t st) {
    if (rp * p > kMaxM || k > kMaxK || p * rn > kMaxP * kSalsaR) return fail_msg(2, "salsa: core shape (%d,%d,%d)-(%d) too large", rp, p, k, rn);
    if (rp * p < k) return fail_msg(2, "salsa: reduced SVD needs r_{j-1}*p >= r_j (%d*%d < %d)", rp, p, k);
    SalsaLeftArgs a{core_prev, core_cur, rp, p, k, rn, min_sv, expand, do_reg, gamma, rank_out, adapted_out};
    k_salsa_left<<<1, 32, 0, st>>>(a);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_salsa_right(double* core_cur, double* core_next, int rl, int p, int k, int rn, const double* min_sv, int do_reg,
                       double* theta, cudaStream_t st) {
    if (rl * p > kMaxM || k > kMaxK || p * rn > kMaxP * kSalsaR) return fail_msg(2, "salsa: core shape (%d,%d,%d)-(%d) too large", rl, p, k, rn);
    if (rl * p < k) return fail_msg(2, "salsa: reduced SVD needs r_j*p >= r_{j+1} (%d*%d < %d)", rl, p, k);
    SalsaRightArgs a{core_cur, core_next, rl, p, k, rn, min_sv, do_reg, theta};
    k_salsa_right<<<1, 32, 0, st>>>(a);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_salsa_scalars(double* state, const double* res_sum, double n_samples, double freq_omega, int mode, cudaStream_t st) {
    k_salsa_scalars<<<1, 32, 0, st>>>(state, res_sum, n_samples, freq_omega, mode);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace bsde
