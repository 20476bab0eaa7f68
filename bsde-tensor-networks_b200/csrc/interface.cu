// K3 / K6 / K7 — per-sample contractions of the tensor train with the basis values.
//
//   k_interface_left   L_{j+1}[s,b] = sum_{a,m} L_j[s,a] phi_j[s,m] C_j[a,m,b]        (als.py:14-17 chained)
//   k_interface_right  R_{j-1}[s,a] = sum_{m,b} C_j[a,m,b] phi_j[s,m] R_j[s,b]
//   k_tt_eval          V(x_s) = G_0[s] G_1[s] ... G_{d-1}[s]                            (utils.py:19-27)
//   k_grad_step        dV/dx_i(x_s) = L_i[s] (C_i . dphi_i[s]) R_i[s]  (+ L_{i+1})      (derivative.py:43-85)
//   k_residual         sum_s ((L_j (x) phi_j (x) R_j)[s] . c_j - y_s)^2                 (als.py:168-172, 288-293)
//
// All are HBM-bound streaming kernels: one thread per sample, sample-contiguous (coalesced) loads and
// stores, the core broadcast from shared memory, basis values evaluated in registers (never stored).
#include "common.cuh"
#include "kernels.cuh"
#include <algorithm>
#include <cstdint>

namespace bsde {

namespace {

constexpr int kThreads = 256;
constexpr int kSB = 2;          // samples carried per thread by the interface kernels

inline int grid_for(int64_t N) {
    int64_t b = (N + kThreads - 1) / kThreads;
    const int64_t cap = 148 * 16;
    return (int)(b < cap ? b : cap);
}

// interface kernels: kSB samples per thread, CTAs a multiple of the SM count when the batch is large
inline int grid_for_blocked(int64_t N) {
    int64_t b = (N + (int64_t)kThreads * kSB - 1) / ((int64_t)kThreads * kSB);
    const int64_t cap = 148 * 16;
    if (b >= 148) b = (b / 148) * 148;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

// Each thread carries kSB samples (a block apart, so every load instruction stays coalesced across the warp): one
// broadcast shared-memory read of a core entry then feeds kSB FMAs.  With one sample per thread the kernels were
// bound by the issue rate of those reads (64 LDS per sample at r = p = 4), at 43 % of HBM bandwidth.

template <int RB, int PB>
__global__ void __launch_bounds__(kThreads)
k_interface_left(Inputs in, int j, const double* __restrict__ core, int r_in, int r_out,
                 const double* __restrict__ Lin, double* __restrict__ Lout) {
    extern __shared__ double csm[];
    const int p = in.p;
    for (int i = threadIdx.x; i < r_in * p * r_out; i += blockDim.x) csm[i] = core[i];
    __syncthreads();
    const int64_t N = in.N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s0 < N; s0 += stride * kSB) {
        double phi[kSB][PB], acc[kSB][RB], lall[kSB][RB];
        int64_t sidx[kSB];
        // all global loads of the chunk are issued before the first use (ncu: 53 % long_scoreboard stalls with the
        // load inside the rank loop): first the interface components, then the inputs of the basis
#pragma unroll
        for (int u = 0; u < kSB; u++) {
            const int64_t s = s0 + u * stride;
            sidx[u] = s < N ? s : N - 1;                 // clamped duplicates are computed and not stored
        }
#pragma unroll
        for (int a = 0; a < RB; a++)
#pragma unroll
            for (int u = 0; u < kSB; u++) lall[u][a] = (a < r_in) ? (Lin ? Lin[(int64_t)a * N + sidx[u]] : 1.0) : 0.0;
#pragma unroll
        for (int u = 0; u < kSB; u++) {
            basis_values<PB>(in, j, sidx[u], 0, phi[u]);
#pragma unroll
            for (int b = 0; b < RB; b++) acc[u][b] = 0.0;
        }
        if (r_in == RB && r_out == RB && p == PB) {
            // full shape (the interior cores of a uniform train): straight-line code, the core read as 16-byte broadcasts
            // (half the shared-memory instructions; same order of the FMAs as the general form, hence the same bits)
            const double2* c2 = reinterpret_cast<const double2*>(csm);
#pragma unroll
            for (int a = 0; a < RB; a++)
#pragma unroll
                for (int m = 0; m < PB; m++) {
                    double t[kSB];
#pragma unroll
                    for (int u = 0; u < kSB; u++) t[u] = lall[u][a] * phi[u][m];
#pragma unroll
                    for (int b = 0; b < RB; b += 2) {
                        const double2 cv = c2[((a * PB + m) * RB + b) >> 1];
#pragma unroll
                        for (int u = 0; u < kSB; u++) {
                            acc[u][b] = fma(t[u], cv.x, acc[u][b]);
                            acc[u][b + 1] = fma(t[u], cv.y, acc[u][b + 1]);
                        }
                    }
                }
        } else {
#pragma unroll
        for (int a = 0; a < RB; a++) {
            if (a >= r_in) break;
#pragma unroll
            for (int m = 0; m < PB; m++) {
                if (m < p) {
                    const double* c = csm + (a * p + m) * r_out;
                    double t[kSB];
#pragma unroll
                    for (int u = 0; u < kSB; u++) t[u] = lall[u][a] * phi[u][m];
#pragma unroll
                    for (int b = 0; b < RB; b++)
                        if (b < r_out) {
                            const double cb = c[b];
#pragma unroll
                            for (int u = 0; u < kSB; u++) acc[u][b] = fma(t[u], cb, acc[u][b]);
                        }
                }
            }
        }
        }
#pragma unroll
        for (int u = 0; u < kSB; u++) {
            const int64_t s = s0 + u * stride;
            if (s < N) {
#pragma unroll
                for (int b = 0; b < RB; b++)
                    if (b < r_out) Lout[(int64_t)b * N + s] = acc[u][b];
            }
        }
    }
}

template <int RB, int PB>
__global__ void __launch_bounds__(kThreads)
k_interface_right(Inputs in, int j, const double* __restrict__ core, int r_out, int r_in,
                  const double* __restrict__ Rin, double* __restrict__ Rout) {
    // core is (r_out, p, r_in); Rin has r_in components (null => ones(1)), Rout gets r_out components
    extern __shared__ double csm[];
    const int p = in.p;
    for (int i = threadIdx.x; i < r_out * p * r_in; i += blockDim.x) csm[i] = core[i];
    __syncthreads();
    const int64_t N = in.N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s0 < N; s0 += stride * kSB) {
        double phi[kSB][PB], rb[kSB][RB];
        int64_t sidx[kSB];
#pragma unroll
        for (int u = 0; u < kSB; u++) {
            const int64_t s = s0 + u * stride;
            sidx[u] = s < N ? s : N - 1;
#pragma unroll
            for (int b = 0; b < RB; b++) rb[u][b] = (b < r_in) ? (Rin ? Rin[(int64_t)b * N + sidx[u]] : 1.0) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kSB; u++) basis_values<PB>(in, j, sidx[u], 0, phi[u]);
        if (r_in == RB && r_out == RB && p == PB) {
            // full shape: straight-line code, 16-byte core reads, all RB outputs kept in registers and stored at the end
            const double2* c2 = reinterpret_cast<const double2*>(csm);
            double out[kSB][RB];
#pragma unroll
            for (int a = 0; a < RB; a++) {
                double acc[kSB];
#pragma unroll
                for (int u = 0; u < kSB; u++) acc[u] = 0.0;
#pragma unroll
                for (int m = 0; m < PB; m++) {
                    double t[kSB];
#pragma unroll
                    for (int u = 0; u < kSB; u++) t[u] = 0.0;
#pragma unroll
                    for (int b = 0; b < RB; b += 2) {
                        const double2 cv = c2[((a * PB + m) * RB + b) >> 1];
#pragma unroll
                        for (int u = 0; u < kSB; u++) { t[u] = fma(cv.x, rb[u][b], t[u]); t[u] = fma(cv.y, rb[u][b + 1], t[u]); }
                    }
#pragma unroll
                    for (int u = 0; u < kSB; u++) acc[u] = fma(phi[u][m], t[u], acc[u]);
                }
#pragma unroll
                for (int u = 0; u < kSB; u++) out[u][a] = acc[u];
            }
#pragma unroll
            for (int u = 0; u < kSB; u++) {
                const int64_t s = s0 + u * stride;
                if (s < N) {
#pragma unroll
                    for (int a = 0; a < RB; a++) Rout[(int64_t)a * N + s] = out[u][a];
                }
            }
            continue;
        }
        for (int a = 0; a < r_out; a++) {
            double acc[kSB];
#pragma unroll
            for (int u = 0; u < kSB; u++) acc[u] = 0.0;
#pragma unroll
            for (int m = 0; m < PB; m++) {
                if (m < p) {
                    const double* c = csm + (a * p + m) * r_in;
                    double t[kSB];
#pragma unroll
                    for (int u = 0; u < kSB; u++) t[u] = 0.0;
#pragma unroll
                    for (int b = 0; b < RB; b++)
                        if (b < r_in) {
                            const double cb = c[b];
#pragma unroll
                            for (int u = 0; u < kSB; u++) t[u] = fma(cb, rb[u][b], t[u]);
                        }
#pragma unroll
                    for (int u = 0; u < kSB; u++) acc[u] = fma(phi[u][m], t[u], acc[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < kSB; u++) {
                const int64_t s = s0 + u * stride;
                if (s < N) Rout[(int64_t)a * N + s] = acc[u];
            }
        }
    }
}

// v <- v . G_j for one core, all in registers.
template <int RB, int PB>
__device__ __forceinline__ void chain_step(double (&v)[RB], const double (&phi)[PB], const double* __restrict__ c,
                                           int rl, int p, int rr) {
    double nv[RB];
#pragma unroll
    for (int b = 0; b < RB; b++) nv[b] = 0.0;
#pragma unroll
    for (int a = 0; a < RB; a++) {
        if (a < rl) {
#pragma unroll
            for (int m = 0; m < PB; m++) {
                if (m < p) {
                    const double t = v[a] * phi[m];
#pragma unroll
                    for (int b = 0; b < RB; b++)
                        if (b < rr) nv[b] = fma(t, __ldg(c + (a * p + m) * rr + b), nv[b]);
                }
            }
        }
    }
#pragma unroll
    for (int b = 0; b < RB; b++) v[b] = nv[b];
}

// Samples carried per thread by k_tt_eval (a grid stride apart, so every load stays coalesced): one read of a core entry
// (same address in every lane: a broadcast out of L1) then feeds that many FMAs.  With one sample per thread the kernel
// issued one load per FMA and ran at 7 TFLOP/s (1.9 ms for 2^20 samples through 100 rank-4 cores, 13 GFLOP); the chain
// product is FP64-bound, not HBM-bound: 128 flops per 8 bytes read at r = p = 4.
template <int RB, int PB>
struct EvalBlocking { static constexpr int S = RB * PB <= 16 ? 4 : 2; };

template <int RB, int PB>
__global__ void __launch_bounds__(kThreads)
k_tt_eval(Inputs in, const double* __restrict__ cores, const int64_t* __restrict__ core_off,
          const int* __restrict__ ranks, double* __restrict__ out) {
    constexpr int S = EvalBlocking<RB, PB>::S;
    const int64_t N = in.N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int p = in.p;
    for (int64_t s0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s0 < N; s0 += stride * S) {
        int64_t sidx[S];
#pragma unroll
        for (int u = 0; u < S; u++) { const int64_t s = s0 + u * stride; sidx[u] = s < N ? s : N - 1; }   // clamped duplicates are not stored
        double v[S][RB];
#pragma unroll
        for (int u = 0; u < S; u++)
#pragma unroll
            for (int b = 0; b < RB; b++) v[u][b] = (b == 0) ? 1.0 : 0.0;
        // (points: the next asset's coordinates are loaded one asset ahead — the kernel runs at one CTA of 8 warps per SM,
        // which does not hide a DRAM round trip per asset)
        const bool points = in.kind != BASIS_PHI;
        double xn[S];
#pragma unroll
        for (int u = 0; u < S; u++) xn[u] = points ? __ldg(in.data + sidx[u]) : 0.0;
        for (int j = 0; j < in.d; j++) {
            double phi[S][PB];
            if (points) {
                double xc[S];
#pragma unroll
                for (int u = 0; u < S; u++) xc[u] = xn[u];
                if (j + 1 < in.d) {
#pragma unroll
                    for (int u = 0; u < S; u++) xn[u] = __ldg(in.data + (int64_t)(j + 1) * N + sidx[u]);
                }
#pragma unroll
                for (int u = 0; u < S; u++) basis_from_x<PB>(in, xc[u], 0, phi[u]);
            } else {
#pragma unroll
                for (int u = 0; u < S; u++) basis_values<PB>(in, j, sidx[u], 0, phi[u]);
            }
            // v <- v . G_j, the same operation order per sample as chain_step
            const double* __restrict__ c = cores + core_off[j];
            const int rl = ranks[j], rr = ranks[j + 1];
            double nv[S][RB];
#pragma unroll
            for (int u = 0; u < S; u++)
#pragma unroll
                for (int b = 0; b < RB; b++) nv[u][b] = 0.0;
            if (RB * PB >= 2 && RB * PB <= 16 && rl == RB && rr == RB && p == PB && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
                // interior core of a train with uniform ranks (uniform over the grid): compile-time shape, the PB x RB
                // entries of row a fetched with 16-byte loads before their FMAs (ncu on the run-time-shaped form below:
                // a third of the stall samples waited for one core entry at a time)
                const double2* __restrict__ c2 = reinterpret_cast<const double2*>(c);
#pragma unroll
                for (int a = 0; a < RB; a++) {
                    double2 cv[PB * RB / 2];
#pragma unroll
                    for (int k = 0; k < PB * RB / 2; k++) cv[k] = __ldg(c2 + a * (PB * RB / 2) + k);
#pragma unroll
                    for (int m = 0; m < PB; m++) {
                        double t[S];
#pragma unroll
                        for (int u = 0; u < S; u++) t[u] = v[u][a] * phi[u][m];
#pragma unroll
                        for (int b = 0; b < RB; b++) {
                            const int e = m * RB + b;
                            const double cvb = (e & 1) ? cv[e >> 1].y : cv[e >> 1].x;
#pragma unroll
                            for (int u = 0; u < S; u++) nv[u][b] = fma(t[u], cvb, nv[u][b]);
                        }
                    }
                }
            } else {
#pragma unroll
            for (int a = 0; a < RB; a++) {
                if (a < rl) {
#pragma unroll
                    for (int m = 0; m < PB; m++) {
                        if (m < p) {
                            double t[S];
#pragma unroll
                            for (int u = 0; u < S; u++) t[u] = v[u][a] * phi[u][m];
#pragma unroll
                            for (int b = 0; b < RB; b++)
                                if (b < rr) {
                                    const double cv = __ldg(c + (a * p + m) * rr + b);
#pragma unroll
                                    for (int u = 0; u < S; u++) nv[u][b] = fma(t[u], cv, nv[u][b]);
                                }
                        }
                    }
                }
            }
            }
#pragma unroll
            for (int u = 0; u < S; u++)
#pragma unroll
                for (int b = 0; b < RB; b++) v[u][b] = nv[u][b];
        }
#pragma unroll
        for (int u = 0; u < S; u++) {
            const int64_t s = s0 + u * stride;
            if (s < N) out[s] = v[u][0];
        }
    }
}

// One asset of multi_derivative: z_i[s] = L_i . (C_i . dphi_i) . R_i, and L_{i+1} = L_i . (C_i . phi_i).
template <int RB, int PB>
__global__ void __launch_bounds__(kThreads)
k_grad_step(Inputs in, int i, const double* __restrict__ core, int rl, int rr,
            const double* __restrict__ Lin, const double* __restrict__ Rin,
            double* __restrict__ Lout, double* __restrict__ z) {
    extern __shared__ double csm[];
    const int p = in.p;
    for (int k = threadIdx.x; k < rl * p * rr; k += blockDim.x) csm[k] = core[k];
    __syncthreads();
    const int64_t N = in.N;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        double phi[PB], dphi[PB];
        basis_values<PB>(in, i, s, 0, phi);
        basis_values<PB>(in, i, s, 1, dphi);
        double rb[RB], acc[RB];
#pragma unroll
        for (int b = 0; b < RB; b++) {
            rb[b] = (b < rr) ? (Rin ? Rin[(int64_t)b * N + s] : 1.0) : 0.0;
            acc[b] = 0.0;
        }
        double zz = 0.0;
        for (int a = 0; a < rl; a++) {
            const double la = Lin ? Lin[(int64_t)a * N + s] : 1.0;
#pragma unroll
            for (int m = 0; m < PB; m++) {
                if (m < p) {
                    const double* c = csm + (a * p + m) * rr;
                    const double t = la * phi[m];
                    double u = 0.0;
#pragma unroll
                    for (int b = 0; b < RB; b++) {
                        if (b < rr) {
                            acc[b] = fma(t, c[b], acc[b]);
                            u = fma(c[b], rb[b], u);
                        }
                    }
                    zz = fma(la * dphi[m], u, zz);
                }
            }
        }
        z[s] = zz;
        if (Lout) {
#pragma unroll
            for (int b = 0; b < RB; b++)
                if (b < rr) Lout[(int64_t)b * N + s] = acc[b];
        }
    }
}

// Row i of the Hessian of the train at every sample (reference: core/calculus/hessian.py:6-26, O(d^2) contractions there):
//   H_ii = L_i (C_i . ddphi_i) R_i,      H_ij = L_i (C_i . dphi_i) [prod_{i<k<j} (C_k . phi_k)] (C_j . dphi_j) R_j   (j > i)
// one thread per sample carries the row vector v = L_i (C_i . dphi_i) prod (C_k . phi_k) to the right.  rbond[j] = R_j
// (r_{j+1} slabs of N, null for the last core); out is H [d][d][N], both triangles written.
struct HessArgs {
    const double* cores;
    const int64_t* core_off;      // device
    const int* ranks;             // device
    const double* const* rbond;   // device array of d pointers
    const double* Lin;            // L_i (ranks[i] slabs) or null for i = 0
    double* out;
    int i;
};

template <int RB, int PB>
__global__ void __launch_bounds__(kThreads)
k_hessian_row(Inputs in, HessArgs h) {
    const int64_t N = in.N;
    const int d = in.d, p = in.p, i = h.i;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        double v[RB], phi[PB], dphi[PB], ddphi[PB];
        const int rl = h.ranks[i], rr = h.ranks[i + 1];
        const double* C = h.cores + h.core_off[i];
        basis_values<PB>(in, i, s, 1, dphi);
        basis_values<PB>(in, i, s, 2, ddphi);
#pragma unroll
        for (int b = 0; b < RB; b++) v[b] = 0.0;
        double w2[RB];                               // L_i (C_i . ddphi_i)
#pragma unroll
        for (int b = 0; b < RB; b++) w2[b] = 0.0;
        for (int a = 0; a < rl; a++) {
            const double la = h.Lin ? h.Lin[(int64_t)a * N + s] : 1.0;
#pragma unroll
            for (int m = 0; m < PB; m++)
                if (m < p) {
                    const double* c = C + (a * p + m) * rr;
#pragma unroll
                    for (int b = 0; b < RB; b++)
                        if (b < rr) { v[b] = fma(la * dphi[m], c[b], v[b]); w2[b] = fma(la * ddphi[m], c[b], w2[b]); }
                }
        }
        {
            const double* R = h.rbond[i];
            double hii = 0.0;
#pragma unroll
            for (int b = 0; b < RB; b++)
                if (b < rr) hii = fma(w2[b], R ? R[(int64_t)b * N + s] : 1.0, hii);
            h.out[((int64_t)i * d + i) * N + s] = hii;
        }
        for (int j = i + 1; j < d; j++) {
            const int ra = h.ranks[j], rb = h.ranks[j + 1];
            const double* Cj = h.cores + h.core_off[j];
            const double* R = h.rbond[j];
            basis_values<PB>(in, j, s, 0, phi);
            basis_values<PB>(in, j, s, 1, dphi);
            double nv[RB], hv = 0.0;
#pragma unroll
            for (int b = 0; b < RB; b++) nv[b] = 0.0;
#pragma unroll
            for (int a = 0; a < RB; a++) {
                if (a < ra) {
#pragma unroll
                    for (int m = 0; m < PB; m++)
                        if (m < p) {
                            const double* c = Cj + (a * p + m) * rb;
                            double u = 0.0;
#pragma unroll
                            for (int b = 0; b < RB; b++)
                                if (b < rb) { nv[b] = fma(v[a] * phi[m], c[b], nv[b]); u = fma(c[b], R ? R[(int64_t)b * N + s] : 1.0, u); }
                            hv = fma(v[a] * dphi[m], u, hv);
                        }
                }
            }
            h.out[((int64_t)i * d + j) * N + s] = hv;
            h.out[((int64_t)j * d + i) * N + s] = hv;
#pragma unroll
            for (int b = 0; b < RB; b++) v[b] = nv[b];
        }
    }
}

// Deterministic block sum: warp shuffles, then warp partials combined in warp order.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) scratch[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += scratch[k];
    return t;
}

// sum_s (pred_s - y_s)^2 with pred = (L (x) phi (x) R) . core ; optionally writes pred.
template <int RB, int PB>
__global__ void __launch_bounds__(kThreads)
k_residual(Inputs in, int j, const double* __restrict__ core, int rl, int rr,
           const double* __restrict__ Lin, const double* __restrict__ Rin, const double* __restrict__ y,
           double* __restrict__ pred_out, double* __restrict__ partial) {
    extern __shared__ double csm[];
    __shared__ double scratch[kThreads / 32];
    const int p = in.p;
    for (int k = threadIdx.x; k < rl * p * rr; k += blockDim.x) csm[k] = core[k];
    __syncthreads();
    const int64_t N = in.N;
    double loc = 0.0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        double phi[PB];
        basis_values<PB>(in, j, s, 0, phi);
        double rb[RB];
#pragma unroll
        for (int b = 0; b < RB; b++) rb[b] = (b < rr) ? (Rin ? Rin[(int64_t)b * N + s] : 1.0) : 0.0;
        double pr = 0.0;
        for (int a = 0; a < rl; a++) {
            const double la = Lin ? Lin[(int64_t)a * N + s] : 1.0;
#pragma unroll
            for (int m = 0; m < PB; m++) {
                if (m < p) {
                    const double* c = csm + (a * p + m) * rr;
                    double u = 0.0;
#pragma unroll
                    for (int b = 0; b < RB; b++)
                        if (b < rr) u = fma(c[b], rb[b], u);
                    pr = fma(la * phi[m], u, pr);
                }
            }
        }
        if (pred_out) pred_out[s] = pr;
        if (y) {
            const double e = pr - y[s];
            loc = fma(e, e, loc);
        }
    }
    const double t = block_sum(loc, scratch);
    if (threadIdx.x == 0 && partial) partial[blockIdx.x] = t;
}

// out[e] = sum_p partial[p * stride + e]   (fixed order => bitwise reproducible)
__global__ void k_sum_partials(const double* __restrict__ partial, int n_partials, int64_t stride, int n_elems,
                               double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int q = 0;
    for (; q + 3 < n_partials; q += 4) {
        a0 += partial[(int64_t)q * stride + e];
        a1 += partial[(int64_t)(q + 1) * stride + e];
        a2 += partial[(int64_t)(q + 2) * stride + e];
        a3 += partial[(int64_t)(q + 3) * stride + e];
    }
    for (; q < n_partials; q++) a0 += partial[(int64_t)q * stride + e];
    out[e] = (a0 + a1) + (a2 + a3);
}

template <int RB>
int pick_pb_left(int p, Inputs in, int j, const double* core, int r_in, int r_out, const double* Lin, double* Lout,
                 cudaStream_t st) {
    const size_t sm = sizeof(double) * r_in * p * r_out;
    const int g = grid_for_blocked(in.N);
    if (p <= 4) k_interface_left<RB, 4><<<g, kThreads, sm, st>>>(in, j, core, r_in, r_out, Lin, Lout);
    else k_interface_left<RB, 8><<<g, kThreads, sm, st>>>(in, j, core, r_in, r_out, Lin, Lout);
    return 0;
}
template <int RB>
int pick_pb_right(int p, Inputs in, int j, const double* core, int r_out, int r_in, const double* Rin, double* Rout,
                  cudaStream_t st) {
    const size_t sm = sizeof(double) * r_in * p * r_out;
    const int g = grid_for_blocked(in.N);
    if (p <= 4) k_interface_right<RB, 4><<<g, kThreads, sm, st>>>(in, j, core, r_out, r_in, Rin, Rout);
    else k_interface_right<RB, 8><<<g, kThreads, sm, st>>>(in, j, core, r_out, r_in, Rin, Rout);
    return 0;
}

#define BSDE_DISPATCH_RB(R, CALL)                    \
    do {                                             \
        if ((R) <= 1) { CALL(1); }                   \
        else if ((R) <= 2) { CALL(2); }              \
        else if ((R) <= 4) { CALL(4); }              \
        else if ((R) <= 8) { CALL(8); }              \
        else { CALL(16); }                           \
    } while (0)

}  // namespace

int launch_interface_left(const Inputs& in, int j, const double* core, int r_in, int r_out, const double* Lin,
                          double* Lout, cudaStream_t st) {
    if (in.p > kMaxP || r_in > kMaxR || r_out > kMaxR) return fail_msg(2, "interface_left: p=%d r=(%d,%d) exceeds limits", in.p, r_in, r_out);
#define CALL(RBV) pick_pb_left<RBV>(in.p, in, j, core, r_in, r_out, Lin, Lout, st)
    BSDE_DISPATCH_RB((r_out > r_in ? r_out : r_in), CALL);
#undef CALL
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_interface_right(const Inputs& in, int j, const double* core, int r_out, int r_in, const double* Rin,
                           double* Rout, cudaStream_t st) {
    if (in.p > kMaxP || r_in > kMaxR || r_out > kMaxR) return fail_msg(2, "interface_right: p=%d r=(%d,%d) exceeds limits", in.p, r_out, r_in);
#define CALL(RBV) pick_pb_right<RBV>(in.p, in, j, core, r_out, r_in, Rin, Rout, st)
    BSDE_DISPATCH_RB(r_in, CALL);
#undef CALL
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_tt_eval(const Inputs& in, const double* cores, const int64_t* core_off, const int* ranks_dev, int max_rank,
                   double* out, cudaStream_t st) {
    if (in.p > kMaxP || max_rank > kMaxR) return fail_msg(2, "tt_eval: p=%d rank=%d exceeds limits", in.p, max_rank);
    // every thread carries EvalBlocking::S samples a grid stride apart: the grid covers N / S threads
#define CALL(RBV)                                                                                                                   \
    if (in.p <= 4) k_tt_eval<RBV, 4><<<grid_for((in.N + EvalBlocking<RBV, 4>::S - 1) / EvalBlocking<RBV, 4>::S), kThreads, 0, st>>>(in, cores, core_off, ranks_dev, out); \
    else k_tt_eval<RBV, 8><<<grid_for((in.N + EvalBlocking<RBV, 8>::S - 1) / EvalBlocking<RBV, 8>::S), kThreads, 0, st>>>(in, cores, core_off, ranks_dev, out)
    BSDE_DISPATCH_RB(max_rank, CALL);
#undef CALL
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_grad_step(const Inputs& in, int i, const double* core, int rl, int rr, const double* Lin, const double* Rin,
                     double* Lout, double* z, cudaStream_t st) {
    if (in.p > kMaxP || rl > kMaxR || rr > kMaxR) return fail_msg(2, "grad_step: p=%d r=(%d,%d) exceeds limits", in.p, rl, rr);
    const int g = grid_for(in.N);
    const size_t sm = sizeof(double) * rl * in.p * rr;
#define CALL(RBV)                                                                                               \
    if (in.p <= 4) k_grad_step<RBV, 4><<<g, kThreads, sm, st>>>(in, i, core, rl, rr, Lin, Rin, Lout, z);          \
    else k_grad_step<RBV, 8><<<g, kThreads, sm, st>>>(in, i, core, rl, rr, Lin, Rin, Lout, z)
    BSDE_DISPATCH_RB(rr, CALL);
#undef CALL
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int residual_grid(int64_t N) { return grid_for(N); }

int launch_residual(const Inputs& in, int j, const double* core, int rl, int rr, const double* Lin, const double* Rin,
                    const double* y, double* pred_out, double* partial, double* out_sum, cudaStream_t st) {
    if (in.p > kMaxP || rl > kMaxR || rr > kMaxR) return fail_msg(2, "residual: p=%d r=(%d,%d) exceeds limits", in.p, rl, rr);
    const int g = grid_for(in.N);
    const size_t sm = sizeof(double) * rl * in.p * rr;
#define CALL(RBV)                                                                                                       \
    if (in.p <= 4) k_residual<RBV, 4><<<g, kThreads, sm, st>>>(in, j, core, rl, rr, Lin, Rin, y, pred_out, partial);       \
    else k_residual<RBV, 8><<<g, kThreads, sm, st>>>(in, j, core, rl, rr, Lin, Rin, y, pred_out, partial)
    BSDE_DISPATCH_RB(rr, CALL);
#undef CALL
    BSDE_CUDA_OK(cudaGetLastError());
    if (partial && out_sum) {
        k_sum_partials<<<1, 32, 0, st>>>(partial, g, 1, 1, out_sum);
        BSDE_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int launch_sum_partials(const double* partial, int n_partials, int64_t stride, int n_elems, double* out, cudaStream_t st) {
    k_sum_partials<<<(n_elems + 127) / 128, 128, 0, st>>>(partial, n_partials, stride, n_elems, out);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_hessian_row(const Inputs& in, int i, const double* cores, const int64_t* core_off, const int* ranks_dev,
                       const double* const* rbond_dev, const double* Lin, int max_rank, double* out, cudaStream_t st) {
    if (in.p > kMaxP || max_rank > 8) return fail_msg(2, "hessian: p = %d / rank %d outside the supported range", in.p, max_rank);
    HessArgs h{cores, core_off, ranks_dev, rbond_dev, Lin, out, i};
    const int grid = (int)std::min<int64_t>((in.N + kThreads - 1) / kThreads, 148 * 8);
    const bool p4 = in.p <= 4;
    if (max_rank <= 4) {
        if (p4) k_hessian_row<4, 4><<<grid, kThreads, 0, st>>>(in, h); else k_hessian_row<4, 8><<<grid, kThreads, 0, st>>>(in, h);
    } else {
        if (p4) k_hessian_row<8, 4><<<grid, kThreads, 0, st>>>(in, h); else k_hessian_row<8, 8><<<grid, kThreads, 0, st>>>(in, h);
    }
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace bsde
