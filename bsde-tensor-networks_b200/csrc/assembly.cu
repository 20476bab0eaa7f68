// K4 — assembly of the local normal equations of one TT core over the sample batch
// (reference: ALS.micro_optimization, core/optimisation/als.py:44-55:  A = V^T V,  Y = V^T y  with
//  V[s,(a,m,b)] = L_j[s,a] phi_j[s,m] R_j[s,b]).
//
// B200 facts this kernel is built on (bench/micro/fp64_peak.cu, measured): FP64 DMMA and DFMA issue at the same
// rate (37 TFLOP/s) and share one pipe, and every f64 mma.sync shape is lowered to DMMA.8x8x4.  So the cost of
// assembly is the number of FP64-pipe slots, and the plain V^T V product (n^2 FMAs / sample, or n(n+1)/2 on
// symmetric tiles) wastes most of them: V is a Kronecker product, hence
//
//   A[(a,m,b),(a',m',b')] = sum_s (L_a L_a')(phi_m phi_m')(R_b R_b')
//
// depends only on the unordered pairs {a,a'}, {m,m'}, {b,b'} — and for monomials phi_m phi_m' = x^(m+m')
// only on m+m'.  We therefore accumulate the *moment tensors*
//
//   M1[beta][lambda][rho] = sum_s B_beta(s) * LL_lambda(s) * RR_rho(s)     (nB x nLL x nRR)
//   M2[m][a][b]           = sum_s (phi_m y)(s) * L_a(s) * R_b(s)           (= V^T y)
//   yy                    = sum_s y_s^2
//
// (r=p=4: 700 + 64 + 1 numbers instead of 2080 + 64 + 1) as two skinny GEMMs over the sample axis on the
// DMMA pipe: rows = basis-pair functions (one 8-row tile when 2p-1 <= 8), columns = LL (x) RR products.
// The solve kernel (dense.cu) expands the moments into A and b.
//
// Execution: each warp owns 32 consecutive samples per iteration.  Phase 1 (lane = sample, coalesced loads)
// evaluates the basis and stages the per-sample factor functions in shared memory; phase 2 runs 8 k-steps of
// DMMA.8x8x4 (4 samples each) over all tiles, building each B fragment as one product of two staged factors.
// Warps never synchronise with each other until the final fixed-order reduction, so results are bitwise
// reproducible for a fixed launch shape.
#include "common.cuh"
#include "kernels.cuh"
#include <algorithm>
#include <cstring>

namespace bsde {

namespace {

constexpr int kAsmThreads = 128;
constexpr int kAsmWarps = kAsmThreads / 32;
constexpr int kStride = 36;      // doubles per staged function per warp: 32 samples + 4 pad (conflict-free fragments)

struct AsmArgs {
    Inputs in;
    int j;
    const double* L;
    const double* R;
    const double* y;
    int rl, rr, mono, nB;
    int F, fLL, fRR, fRow1, fL, fR, fRow2;
    int warp_doubles;            // shared-memory doubles per warp
    const uint16_t* tab;
    const uint8_t* flags;
    double* partial;
    int64_t partial_stride;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// One 8x8 output tile over the 32 staged samples: 8 DMMA.8x8x4, the B fragment formed as a product of two staged
// factors.  `w` already includes the lane's sample slot (lane & 3); the offsets are in doubles.
__device__ __forceinline__ void tile_pair(const double* __restrict__ w, unsigned long long ta, unsigned long long tb,
                                          double (&ca)[2], double (&cb)[2]) {
    const int ra = (int)(ta & 0xffffu), a1 = (int)((ta >> 16) & 0xffffu), a2 = (int)((ta >> 32) & 0xffffu);
    const int rb = (int)(tb & 0xffffu), b1 = (int)((tb >> 16) & 0xffffu), b2 = (int)((tb >> 32) & 0xffffu);
#pragma unroll
    for (int step = 0; step < 8; step++) {
        const double xa = w[ra + step * 4], pa = w[a1 + step * 4] * w[a2 + step * 4];
        const double xb = w[rb + step * 4], pb = w[b1 + step * 4] * w[b2 + step * 4];
        dmma884(ca[0], ca[1], xa, pa);
        dmma884(cb[0], cb[1], xb, pb);
    }
}

template <int NJ, int RB, int PB>
__global__ void __launch_bounds__(kAsmThreads, (NJ >= 16 ? 4 : 6))
k_assemble(AsmArgs a) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ws = sm + (size_t)warp * a.warp_doubles;
    const int group = blockIdx.y;
    const int64_t N = a.in.N;

    // this group's tile table: per (tile, lane group) the staged-function offsets (doubles) of the A fragment row
    // and of the two B fragment factors, packed 3 x 16 bit
    unsigned long long* stab = reinterpret_cast<unsigned long long*>(sm + (size_t)kAsmWarps * a.warp_doubles);
    for (int e = threadIdx.x; e < NJ * 8; e += kAsmThreads) {
        const uint16_t* t = a.tab + ((size_t)(group * NJ) * 8 + e) * 3;
        stab[e] = (unsigned long long)(t[0] * kStride) | ((unsigned long long)(t[1] * kStride) << 16) |
                  ((unsigned long long)(t[2] * kStride) << 32);
    }
    __syncthreads();
    const unsigned long long* mytab = stab + (lane >> 2);

    double acc[NJ][2];
#pragma unroll
    for (int jb = 0; jb < NJ; jb++) acc[jb][0] = acc[jb][1] = 0.0;
    double accyy = 0.0;

    const int64_t nchunks = (N + 31) / 32;
    const int64_t warps_total = (int64_t)gridDim.x * kAsmWarps;
    for (int64_t chunk = (int64_t)blockIdx.x * kAsmWarps + warp; chunk < nchunks; chunk += warps_total) {
        // ---------------- phase 1: lane = sample ----------------
        const int64_t s = chunk * 32 + lane;
        const bool valid = s < N;
        const int64_t sc = valid ? s : N - 1;
        const double vm = valid ? 1.0 : 0.0;
        double Lv[RB], Rv[RB];
#pragma unroll
        for (int q = 0; q < RB; q++) {
            Lv[q] = (q < a.rl) ? (a.L ? a.L[(int64_t)q * N + sc] : 1.0) : 0.0;
            Rv[q] = (q < a.rr) ? (a.R ? a.R[(int64_t)q * N + sc] : 1.0) : 0.0;
        }
        const double yv = a.y[sc] * vm;
        accyy = fma(yv, yv, accyy);
        ws[lane] = 0.0;   // function 0 = ZERO
        {
            int idx = a.fLL;
#pragma unroll
            for (int q = 0; q < RB; q++)
#pragma unroll
                for (int q2 = q; q2 < RB; q2++)
                    if (q < a.rl && q2 < a.rl) { ws[idx * kStride + lane] = Lv[q] * Lv[q2]; idx++; }
            idx = a.fRR;
#pragma unroll
            for (int q = 0; q < RB; q++)
#pragma unroll
                for (int q2 = q; q2 < RB; q2++)
                    if (q < a.rr && q2 < a.rr) { ws[idx * kStride + lane] = Rv[q] * Rv[q2]; idx++; }
#pragma unroll
            for (int q = 0; q < RB; q++) {
                if (q < a.rl) ws[(a.fL + q) * kStride + lane] = Lv[q];
                if (q < a.rr) ws[(a.fR + q) * kStride + lane] = Rv[q];
            }
        }
        if (a.mono) {
            const double x = a.in.data[(int64_t)a.j * N + sc];
            double pw = vm;
#pragma unroll
            for (int k = 0; k < 2 * PB - 1; k++) {
                if (k < a.nB) ws[(a.fRow1 + k) * kStride + lane] = pw;
                if (k < a.in.p) ws[(a.fRow2 + k) * kStride + lane] = pw * yv;
                pw *= x;
            }
        } else {
            double phi[PB];
            basis_values<PB>(a.in, a.j, sc, 0, phi);
            int idx = a.fRow1;
#pragma unroll
            for (int m = 0; m < PB; m++)
#pragma unroll
                for (int m2 = m; m2 < PB; m2++)
                    if (m < a.in.p && m2 < a.in.p) { ws[idx * kStride + lane] = vm * phi[m] * phi[m2]; idx++; }
#pragma unroll
            for (int m = 0; m < PB; m++)
                if (m < a.in.p) ws[(a.fRow2 + m) * kStride + lane] = phi[m] * yv;
        }
        __syncwarp();
        // ---------------- phase 2: DMMA over the staged 32 samples, two independent tiles at a time ----------------
        const double* w = ws + (lane & 3);
#pragma unroll
        for (int jb = 0; jb < NJ; jb += 2) tile_pair(w, mytab[jb * 8], mytab[(jb + 1) * 8], acc[jb], acc[jb + 1]);
        __syncwarp();
    }

    // ---------------- epilogue: fixed-order sum over the CTA's warps ----------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accyy += __shfl_down_sync(0xffffffffu, accyy, o);
    __syncthreads();
#pragma unroll
    for (int jb = 0; jb < NJ; jb++) {
        // C fragment: row = lane/4, cols 2*(lane%4), +1
        ws[jb * 64 + (lane >> 2) * 8 + (lane & 3) * 2 + 0] = acc[jb][0];
        ws[jb * 64 + (lane >> 2) * 8 + (lane & 3) * 2 + 1] = acc[jb][1];
    }
    if (lane == 0) ws[NJ * 64] = accyy;
    __syncthreads();
    double* out = a.partial + (int64_t)blockIdx.x * a.partial_stride;
    for (int e = threadIdx.x; e < NJ * 64; e += kAsmThreads) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kAsmWarps; w++) t += sm[(size_t)w * a.warp_doubles + e];
        out[(int64_t)group * NJ * 64 + e] = t;
    }
    if (threadIdx.x == 0 && group == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kAsmWarps; w++) t += sm[(size_t)w * a.warp_doubles + NJ * 64];
        out[a.partial_stride - 1] = t;
    }
}

// moments[e] = sum over partial rows of partial[row][gather[e]].  One block = 32 consecutive moments x 8 row
// lanes: warp w adds rows w, w+8, ... in order, then the 8 warp sums are added in warp order, so the result is
// a fixed summation tree of the launch shape (bitwise reproducible).
constexpr int kRedWarps = 8;
__global__ void __launch_bounds__(32 * kRedWarps)
k_moment_reduce(const double* __restrict__ partial, int rows, int64_t stride,
                const int32_t* __restrict__ gather, int n_mom, double* __restrict__ moments) {
    __shared__ double part[kRedWarps][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    double a0 = 0.0, a1 = 0.0;
    if (e < n_mom) {
        const int32_t g = gather[e];
        int q = w;
        for (; q + kRedWarps < rows; q += 2 * kRedWarps) {
            a0 += partial[(int64_t)q * stride + g];
            a1 += partial[(int64_t)(q + kRedWarps) * stride + g];
        }
        if (q < rows) a0 += partial[(int64_t)q * stride + g];
    }
    part[w][lane] = a0 + a1;
    __syncthreads();
    if (w == 0 && e < n_mom) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kRedWarps; k++) t += part[k][lane];
        moments[e] = t;
    }
}

template <int NJ, int RB, int PB>
int launch_t(const AsmPlan& plan, const AsmArgs& args, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        BSDE_CUDA_OK(cudaFuncSetAttribute(k_assemble<NJ, RB, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    dim3 grid(plan.grid_x, plan.n_groups);
    k_assemble<NJ, RB, PB><<<grid, kAsmThreads, plan.smem_bytes, st>>>(args);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int NJ>
int launch_nj(const AsmPlan& plan, const AsmArgs& args, cudaStream_t st) {
    const int rb = std::max(plan.rl, plan.rr);
    const bool p4 = plan.p <= 4;
    if (rb <= 4) return p4 ? launch_t<NJ, 4, 4>(plan, args, st) : launch_t<NJ, 4, 8>(plan, args, st);
    return p4 ? launch_t<NJ, 8, 4>(plan, args, st) : launch_t<NJ, 8, 8>(plan, args, st);
}


template <int NJ, int RB, int PB>
int occ_t(size_t smem, int* out) {
    BSDE_CUDA_OK(cudaFuncSetAttribute(k_assemble<NJ, RB, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BSDE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k_assemble<NJ, RB, PB>, kAsmThreads, smem));
    return 0;
}
template <int NJ>
int occ_nj(const AsmPlan& plan, int* out) {
    const int rb = std::max(plan.rl, plan.rr);
    const bool p4 = plan.p <= 4;
    if (rb <= 4) return p4 ? occ_t<NJ, 4, 4>(plan.smem_bytes, out) : occ_t<NJ, 4, 8>(plan.smem_bytes, out);
    return p4 ? occ_t<NJ, 8, 4>(plan.smem_bytes, out) : occ_t<NJ, 8, 8>(plan.smem_bytes, out);
}
// resident CTAs per SM of the kernel instance this plan launches (registers and shared memory)
int asm_blocks_per_sm(const AsmPlan& plan, int* out) {
    if (plan.NJ == 4) return occ_nj<4>(plan, out);
    if (plan.NJ == 8) return occ_nj<8>(plan, out);
    return occ_nj<16>(plan, out);
}

}  // namespace

int asm_plan_build(AsmPlan& P, int rl, int p, int rr, int mono, int64_t N, int sm_count) {
    if (rl < 1 || rr < 1 || p < 1 || rl > 8 || rr > 8 || p > kMaxP)
        return fail_msg(2, "assembly: core shape (%d,%d,%d) outside the supported range (ranks <= 8, p <= %d)", rl, p, rr, kMaxP);
    P.rl = rl; P.p = p; P.rr = rr; P.mono = mono;
    P.n = rl * p * rr;
    P.nLL = rl * (rl + 1) / 2;
    P.nRR = rr * (rr + 1) / 2;
    P.nB = mono ? 2 * p - 1 : p * (p + 1) / 2;
    P.fLL = 1;
    P.fRR = P.fLL + P.nLL;
    P.fRow1 = P.fRR + P.nRR;
    P.fL = P.fRow1 + P.nB;
    P.fR = P.fL + rl;
    P.fRow2 = P.fR + rr;
    P.F = P.fRow2 + p;

    struct Col { int f1, f2, key1, key2; };
    // GEMM 1 columns: (lambda, rho) in two parts so that most fragment factors stay in registers
    std::vector<Col> cols1;
    const int full = P.nRR / 8;
    for (int q = 0; q < full; q++)
        for (int lam = 0; lam < P.nLL; lam++)
            for (int i = 0; i < 8; i++) cols1.push_back({P.fLL + lam, P.fRR + 8 * q + i, lam, 8 * q + i});
    for (int rho = full * 8; rho < P.nRR; rho++)
        for (int lam = 0; lam < P.nLL; lam++) cols1.push_back({P.fLL + lam, P.fRR + rho, lam, rho});
    while (cols1.size() % 8) cols1.push_back({0, 0, -1, -1});
    std::vector<Col> cols2;
    for (int a = 0; a < rl; a++)
        for (int b = 0; b < rr; b++) cols2.push_back({P.fL + a, P.fR + b, a, b});
    while (cols2.size() % 8) cols2.push_back({0, 0, -1, -1});
    const int RT1 = (P.nB + 7) / 8, CT1 = (int)cols1.size() / 8;
    const int RT2 = (p + 7) / 8, CT2 = (int)cols2.size() / 8;
    P.n_jobs = RT1 * CT1 + RT2 * CT2;
    P.NJ = P.n_jobs <= 4 ? 4 : (P.n_jobs <= 8 ? 8 : 16);
    P.n_groups = (P.n_jobs + P.NJ - 1) / P.NJ;
    const int total = P.n_groups * P.NJ;
    P.partial_stride = (int64_t)total * 64 + 1;
    P.n_mom = P.nB * P.nLL * P.nRR + p * rl * rr + 1;

    std::vector<uint16_t> tab((size_t)total * 8 * 3, 0);
    std::vector<uint8_t> flags(total, 7);
    P.h_gather.assign(P.n_mom, 0);
    int job = 0;
    for (int ct = 0; ct < CT1; ct++)
        for (int rt = 0; rt < RT1; rt++, job++)
            for (int i = 0; i < 8; i++) {
                const int beta = rt * 8 + i;
                tab[((size_t)job * 8 + i) * 3 + 0] = (uint16_t)(beta < P.nB ? P.fRow1 + beta : 0);
                const Col& c = cols1[ct * 8 + i];
                tab[((size_t)job * 8 + i) * 3 + 1] = (uint16_t)c.f1;
                tab[((size_t)job * 8 + i) * 3 + 2] = (uint16_t)c.f2;
                if (c.key1 >= 0)
                    for (int r8 = 0; r8 < 8; r8++) {
                        const int b2 = rt * 8 + r8;
                        if (b2 < P.nB) P.h_gather[((size_t)b2 * P.nLL + c.key1) * P.nRR + c.key2] = job * 64 + r8 * 8 + i;
                    }
            }
    const int base2 = P.nB * P.nLL * P.nRR;
    for (int ct = 0; ct < CT2; ct++)
        for (int rt = 0; rt < RT2; rt++, job++)
            for (int i = 0; i < 8; i++) {
                const int m = rt * 8 + i;
                tab[((size_t)job * 8 + i) * 3 + 0] = (uint16_t)(m < p ? P.fRow2 + m : 0);
                const Col& c = cols2[ct * 8 + i];
                tab[((size_t)job * 8 + i) * 3 + 1] = (uint16_t)c.f1;
                tab[((size_t)job * 8 + i) * 3 + 2] = (uint16_t)c.f2;
                if (c.key1 >= 0)
                    for (int r8 = 0; r8 < 8; r8++) {
                        const int m2 = rt * 8 + r8;
                        if (m2 < p) P.h_gather[base2 + ((size_t)m2 * rl + c.key1) * rr + c.key2] = job * 64 + r8 * 8 + i;
                    }
            }
    P.h_gather[P.n_mom - 1] = (int32_t)(P.partial_stride - 1);
    // reload flags: a fragment source is re-read only when some lane group's function differs from the previous tile
    for (int jb = 0; jb < total; jb++) {
        if (jb % P.NJ == 0) { flags[jb] = 7; continue; }
        uint8_t f = 0;
        for (int i = 0; i < 8; i++)
            for (int k = 0; k < 3; k++)
                if (tab[((size_t)jb * 8 + i) * 3 + k] != tab[((size_t)(jb - 1) * 8 + i) * 3 + k]) f |= (uint8_t)(1 << k);
        flags[jb] = f;
    }

    const int warp_doubles = std::max(P.F * kStride, P.NJ * 64 + 1);
    P.smem_bytes = (size_t)kAsmWarps * warp_doubles * sizeof(double) + (size_t)P.NJ * 8 * sizeof(unsigned long long);
    if (P.smem_bytes > 200 * 1024) return fail_msg(2, "assembly: core shape (%d,%d,%d) needs %zu B of shared memory", rl, p, rr, P.smem_bytes);
    const int64_t nchunks = (N + 31) / 32;
    const int64_t want = (nchunks + kAsmWarps - 1) / kAsmWarps;
    int per_sm = 1;
    { const int rc = asm_blocks_per_sm(P, &per_sm); if (rc) return rc; }
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;
    P.grid_x = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count * per_sm));

    BSDE_CUDA_OK(cudaMalloc(&P.d_tab, tab.size() * sizeof(uint16_t)));
    BSDE_CUDA_OK(cudaMalloc(&P.d_flags, flags.size()));
    BSDE_CUDA_OK(cudaMalloc(&P.d_gather, P.h_gather.size() * sizeof(int32_t)));
    BSDE_CUDA_OK(cudaMemcpy(P.d_tab, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    BSDE_CUDA_OK(cudaMemcpy(P.d_flags, flags.data(), flags.size(), cudaMemcpyHostToDevice));
    BSDE_CUDA_OK(cudaMemcpy(P.d_gather, P.h_gather.data(), P.h_gather.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return 0;
}

void asm_plan_free(AsmPlan& P) {
    if (P.d_tab) cudaFree(P.d_tab);
    if (P.d_flags) cudaFree(P.d_flags);
    if (P.d_gather) cudaFree(P.d_gather);
    P.d_tab = nullptr; P.d_flags = nullptr; P.d_gather = nullptr;
}

int launch_assembly(const AsmPlan& P, const Inputs& in, int j, const double* L, const double* R, const double* y,
                    double* partial, double* moments, cudaStream_t st, cudaEvent_t mid) {
    AsmArgs a;
    a.in = in; a.j = j; a.L = L; a.R = R; a.y = y;
    a.rl = P.rl; a.rr = P.rr; a.mono = P.mono; a.nB = P.nB;
    a.F = P.F; a.fLL = P.fLL; a.fRR = P.fRR; a.fRow1 = P.fRow1; a.fL = P.fL; a.fR = P.fR; a.fRow2 = P.fRow2;
    a.warp_doubles = (int)((P.smem_bytes - (size_t)P.NJ * 8 * sizeof(unsigned long long)) / sizeof(double) / kAsmWarps);
    a.tab = P.d_tab; a.flags = P.d_flags;
    a.partial = partial; a.partial_stride = P.partial_stride;
    int rc;
    if (P.NJ == 4) rc = launch_nj<4>(P, a, st);
    else if (P.NJ == 8) rc = launch_nj<8>(P, a, st);
    else rc = launch_nj<16>(P, a, st);
    if (rc) return rc;
    if (mid) BSDE_CUDA_OK(cudaEventRecord(mid, st));
    k_moment_reduce<<<(P.n_mom + 31) / 32, 32 * kRedWarps, 0, st>>>(partial, P.grid_x, P.partial_stride, P.d_gather, P.n_mom, moments);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace bsde
