// K4 — assembly of the local normal equations of one TT core over the sample batch
// (reference: ALS.micro_optimization, core/optimisation/als.py:44-55:  A = V^T V,  Y = V^T y  with
//  V[s,(a,m,b)] = L_j[s,a] phi_j[s,m] R_j[s,b]).
//
// B200 facts this kernel is built on (bench/micro/fp64_peak.cu, measured): FP64 DMMA and DFMA issue at the same
// rate (37 TFLOP/s) and every f64 mma.sync shape is lowered to DMMA.8x8x4 (16 cycles per SM sub-partition).  So the
// cost of assembly is the number of DMMA slots, and the plain V^T V product (n^2 FMAs / sample, or n(n+1)/2 on
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
// (r=p=4: 700 + 64 + 1 numbers instead of 2080 + 64 + 1) as skinny GEMMs over the sample axis on the DMMA pipe.
// The solve kernel (dense.cu) expands the moments into A and b.
//
// Execution: each warp owns 32 consecutive samples per iteration.  Phase 1 (lane = sample, coalesced loads)
// evaluates the basis and stages the per-sample factor functions in shared memory; phase 2 runs 8 k-steps of
// DMMA.8x8x4 (4 samples each) over all output tiles.  Two phase-2 variants:
//
//   k_assemble_rb  register-blocked:  rows = (beta, lambda) pairs, columns = rho.  Per k-step every A fragment
//                  (a product of two staged factors) is used for all column tiles and every B fragment for all row
//                  tiles, so shared memory supplies ~2.4 wavefronts per DMMA and the kernel is DMMA-bound.
//                  Instantiated for a few tile-count classes; shapes are padded up to the next class.
//   k_assemble     generic, table driven:  rows = beta, columns = (lambda, rho) pairs, one independent fragment
//                  triple per tile.  Fewer tiles (15 instead of 20 at r=p=4) but ~5 shared-memory wavefronts per
//                  DMMA, which makes it LSU-bound (ncu: 82 % of the shared-memory pipe, DMMA pipe 52 %).
//                  Fallback for shapes outside the register-blocked classes.
//
// Warps never synchronise with each other until the final fixed-order reduction, so results are bitwise
// reproducible for a fixed launch shape.
#include "common.cuh"
#include "kernels.cuh"
#include <algorithm>
#include <cstring>

namespace bsde {

int g_pdl_launch = 0;      // see common.cuh: launches of a captured sweep carry the programmatic-serialisation attribute
int g_asm_r4_full = 1;     // option "r4_full": 0 = never take k_assemble_r4's specialised instantiation (A/B measurement)

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
    const uint16_t* tab;         // generic: [tiles][8][3]; register-blocked: see asm_plan_build
    int tab_entries;             // register-blocked: uint16 entries to copy into shared memory
    double* partial;
    int64_t partial_stride;
    // fused interface update (FusedInterface, kernels.cuh): the interface on the side the sweep comes from is not read
    // from a.L / a.R but formed from the neighbouring bond, the neighbouring core and that asset's basis, and written
    // back as a by-product
    int fuse;                    // 0 none, 1 left (a.L is the OUTPUT), 2 right (a.R is the OUTPUT)
    int fj, fr_in;               // asset of the neighbouring core; components of f_in
    const double* fcore;
    const double* f_in;          // neighbouring bond, null = ones(1)
    double* f_out;
    int core_off;                // shared-memory offset (doubles) of the staged neighbouring core
    const double* ldL;           // what load_raw reads into RawSample::L / R, and how many components
    const double* ldR;
    int nldL, nldR;
};

// volatile: keeps the DMMAs of one k-step together in program order (independent accumulators back to back)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- phase 1: lane = sample ----
// The raw per-sample inputs of the NEXT chunk are fetched into registers before phase 2 of the current one, so the
// DRAM latency is hidden behind the DMMA work of the same warp (ncu: long_scoreboard was 20 % of the stalls).
template <int RB, int PB>
struct RawSample {
    double L[RB], R[RB], y, vm;
    double xb[PB];      // xb[0] = x (fused bases) or the p explicit basis values
    double xf;          // fused interface update: x of the neighbouring asset
};

template <int RB, int PB>
__device__ __forceinline__ void load_raw(const AsmArgs& a, int lane, int64_t chunk, int64_t nchunks, RawSample<RB, PB>& r) {
    const int64_t N = a.in.N;
    if (chunk >= nchunks) chunk = nchunks - 1;            // past the end: a harmless re-read, never staged
    const int64_t s = chunk * 32 + lane;
    const bool valid = s < N;
    const int64_t sc = valid ? s : N - 1;
    r.vm = valid ? 1.0 : 0.0;
    // (a.ldL, a.ldR: the bonds the raw loads come from — with a fused update one of them is the NEIGHBOURING bond, which
    // fuse_interface turns into this core's interface; selected on the host so that nothing here depends on loaded data)
#pragma unroll
    for (int q = 0; q < RB; q++) {
        r.L[q] = (q < a.nldL) ? (a.ldL ? a.ldL[(int64_t)q * N + sc] : 1.0) : 0.0;
        r.R[q] = (q < a.nldR) ? (a.ldR ? a.ldR[(int64_t)q * N + sc] : 1.0) : 0.0;
    }
    if (a.fuse) r.xf = a.in.data[(int64_t)a.fj * N + sc];
    r.y = a.y[sc];
    if (a.in.kind == BASIS_PHI) {
        const double* src = a.in.data + ((int64_t)a.j * a.in.p) * N + sc;
#pragma unroll
        for (int m = 0; m < PB; m++) r.xb[m] = (m < a.in.p) ? src[(int64_t)m * N] : 0.0;
    } else {
        r.xb[0] = a.in.data[(int64_t)a.j * N + sc];
    }
}

// Fused interface update, bit-identical to k_interface_left / k_interface_right (interface.cu: same basis routine, same
// order of the FMAs):  left   L_j[b]     = sum_{a,m} (L_{j-1}[a] phi_{j-1}[m]) C_{j-1}[a,m,b]
//                      right  R_j[a]     = sum_m phi_{j+1}[m] (sum_b C_{j+1}[a,m,b] R_{j+1}[b])
// The result replaces the raw neighbouring bond in r and is stored for the micro-steps that follow.
template <int RB, int PB>
__device__ __forceinline__ void fuse_interface(const AsmArgs& a, const double* __restrict__ core, int64_t s,
                                               RawSample<RB, PB>& r) {
    double phi[PB];
    double out[RB];
    // full interior shape (monomials, p = PB, both bond ranks = RB): straight-line code, the core read as 16-byte
    // broadcasts — the C3 hot path; every other shape takes the predicated general form below
    if (a.in.kind != BASIS_PHI && a.in.p == PB && a.fr_in == RB && (a.fuse == 1 ? a.rl : a.rr) == RB) {
        if (a.in.kind == BASIS_POLY) monomials_cr<PB>(r.xf, phi);
        else basis_from_x<PB>(a.in, r.xf, 0, phi);
        const double2* c2 = reinterpret_cast<const double2*>(core);
        if (a.fuse == 1) {
#pragma unroll
            for (int b = 0; b < RB; b++) out[b] = 0.0;
#pragma unroll
            for (int q = 0; q < RB; q++)
#pragma unroll
                for (int m = 0; m < PB; m++) {
                    const double t = r.L[q] * phi[m];
#pragma unroll
                    for (int b = 0; b < RB; b += 2) {
                        const double2 cv = c2[((q * PB + m) * RB + b) >> 1];
                        out[b] = fma(t, cv.x, out[b]);
                        out[b + 1] = fma(t, cv.y, out[b + 1]);
                    }
                }
#pragma unroll
            for (int b = 0; b < RB; b++) {
                r.L[b] = out[b];
                if (r.vm != 0.0) a.f_out[(int64_t)b * a.in.N + s] = out[b];
            }
        } else {
#pragma unroll
            for (int q = 0; q < RB; q++) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < PB; m++) {
                    double t = 0.0;
#pragma unroll
                    for (int b = 0; b < RB; b += 2) {
                        const double2 cv = c2[((q * PB + m) * RB + b) >> 1];
                        t = fma(cv.x, r.R[b], t);
                        t = fma(cv.y, r.R[b + 1], t);
                    }
                    acc = fma(phi[m], t, acc);
                }
                out[q] = acc;
            }
#pragma unroll
            for (int q = 0; q < RB; q++) {
                r.R[q] = out[q];
                if (r.vm != 0.0) a.f_out[(int64_t)q * a.in.N + s] = out[q];
            }
        }
        return;
    }
    basis_from_x<PB>(a.in, r.xf, 0, phi);
    const int p = a.in.p;
    if (a.fuse == 1) {
        const int r_out = a.rl;
#pragma unroll
        for (int b = 0; b < RB; b++) out[b] = 0.0;
#pragma unroll
        for (int q = 0; q < RB; q++) {
            if (q >= a.fr_in) break;
#pragma unroll
            for (int m = 0; m < PB; m++) {
                if (m < p) {
                    const double* c = core + (q * p + m) * r_out;
                    const double t = r.L[q] * phi[m];
#pragma unroll
                    for (int b = 0; b < RB; b++)
                        if (b < r_out) out[b] = fma(t, c[b], out[b]);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < RB; b++) {
            r.L[b] = (b < r_out) ? out[b] : 0.0;
            if (b < r_out && r.vm != 0.0) a.f_out[(int64_t)b * a.in.N + s] = out[b];
        }
    } else {
        const int r_out = a.rr;
#pragma unroll
        for (int q = 0; q < RB; q++) {
            double acc = 0.0;
            if (q < r_out) {
#pragma unroll
                for (int m = 0; m < PB; m++) {
                    if (m < p) {
                        const double* c = core + (q * p + m) * a.fr_in;
                        double t = 0.0;
#pragma unroll
                        for (int b = 0; b < RB; b++)
                            if (b < a.fr_in) t = fma(c[b], r.R[b], t);
                        acc = fma(phi[m], t, acc);
                    }
                }
            }
            out[q] = acc;
        }
#pragma unroll
        for (int q = 0; q < RB; q++) {
            r.R[q] = (q < r_out) ? out[q] : 0.0;
            if (q < r_out && r.vm != 0.0) a.f_out[(int64_t)q * a.in.N + s] = out[q];
        }
    }
}

// Stages function 0 = ZERO, LL pairs, RR pairs, basis rows, L, R, phi*y for the 32 samples of one chunk.
template <int RB, int PB>
__device__ __forceinline__ double stage_functions(const AsmArgs& a, double* __restrict__ ws, int lane, const RawSample<RB, PB>& r) {
    const double vm = r.vm;
    const double yv = r.y * vm;
    ws[lane] = 0.0;
    if (a.rl == RB && a.rr == RB && a.in.p == PB) {
        // full shape: every slot index is a compile-time constant — the predicated form below spends ~400 integer /
        // predicate instructions per 32-sample chunk on its run-time shape (ncu source page of k_assemble_rb at rank 2,
        // round 2).  Same values into the same slots as the general form.
        constexpr int nP = RB * (RB + 1) / 2;
        constexpr int fLL = 1, fRR = fLL + nP, fRow1 = fRR + nP;
        int idx = 0;
#pragma unroll
        for (int q = 0; q < RB; q++)
#pragma unroll
            for (int q2 = q; q2 < RB; q2++) {
                ws[(fLL + idx) * kStride + lane] = r.L[q] * r.L[q2];
                ws[(fRR + idx) * kStride + lane] = r.R[q] * r.R[q2];
                idx++;
            }
        if (a.mono) {
            constexpr int nBm = 2 * PB - 1, fL = fRow1 + nBm, fR = fL + RB, fRow2 = fR + RB;
#pragma unroll
            for (int q = 0; q < RB; q++) {
                ws[(fL + q) * kStride + lane] = r.L[q];
                ws[(fR + q) * kStride + lane] = r.R[q];
            }
            const double x = r.xb[0];
            double pw = vm;
#pragma unroll
            for (int k = 0; k < nBm; k++) {
                ws[(fRow1 + k) * kStride + lane] = pw;
                if (k < PB) ws[(fRow2 + k) * kStride + lane] = pw * yv;
                pw *= x;
            }
            return yv * yv;
        }
        constexpr int nBf = PB * (PB + 1) / 2, fL = fRow1 + nBf, fR = fL + RB, fRow2 = fR + RB;
#pragma unroll
        for (int q = 0; q < RB; q++) {
            ws[(fL + q) * kStride + lane] = r.L[q];
            ws[(fR + q) * kStride + lane] = r.R[q];
        }
        double phi[PB];
        if (a.in.kind == BASIS_PHI) {
#pragma unroll
            for (int m = 0; m < PB; m++) phi[m] = r.xb[m];
        } else {
            basis_from_x<PB>(a.in, r.xb[0], 0, phi);
        }
        idx = 0;
#pragma unroll
        for (int m = 0; m < PB; m++)
#pragma unroll
            for (int m2 = m; m2 < PB; m2++) { ws[(fRow1 + idx) * kStride + lane] = vm * phi[m] * phi[m2]; idx++; }
#pragma unroll
        for (int m = 0; m < PB; m++) ws[(fRow2 + m) * kStride + lane] = phi[m] * yv;
        return yv * yv;
    }
    {
        int idx = a.fLL;
#pragma unroll
        for (int q = 0; q < RB; q++)
#pragma unroll
            for (int q2 = q; q2 < RB; q2++)
                if (q < a.rl && q2 < a.rl) { ws[idx * kStride + lane] = r.L[q] * r.L[q2]; idx++; }
        idx = a.fRR;
#pragma unroll
        for (int q = 0; q < RB; q++)
#pragma unroll
            for (int q2 = q; q2 < RB; q2++)
                if (q < a.rr && q2 < a.rr) { ws[idx * kStride + lane] = r.R[q] * r.R[q2]; idx++; }
#pragma unroll
        for (int q = 0; q < RB; q++) {
            if (q < a.rl) ws[(a.fL + q) * kStride + lane] = r.L[q];
            if (q < a.rr) ws[(a.fR + q) * kStride + lane] = r.R[q];
        }
    }
    if (a.mono) {
        const double x = r.xb[0];
        double pw = vm;
#pragma unroll
        for (int k = 0; k < 2 * PB - 1; k++) {
            if (k < a.nB) ws[(a.fRow1 + k) * kStride + lane] = pw;
            if (k < a.in.p) ws[(a.fRow2 + k) * kStride + lane] = pw * yv;
            pw *= x;
        }
    } else {
        double phi[PB];
        if (a.in.kind == BASIS_PHI) {
#pragma unroll
            for (int m = 0; m < PB; m++) phi[m] = r.xb[m];
        } else {
            basis_from_x<PB>(a.in, r.xb[0], 0, phi);
        }
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
    return yv * yv;
}

// ---- epilogue: fixed-order sum of the warps' tiles, one partial row per CTA ----
template <int NT, int WARPS = kAsmWarps>
__device__ __forceinline__ void write_partials(const AsmArgs& a, double* sm, double* ws, const double (&acc)[NT][2],
                                               double accyy, int lane, int tile_base, bool write_yy) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accyy += __shfl_down_sync(0xffffffffu, accyy, o);
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NT; t++) {
        // C fragment: row = lane/4, cols 2*(lane%4), +1
        ws[t * 64 + (lane >> 2) * 8 + (lane & 3) * 2 + 0] = acc[t][0];
        ws[t * 64 + (lane >> 2) * 8 + (lane & 3) * 2 + 1] = acc[t][1];
    }
    if (lane == 0) ws[NT * 64] = accyy;
    __syncthreads();
    double* out = a.partial + (int64_t)blockIdx.x * a.partial_stride;
    for (int e = threadIdx.x; e < NT * 64; e += WARPS * 32) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) t += sm[(size_t)w * a.warp_doubles + e];
        out[(int64_t)tile_base * 64 + e] = t;
    }
    if (threadIdx.x == 0 && write_yy) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) t += sm[(size_t)w * a.warp_doubles + NT * 64];
        out[a.partial_stride - 1] = t;
    }
}

// =====================================================================================================
// register-blocked variant
// =====================================================================================================
// shared table (uint16 function indices, premultiplied by kStride on load):
//   rows1 [RT1][8][2] = two staged functions whose product is the A fragment of M1, cols1 [CT1][8] = its B fragment
//   rows2 [RT2][8][2] / cols2 [8] = the same for M2                 (which functions: see asm_plan_build's layouts)
// resident CTAs per SM the register allocation is tuned for: 3 (168 registers).  With 4 (128 registers) the operand
// look-ahead spills and the kernel is 30 % slower (measured at both 15 and 20 accumulator tiles).
#ifndef BSDE_RB_MIN_BLOCKS
#define BSDE_RB_MIN_BLOCKS(nt) ((nt) > 20 ? 2 : 3)      // 28 accumulator tiles (general basis at r = p = 4) need more than 168 registers
#endif
// Sections 3 and 4 (RT3 > 0) serve shapes whose column index is just above one tile wide — a general basis at r = p = 4
// has nB = nLL = nRR = 10: rows (lambda, rho) x columns beta would need 13 x 2 tiles for 10 columns.  There the columns
// beta >= 8 are turned into rows:   section 1  rows (lambda, rho)   x columns beta < 8      13 tiles
//                                   section 3  rows (beta', lambda) x columns rho < 8        3 tiles   (beta' = beta - 8)
//                                   section 4  rows (beta', rho')   x columns lambda         1 x 2 tiles (rho' = rho - 8)
// 20 accumulator tiles with M2 instead of 28.  Table: rows1, cols1, rows2, cols2, rows3 [RT3][8][2], cols3 [8], rows4
// [RT4][8][2], cols4 [CT4][8].
#ifndef BSDE_RB_SECT_MIN_BLOCKS
#define BSDE_RB_SECT_MIN_BLOCKS 2      // 19 row tiles of fragment sources + 20 accumulator tiles: 168 registers spill (measured: 167 vs 146 us)
#endif
template <int RT1, int CT1, int RT2, int RB, int PB, int RT3 = 0, int RT4 = 0, int CT4 = 1>
__global__ void __launch_bounds__(kAsmThreads, RT3 > 0 ? BSDE_RB_SECT_MIN_BLOCKS : BSDE_RB_MIN_BLOCKS(RT1 * CT1 + RT2 + RT3 + RT4 * CT4))
k_assemble_rb(AsmArgs a) {
    extern __shared__ double sm[];
    pdl_wait();
    pdl_release();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ws = sm + (size_t)warp * a.warp_doubles;
    unsigned short* stab = reinterpret_cast<unsigned short*>(sm + (size_t)kAsmWarps * a.warp_doubles);
    for (int e = threadIdx.x; e < a.tab_entries; e += kAsmThreads) stab[e] = (unsigned short)(a.tab[e] * kStride);
    const double* score = sm + a.core_off;
    if (a.fuse) {
        const int nc = (a.fuse == 1 ? a.fr_in * a.rl : a.rr * a.fr_in) * a.in.p;
        for (int e = threadIdx.x; e < nc; e += kAsmThreads) sm[a.core_off + e] = a.fcore[e];
    }
    __syncthreads();
    const int g8 = lane >> 2;
    // per-lane fragment sources (offsets in doubles): row tiles of the sections in order 1, 2 (M2), 3, 4, two factors
    // each; column tiles of every section
    constexpr int RT = RT1 + RT2 + RT3 + RT4;
    constexpr int CT4n = RT4 > 0 ? CT4 : 1;
    int rlo[RT], rhi[RT];
    int colA[CT1], colB, colC = 0, colD[CT4n];
    {
        const unsigned short* t = stab;
#pragma unroll
        for (int r = 0; r < RT1; r++) { rlo[r] = t[(r * 8 + g8) * 2]; rhi[r] = t[(r * 8 + g8) * 2 + 1]; }
        t += RT1 * 16;
#pragma unroll
        for (int c = 0; c < CT1; c++) colA[c] = t[c * 8 + g8];
        t += CT1 * 8;
#pragma unroll
        for (int r = 0; r < RT2; r++) { rlo[RT1 + r] = t[(r * 8 + g8) * 2]; rhi[RT1 + r] = t[(r * 8 + g8) * 2 + 1]; }
        t += RT2 * 16;
        colB = t[g8];
        t += 8;
#pragma unroll
        for (int r = 0; r < RT3; r++) { rlo[RT1 + RT2 + r] = t[(r * 8 + g8) * 2]; rhi[RT1 + RT2 + r] = t[(r * 8 + g8) * 2 + 1]; }
        if (RT3 > 0) { t += RT3 * 16; colC = t[g8]; t += 8; }
#pragma unroll
        for (int r = 0; r < RT4; r++) { rlo[RT1 + RT2 + RT3 + r] = t[(r * 8 + g8) * 2]; rhi[RT1 + RT2 + RT3 + r] = t[(r * 8 + g8) * 2 + 1]; }
        if (RT4 > 0) t += RT4 * 16;
#pragma unroll
        for (int c = 0; c < CT4n; c++) colD[c] = RT4 > 0 ? t[c * 8 + g8] : 0;
    }

    constexpr int NT = RT1 * CT1 + RT2 + RT3 + RT4 * CT4;
    constexpr int T2 = RT1 * CT1, T3 = T2 + RT2, T4 = T3 + RT3;      // first accumulator tile of sections 2, 3, 4
    double acc[NT][2];
#pragma unroll
    for (int t = 0; t < NT; t++) acc[t][0] = acc[t][1] = 0.0;
    double accyy = 0.0;

    const int64_t nchunks = (a.in.N + 31) / 32;
    const int64_t warps_total = (int64_t)gridDim.x * kAsmWarps;
    RawSample<RB, PB> raw;
    load_raw<RB, PB>(a, lane, (int64_t)blockIdx.x * kAsmWarps + warp, nchunks, raw);
    for (int64_t chunk = (int64_t)blockIdx.x * kAsmWarps + warp; chunk < nchunks; chunk += warps_total) {
        if (a.fuse) fuse_interface<RB, PB>(a, score, chunk * 32 + lane, raw);
        accyy += stage_functions<RB, PB>(a, ws, lane, raw);
        load_raw<RB, PB>(a, lane, chunk + warps_total, nchunks, raw);      // prefetch the next chunk
        __syncwarp();
        const double* w = ws + (lane & 3);
        // software pipeline inside one warp: the A fragment (a product of two staged factors) of row tile r+1 is
        // formed, and the factors of row tile r+3 are requested, before the DMMAs of row tile r issue — the LDS and
        // DMUL latencies are then covered by DMMA slots of the same warp
        constexpr int R1 = (RT > 1) ? 1 : 0, R2 = (RT > 2) ? 2 : (2 % RT);
        double av = w[rlo[0]] * w[rhi[0]];
        double n0a = w[rlo[R1] + (RT > 1 ? 0 : 4)], n0b = w[rhi[R1] + (RT > 1 ? 0 : 4)];
        double n1a = w[rlo[R2] + (RT > 2 ? 0 : 4)], n1b = w[rhi[R2] + (RT > 2 ? 0 : 4)];
#pragma unroll 1
        for (int step = 0; step < 8; step++, w += 4) {
            double bc[CT1], bd[CT4n];
#pragma unroll
            for (int c = 0; c < CT1; c++) bc[c] = w[colA[c]];
            const double b2 = w[colB];
            const double b3 = RT3 > 0 ? w[colC] : 0.0;
#pragma unroll
            for (int c = 0; c < CT4n; c++) bd[c] = RT4 > 0 ? w[colD[c]] : 0.0;
#pragma unroll
            for (int r = 0; r < RT; r++) {
                const double cur = av;
                av = n0a * n0b;                       // row tile r+1
                n0a = n1a; n0b = n1b;                 // row tile r+2
                // row tile r+3: of this step, or of the next one (the 4 pad samples after the 32 staged ones make
                // the look-ahead of the last step a harmless read)
                const int nr = (r + 3) % RT;
                const int no = ((r + 3) / RT) * 4;
                n1a = w[rlo[nr] + no];
                n1b = w[rhi[nr] + no];
                if (r < RT1) {
#pragma unroll
                    for (int c = 0; c < CT1; c++) dmma884(acc[r * CT1 + c][0], acc[r * CT1 + c][1], cur, bc[c]);
                } else if (r < RT1 + RT2) {
                    dmma884(acc[T2 + (r - RT1)][0], acc[T2 + (r - RT1)][1], cur, b2);
                } else if (r < RT1 + RT2 + RT3) {
                    dmma884(acc[T3 + (r - RT1 - RT2)][0], acc[T3 + (r - RT1 - RT2)][1], cur, b3);
                } else {
#pragma unroll
                    for (int c = 0; c < CT4n; c++)
                        dmma884(acc[T4 + (r - RT1 - RT2 - RT3) * CT4n + c][0], acc[T4 + (r - RT1 - RT2 - RT3) * CT4n + c][1], cur, bd[c]);
                }
            }
        }
        __syncwarp();
    }
    write_partials<NT>(a, sm, ws, acc, accyy, lane, 0, true);
}

// =====================================================================================================
// rank-4 / 4-monomial variant ("RR-stationary"): the interior cores of the C3..C5 workloads (r_l = r_r = p = 4)
// =====================================================================================================
// k_assemble_rb needs two shared-memory fragment loads per DMMA (the two factors of the A fragment), 660 shared-memory
// wavefronts per 32-sample chunk — ncu: the LSU pipe is 76 % busy, which is what caps its DMMA rate at 59 %.  This
// variant arranges the 100 rows (lambda, rho) of M1 so that almost every operand is reused from registers:
//
//   tile lambda = 0..9 :  row g = rho = 0..7                     A = RR_g(s) L_a(s),  B = B_g(s) L_a'(s)
//   tile 10, 11        :  row g = lambda = 0..7, rho = 8 | 9      A = LL_g(s) * RR_8|9(s)
//   tile 12            :  row g < 4: lambda = 8 + g/2, rho = 8 + g%2
//   tile 13 (M2)       :  row g: a = g/2, m_hi = g%2;  column c: b = c/2, m_lo = c%2;  m = 2 m_hi + m_lo
//                          A = (L_a x^(2 m_hi))(s),  B = (R_b x^(m_lo) y)(s)       — the 4 x 4 x 4 entries of V^T y fill
//                          exactly one tile (until round 2: two half-empty tiles, rows (a, b) x columns m, with the
//                          A fragments multiplied out per k-step: 15 DMMA + 13 DMUL per k-step instead of 14 + 11)
//
// (g = lane / 4 = row of the fragment, s = the lane's sample of the k-step).  RR_g, LL_g, the M2 fragments and B_g
// are one fragment load per k-step each; RR_8, RR_9 and L_0..3 depend on the sample only, are staged sample-major and
// read with 3 LDS.128 per k-step (a 16-byte load costs 4 wavefronts however few distinct addresses it has, which is why
// LL_lambda(s) = L_a(s) L_a'(s) is formed in registers as (RR_g L_a) L_a' rather than broadcast from shared memory).
// 8 shared-memory instructions and 22 wavefronts per k-step of 14 DMMAs instead of 32 and 64.
constexpr int kR4Tiles = 14;
constexpr int kR4FmRows = 40;            // function-major rows: RR_0..7 | B_0..6, ZERO | LL_0..7 | L_a x^(2h) (8) | R_b x^l y (8)
constexpr int kR4SmStride = 10;          // sample-major doubles per sample: RR_8, RR_9, L_0..3, 4 pad (conflict-free 16-byte reads)
constexpr int kR4WarpDoubles = kR4FmRows * kStride + 32 * kR4SmStride;

// Warps per CTA (12 / kR4Warps CTAs per SM, 3 warps per scheduler either way).  Measured in round 2, same box, A/B: one CTA
// of 12 warps per SM — a third of the partial rows for k_moment_reduce (148 instead of 444) — is SLOWER: assembly 96.2 vs
// 93.0 us per launch (12-way epilogue sum, one barrier domain), reduce 8.35 vs 8.44 us (it is latency-, not data-bound):
// 257.7 vs 251.8 ms per C3 fit.  Three CTAs of 4 warps stay.
#ifndef BSDE_R4_WARPS
#define BSDE_R4_WARPS 4
#endif
constexpr int kR4Warps = BSDE_R4_WARPS;
constexpr int kR4Threads = kR4Warps * 32;
constexpr int kR4MinBlocks = 12 / kR4Warps;
// FULL = the interior launches of a C3..C5 sweep (192 of 196 per sweep at d = 100): fused interface update from a rank-4
// neighbour, both bonds present with 4 components, N a multiple of 32.  Everything load_raw / fuse_interface decide at run
// time for the general case (component counts, null bonds, tail masking, clamped sample index: ~150 integer / predicate
// instructions per 32-sample chunk, ncu source page of round 1) is compiled out; the arithmetic is the same instruction
// sequence, so the results are bit-identical to the general form.
template <bool FULL>
__device__ __forceinline__ void r4_load_raw(const AsmArgs& a, int lane, int64_t chunk, int64_t nchunks, RawSample<4, 4>& r) {
    if (!FULL) { load_raw<4, 4>(a, lane, chunk, nchunks, r); return; }
    const int64_t N = a.in.N;
    if (chunk >= nchunks) chunk = nchunks - 1;            // past the end: a harmless re-read, never staged
    const int64_t s = chunk * 32 + lane;
    const double* pl = a.ldL + s;
    const double* pr = a.ldR + s;
    r.vm = 1.0;
#pragma unroll
    for (int q = 0; q < 4; q++) { r.L[q] = pl[(int64_t)q * N]; r.R[q] = pr[(int64_t)q * N]; }
    r.xf = a.in.data[(int64_t)a.fj * N + s];
    r.xb[0] = a.in.data[(int64_t)a.j * N + s];
    r.y = a.y[s];
}

template <bool FULL>
__device__ __forceinline__ void r4_fuse(const AsmArgs& a, const double* __restrict__ core, int64_t s, RawSample<4, 4>& r) {
    if (!FULL) { if (a.fuse) fuse_interface<4, 4>(a, core, s, r); return; }
    double phi[4], out[4];
    monomials_cr<4>(r.xf, phi);
    const double2* c2 = reinterpret_cast<const double2*>(core);
    if (a.fuse == 1) {                                   // same operation order as fuse_interface's straight-line form
#pragma unroll
        for (int b = 0; b < 4; b++) out[b] = 0.0;
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const double t = r.L[q] * phi[m];
#pragma unroll
                for (int b = 0; b < 4; b += 2) {
                    const double2 cv = c2[((q * 4 + m) * 4 + b) >> 1];
                    out[b] = fma(t, cv.x, out[b]);
                    out[b + 1] = fma(t, cv.y, out[b + 1]);
                }
            }
#pragma unroll
        for (int b = 0; b < 4; b++) { r.L[b] = out[b]; a.f_out[(int64_t)b * a.in.N + s] = out[b]; }
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double acc = 0.0;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                double t = 0.0;
#pragma unroll
                for (int b = 0; b < 4; b += 2) {
                    const double2 cv = c2[((q * 4 + m) * 4 + b) >> 1];
                    t = fma(cv.x, r.R[b], t);
                    t = fma(cv.y, r.R[b + 1], t);
                }
                acc = fma(phi[m], t, acc);
            }
            out[q] = acc;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) { r.R[q] = out[q]; a.f_out[(int64_t)q * a.in.N + s] = out[q]; }
    }
}

template <bool FULL>
__global__ void __launch_bounds__(kR4Threads, kR4MinBlocks)
k_assemble_r4(AsmArgs a) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nchunks = (a.in.N + 31) / 32;
    const int64_t warps_total = (int64_t)gridDim.x * kR4Warps;
    RawSample<4, 4> raw;
    // (fetching the first chunk ahead of the wait — its inputs are older than the preceding solve whenever the interface
    // update is fused — was measured and gains nothing: 238.9-239.5 against 238.8 ms per C3 fit)
    pdl_wait();
    pdl_release();
    double* ws = sm + (size_t)warp * a.warp_doubles;
    double* wsm = ws + kR4FmRows * kStride;
    const double* score = sm + a.core_off;
    if (a.fuse) {
        const int nc = (a.fuse == 1 ? a.fr_in * a.rl : a.rr * a.fr_in) * a.in.p;
        for (int e = threadIdx.x; e < nc; e += kR4Threads) sm[a.core_off + e] = a.fcore[e];
    }
    ws[15 * kStride + lane] = 0.0;       // the ZERO function (samples 0..31 are all that is ever read)
    __syncthreads();
    const int g = lane >> 2;
    // per-lane fragment sources
    const double* fm = ws + (lane & 3);
    const int o_rr = g * kStride, o_bf = (8 + g) * kStride, o_llg = (16 + g) * kStride;
    const int o_a13 = (24 + g) * kStride, o_b13 = (32 + g) * kStride;
    const double2* smp = reinterpret_cast<const double2*>(wsm + (lane & 3) * kR4SmStride);
    const bool t12 = g < 4, t12l = (g & 2) != 0, t12r = (g & 1) != 0;

    double acc[kR4Tiles][2];
#pragma unroll
    for (int t = 0; t < kR4Tiles; t++) acc[t][0] = acc[t][1] = 0.0;
    double accyy = 0.0;

#ifndef BSDE_R4_NOPREFETCH
    r4_load_raw<FULL>(a, lane, (int64_t)blockIdx.x * kR4Warps + warp, nchunks, raw);
#endif
    for (int64_t chunk = (int64_t)blockIdx.x * kR4Warps + warp; chunk < nchunks; chunk += warps_total) {
#ifdef BSDE_R4_NOPREFETCH
        r4_load_raw<FULL>(a, lane, chunk, nchunks, raw);      // 4 CTAs per SM: the other warps cover the DRAM latency
#endif
#ifdef BSDE_R4_EXPERIMENT_PHASE2_ONLY      // measurement only (wrong results): stage the first chunk, then DMMA phases alone
        if (chunk == (int64_t)blockIdx.x * kR4Warps + warp)
#endif
        {
        r4_fuse<FULL>(a, score, chunk * 32 + lane, raw);
        {   // ---- phase 1: lane = sample ----
            const double vm = FULL ? 1.0 : raw.vm, yv = FULL ? raw.y : raw.y * vm;
            double ll[10], rr[10];
            int idx = 0;
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int q2 = q; q2 < 4; q2++) { ll[idx] = raw.L[q] * raw.L[q2]; rr[idx] = raw.R[q] * raw.R[q2]; idx++; }
            double2* srow = reinterpret_cast<double2*>(wsm + lane * kR4SmStride);
            srow[0] = make_double2(rr[8], rr[9]);
            srow[1] = make_double2(raw.L[0], raw.L[1]);
            srow[2] = make_double2(raw.L[2], raw.L[3]);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                ws[k * kStride + lane] = rr[k];
                ws[(16 + k) * kStride + lane] = ll[k];
            }
            const double x = raw.xb[0];
            double pw = vm;
#pragma unroll
            for (int k = 0; k < 7; k++) {
                ws[(8 + k) * kStride + lane] = pw;
                pw *= x;
            }
            {   // M2 fragments: rows (a, m_hi) = L_a x^(2 m_hi), columns (b, m_lo) = R_b x^(m_lo) y
                const double x2 = x * x, xy = x * yv;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    ws[(24 + 2 * q) * kStride + lane] = raw.L[q];
                    ws[(25 + 2 * q) * kStride + lane] = raw.L[q] * x2;
                    ws[(32 + 2 * q) * kStride + lane] = raw.R[q] * yv;
                    ws[(33 + 2 * q) * kStride + lane] = raw.R[q] * xy;
                }
            }
            accyy += yv * yv;
        }
#if !defined(BSDE_R4_NOPREFETCH) && !defined(BSDE_R4_EXPERIMENT_PHASE2_ONLY)
        r4_load_raw<FULL>(a, lane, chunk + warps_total, nchunks, raw);      // prefetch the next chunk
#endif
        }
        __syncwarp();
        // ---- phase 2: 8 k-steps of 4 samples, 15 DMMA each ----
#ifndef BSDE_R4_UNROLL
#define BSDE_R4_UNROLL 8
#endif
        constexpr int kUnroll = BSDE_R4_UNROLL;
#pragma unroll kUnroll
        for (int step = 0; step < 8; step++) {
            const double* w = fm + step * 4;
            const double2* sp = smp + step * (4 * kR4SmStride / 2);
            const double rrg = w[o_rr], bf = w[o_bf], llg = w[o_llg], a13 = w[o_a13], b13 = w[o_b13];
            const double2 r89 = sp[0], s01 = sp[1], s23 = sp[2];
            // tile lambda = (a, a'):  A = RR_g L_a,  B = B_g L_a'  — 8 products feed 10 DMMAs (the FP64 pipe is shared by
            // DMMA and DMUL, and a 16-byte broadcast load costs 4 wavefronts however few distinct addresses it has, so
            // LL_lambda(s) is neither loaded nor multiplied out)
            const double u0 = rrg * s01.x, u1 = rrg * s01.y, u2 = rrg * s23.x, u3 = rrg * s23.y;
            const double v0 = bf * s01.x, v1 = bf * s01.y, v2 = bf * s23.x, v3 = bf * s23.y;
            dmma884(acc[0][0], acc[0][1], u0, v0);
            dmma884(acc[1][0], acc[1][1], u0, v1);
            dmma884(acc[2][0], acc[2][1], u0, v2);
            dmma884(acc[3][0], acc[3][1], u0, v3);
            dmma884(acc[4][0], acc[4][1], u1, v1);
            dmma884(acc[5][0], acc[5][1], u1, v2);
            dmma884(acc[6][0], acc[6][1], u1, v3);
            dmma884(acc[7][0], acc[7][1], u2, v2);
            dmma884(acc[8][0], acc[8][1], u2, v3);
            dmma884(acc[9][0], acc[9][1], u3, v3);
            dmma884(acc[10][0], acc[10][1], llg * r89.x, bf);
            dmma884(acc[11][0], acc[11][1], llg * r89.y, bf);
            // tile 12: LL_8 = L_2 L_3, LL_9 = L_3 L_3  ->  A = L_2|3 RR_8|9,  B = B_g L_3
            const double c12 = t12 ? (t12l ? s23.y : s23.x) : 0.0;
            dmma884(acc[12][0], acc[12][1], c12 * (t12r ? r89.y : r89.x), v3);
            dmma884(acc[13][0], acc[13][1], a13, b13);
        }
        __syncwarp();
    }
    write_partials<kR4Tiles, kR4Warps>(a, sm, ws, acc, accyy, lane, 0, true);
}

// =====================================================================================================
// generic table-driven variant
// =====================================================================================================
struct TileOff { int r, c1, c2; };
__device__ __forceinline__ TileOff unpack_tile(unsigned long long t) {
    return {(int)(t & 0xffffu), (int)((t >> 16) & 0xffffu), (int)((t >> 32) & 0xffffu)};
}

// Four 8x8 output tiles over the 32 staged samples: per 4-sample step one DMMA.8x8x4 on each of four independent
// accumulators, the B fragment formed as a product of two staged factors.
__device__ __forceinline__ void tile_quad(const double* __restrict__ w, const unsigned long long* __restrict__ tab,
                                          double (&c0)[2], double (&c1)[2], double (&c2)[2], double (&c3)[2]) {
    const TileOff t0 = unpack_tile(tab[0]), t1 = unpack_tile(tab[8]), t2 = unpack_tile(tab[16]), t3 = unpack_tile(tab[24]);
#pragma unroll 1
    for (int step = 0; step < 8; step++) {     // not unrolled: ptxas otherwise regroups the DMMAs by accumulator
        const int o = step * 4;
        const double x0 = w[t0.r + o], p0 = w[t0.c1 + o] * w[t0.c2 + o];
        const double x1 = w[t1.r + o], p1 = w[t1.c1 + o] * w[t1.c2 + o];
        const double x2 = w[t2.r + o], p2 = w[t2.c1 + o] * w[t2.c2 + o];
        const double x3 = w[t3.r + o], p3 = w[t3.c1 + o] * w[t3.c2 + o];
        dmma884(c0[0], c0[1], x0, p0);
        dmma884(c1[0], c1[1], x1, p1);
        dmma884(c2[0], c2[1], x2, p2);
        dmma884(c3[0], c3[1], x3, p3);
    }
}

template <int NJ, int RB, int PB>
__global__ void __launch_bounds__(kAsmThreads, (NJ >= 16 ? 4 : 6))
k_assemble(AsmArgs a) {
    extern __shared__ double sm[];
    pdl_wait();
    pdl_release();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ws = sm + (size_t)warp * a.warp_doubles;
    const int group = blockIdx.y;

    // this group's tile table: per (tile, lane group) the staged-function offsets (doubles) of the A fragment row
    // and of the two B fragment factors, packed 3 x 16 bit
    unsigned long long* stab = reinterpret_cast<unsigned long long*>(sm + (size_t)kAsmWarps * a.warp_doubles);
    for (int e = threadIdx.x; e < NJ * 8; e += kAsmThreads) {
        const uint16_t* t = a.tab + ((size_t)(group * NJ) * 8 + e) * 3;
        stab[e] = (unsigned long long)(t[0] * kStride) | ((unsigned long long)(t[1] * kStride) << 16) |
                  ((unsigned long long)(t[2] * kStride) << 32);
    }
    const double* score = sm + a.core_off;
    if (a.fuse) {
        const int nc = (a.fuse == 1 ? a.fr_in * a.rl : a.rr * a.fr_in) * a.in.p;
        for (int e = threadIdx.x; e < nc; e += kAsmThreads) sm[a.core_off + e] = a.fcore[e];
    }
    __syncthreads();
    const unsigned long long* mytab = stab + (lane >> 2);

    double acc[NJ][2];
#pragma unroll
    for (int jb = 0; jb < NJ; jb++) acc[jb][0] = acc[jb][1] = 0.0;
    double accyy = 0.0;

    const int64_t nchunks = (a.in.N + 31) / 32;
    const int64_t warps_total = (int64_t)gridDim.x * kAsmWarps;
    RawSample<RB, PB> raw;
    load_raw<RB, PB>(a, lane, (int64_t)blockIdx.x * kAsmWarps + warp, nchunks, raw);
    for (int64_t chunk = (int64_t)blockIdx.x * kAsmWarps + warp; chunk < nchunks; chunk += warps_total) {
        // (with several tile groups every group computes the same interface and stores identical values)
        if (a.fuse) fuse_interface<RB, PB>(a, score, chunk * 32 + lane, raw);
        accyy += stage_functions<RB, PB>(a, ws, lane, raw);
        load_raw<RB, PB>(a, lane, chunk + warps_total, nchunks, raw);      // prefetch the next chunk
        __syncwarp();
        const double* w = ws + (lane & 3);
#pragma unroll
        for (int jb = 0; jb < NJ; jb += 4) tile_quad(w, mytab + jb * 8, acc[jb], acc[jb + 1], acc[jb + 2], acc[jb + 3]);
        __syncwarp();
    }
    write_partials<NJ>(a, sm, ws, acc, accyy, lane, group * NJ, group == 0);
}

// sums[e] = sum over the CTAs' partial rows of partial[row][e], e = slot of a partial row (tile * 64 + r * 8 + c, last
// slot = y.y).  Coalesced: one block = 32 consecutive slots x 32 row lanes; warp w adds rows w, w+32, ... in order
// (14 independent loads in flight at 444 rows), then the warp sums are added in warp order — a fixed summation tree of
// the launch shape, hence bitwise reproducible.  The solve kernel reads the sums through AsmPlan::d_expand.
#ifndef BSDE_RED_WARPS
#define BSDE_RED_WARPS 16      // measured on the C3 fit (3 row groups): 8 warps 244.7, 16 warps 237.3, 32 warps 239.1 ms per fit
#endif
constexpr int kRedWarps = BSDE_RED_WARPS;
__global__ void __launch_bounds__(32 * kRedWarps)
k_moment_reduce(const double* __restrict__ partial, int rows, int64_t stride, int n_slots, double* __restrict__ sums,
                PeerExchange px) {
    __shared__ double part[kRedWarps][33];
    pdl_wait();
    pdl_release();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    // gridDim.y row groups (kRedGroups on a single GPU): group g sums its contiguous share of the partial rows into row g of
    // `sums`; the solve kernels add the few group rows in order while staging the moments.  One group of 444 rows keeps
    // only 24 SMs busy for 11 us; six groups finish in a third of that.  Still a fixed summation tree.
    {
        const int per = (rows + gridDim.y - 1) / gridDim.y;
        const int first = blockIdx.y * per;
        partial += (int64_t)first * stride;
        sums += (int64_t)blockIdx.y * stride;
        rows = max(0, min(per, rows - first));
    }
    double acc[16];
#pragma unroll
    for (int u = 0; u < 16; u++) acc[u] = 0.0;
    if (e < n_slots) {
        for (int q0 = w; q0 < rows; q0 += 16 * kRedWarps) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int q = q0 + u * kRedWarps;
                if (q < rows) acc[u] += partial[(int64_t)q * stride + e];
            }
        }
    }
#pragma unroll
    for (int h = 8; h > 0; h >>= 1)
#pragma unroll
        for (int u = 0; u < h; u++) acc[u] += acc[u + h];
    part[w][lane] = acc[0];
    __syncthreads();
    if (w != 0) return;
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kRedWarps; k++) t += part[k][lane];
    if (px.world > 1) {
        // ---- fused all-reduce over NVLink peer memory, ONE one-way trip on the critical path.  Every lane stores its sum
        // straight into the slot (parity, source rank, slot) of every peer's buffer and then polls its own buffer until the
        // peers' values have arrived: the data word is its own flag (8-byte stores are single-copy atomic; an empty slot
        // holds a bit pattern no sum can take), so there is no fence, no flag store and no second round trip as in round 1
        // (data, __threadfence_system, flag; then spin on the flags).  Consumed slots are emptied again by their reader; the
        // writer can touch them next two exchanges later, which it starts only after it has seen this rank's NEXT exchange —
        // issued by the next kernel of this stream, i.e. after the kernel boundary has made the emptying stores visible.
        // Every rank adds the same numbers in rank order, so all replicas end with bit-identical sums.
        const unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;
        const int bid = blockIdx.y * gridDim.x + blockIdx.x;          // (row group, slot block): every group is exchanged on its own
        const unsigned long long seq = px.step[bid];
        const size_t par = (size_t)(seq & 1ull) * px.world;
        const int slot = bid * 32 + lane;
        for (int p = 0; p < px.world; p++)
            if (p != px.rank)
                asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(px.data[p] + (par + px.rank) * kXSlots + slot), "d"(t) : "memory");
        double* mine = px.data[px.rank] + par * kXSlots + slot;
        const double own = t;
        t = 0.0;
        bool late = false;
        const long long t0 = clock64();
        for (int q = 0; q < px.world; q++) {
            double v = own;
            if (q != px.rank) {
                unsigned long long bits;
                do {
                    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(bits) : "l"(mine + (size_t)q * kXSlots) : "memory");
                    if (bits == kEmpty && clock64() - t0 > px.timeout_cycles) { late = true; break; }
                } while (bits == kEmpty);
                v = __longlong_as_double((long long)bits);
                mine[(size_t)q * kXSlots] = __longlong_as_double((long long)kEmpty);
            }
            t += v;
        }
        if (late) { *px.error = 1; t = nan(""); }       // a peer is gone: the fit must not go on with partial sums
        __syncwarp();
        if (lane == 0) px.step[bid] = seq + 1;
    }
    if (e < n_slots) sums[e] = t;
}

// ---- dispatch ----------------------------------------------------------------------------------------
enum { OP_LAUNCH = 0, OP_OCCUPANCY = 1 };

template <class K>
int run_kernel(K kernel, int op, const AsmPlan& plan, const AsmArgs* args, cudaStream_t st, int* occ) {
    if (op == OP_OCCUPANCY) {
        BSDE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        BSDE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kernel, plan.warps * 32, plan.smem_bytes));
        return 0;
    }
    dim3 grid(plan.grid_x, plan.n_groups);
    if (g_pdl_launch) {
        BSDE_CUDA_OK(launch_chain(kernel, grid, dim3(plan.warps * 32), plan.smem_bytes, st, true, *args));
        return 0;
    }
    kernel<<<grid, plan.warps * 32, plan.smem_bytes, st>>>(*args);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

constexpr int kRbClassR4 = 100;       // AsmPlan::rb_class of k_assemble_r4
constexpr int kRbClassSectioned = 8;  // k_assemble_rb<13, 1, 2, RB, PB, 3, 1, 2>: columns beta >= 8 turned into rows (see the kernel)

template <int RB, int PB>
int dispatch_rbpb(int op, const AsmPlan& P, const AsmArgs* a, cudaStream_t st, int* occ) {
    if (P.rb_class == kRbClassR4) {
        if (op == OP_OCCUPANCY) {          // both instantiations need their shared-memory attribute set
            const int rc = run_kernel(k_assemble_r4<true>, op, P, a, st, occ);
            if (rc) return rc;
            return run_kernel(k_assemble_r4<false>, op, P, a, st, occ);
        }
        const bool full = a && a->fuse && a->fr_in == 4 && a->nldL == 4 && a->nldR == 4 && a->ldL && a->ldR && a->in.kind == BASIS_POLY &&
                          a->in.p == 4 && (a->in.N % 32) == 0 && g_asm_r4_full;
        if (full) return run_kernel(k_assemble_r4<true>, op, P, a, st, occ);
        return run_kernel(k_assemble_r4<false>, op, P, a, st, occ);
    }
    if (P.rb_class >= 0) {
        switch (P.rb_class) {
            case 0: return run_kernel(k_assemble_rb<1, 1, 1, RB, PB>, op, P, a, st, occ);
            case 1: return run_kernel(k_assemble_rb<2, 1, 1, RB, PB>, op, P, a, st, occ);
            case 2: return run_kernel(k_assemble_rb<2, 2, 1, RB, PB>, op, P, a, st, occ);
            case 3: return run_kernel(k_assemble_rb<4, 1, 2, RB, PB>, op, P, a, st, occ);
            case 4: return run_kernel(k_assemble_rb<5, 1, 2, RB, PB>, op, P, a, st, occ);
            case 5: return run_kernel(k_assemble_rb<13, 1, 2, RB, PB>, op, P, a, st, occ);
            case 6: return run_kernel(k_assemble_rb<9, 2, 2, RB, PB>, op, P, a, st, occ);
            case 7: return run_kernel(k_assemble_rb<13, 2, 2, RB, PB>, op, P, a, st, occ);
            default: return run_kernel(k_assemble_rb<13, 1, 2, RB, PB, 3, 1, 2>, op, P, a, st, occ);     // kRbClassSectioned
        }
    }
    if (P.NJ == 4) return run_kernel(k_assemble<4, RB, PB>, op, P, a, st, occ);
    if (P.NJ == 8) return run_kernel(k_assemble<8, RB, PB>, op, P, a, st, occ);
    return run_kernel(k_assemble<16, RB, PB>, op, P, a, st, occ);
}

int dispatch(int op, const AsmPlan& P, const AsmArgs* a, cudaStream_t st, int* occ) {
    int rb = std::max(P.rl, P.rr);
    if (a && a->fuse) rb = std::max(rb, a->fr_in);      // the fused update holds the neighbouring bond in the same registers
    const bool p4 = P.p <= 4;
    if (op == OP_OCCUPANCY) {
        // a fused update from a wider neighbour pushes a launch into a wider instantiation: every instantiation this plan
        // can reach gets its shared-memory attribute; the grid follows the occupancy of the narrowest one
        int o2 = 0, o4 = 0, o8 = 0;
        int rc = p4 ? dispatch_rbpb<8, 4>(op, P, a, st, &o8) : dispatch_rbpb<8, 8>(op, P, a, st, &o8);
        if (rc) return rc;
        if (rb <= 4) { rc = p4 ? dispatch_rbpb<4, 4>(op, P, a, st, &o4) : dispatch_rbpb<4, 8>(op, P, a, st, &o4); if (rc) return rc; }
        if (rb <= 2) { rc = p4 ? dispatch_rbpb<2, 4>(op, P, a, st, &o2) : dispatch_rbpb<2, 8>(op, P, a, st, &o2); if (rc) return rc; }
#ifdef BSDE_NO_RB2
        if (rb <= 2) o2 = o4;
#endif
        *occ = rb <= 2 ? o2 : (rb <= 4 ? o4 : o8);
        return 0;
    }
    // (rank <= 2 trains — SALSA starts there — have their own bound: the full-shape fast paths then apply at rank 2)
#ifndef BSDE_NO_RB2
    if (rb <= 2) return p4 ? dispatch_rbpb<2, 4>(op, P, a, st, occ) : dispatch_rbpb<2, 8>(op, P, a, st, occ);
#endif
    if (rb <= 4) return p4 ? dispatch_rbpb<4, 4>(op, P, a, st, occ) : dispatch_rbpb<4, 8>(op, P, a, st, occ);
    return p4 ? dispatch_rbpb<8, 4>(op, P, a, st, occ) : dispatch_rbpb<8, 8>(op, P, a, st, occ);
}

// tile-count classes of the register-blocked kernel: (row tiles of M1, column tiles of M1, row tiles of M2)
// ordered by total tile count RT1*CT1 + RT2
constexpr int kNumRbClasses = 8;
const int kRbClasses[kNumRbClasses][3] = {{1, 1, 1}, {2, 1, 1}, {2, 2, 1}, {4, 1, 2}, {5, 1, 2}, {13, 1, 2}, {9, 2, 2}, {13, 2, 2}};

}  // namespace

int asm_plan_build(AsmPlan& P, int rl, int p, int rr, int mono, int64_t N, int sm_count, int force_generic) {
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
    P.n_mom = P.nB * P.nLL * P.nRR + p * rl * rr + 1;
    P.h_gather.assign(P.n_mom, 0);
    const int base2 = P.nB * P.nLL * P.nRR;
    std::vector<uint16_t> tab;
    size_t tab_smem = 0;
    int n_tiles = 0;

    // ---- register-blocked layouts (k_assemble_rb: A fragment = product of two staged functions, B fragment = one) ----
    //   layout 0: M1 rows (beta, lambda), columns rho;     M2 rows (m, a), columns b
    //   layout 1: M1 rows (lambda, rho),  columns beta;    M2 rows (a, b), columns m      (15 tiles at r = p = 4)
    // The layout and tile-count class with the fewest DMMA tiles is chosen.
    P.rb_class = -1;
    int layout = 0;
    const bool use_r4 = force_generic == 0 && rl == 4 && rr == 4 && p == 4 && mono;      // force_generic 2: rb classes only
    if (force_generic != 1 && !use_r4) {
        int best_tiles = 1 << 30;
        for (int lay = 0; lay < 2; lay++) {
            const int rows1 = lay == 0 ? P.nB * P.nLL : P.nLL * P.nRR, cols1 = lay == 0 ? P.nRR : P.nB;
            const int rows2 = lay == 0 ? p * rl : rl * rr, cols2 = lay == 0 ? rr : p;
            if (cols2 > 8) continue;
            const int needRT1 = (rows1 + 7) / 8, needCT1 = (cols1 + 7) / 8, needRT2 = (rows2 + 7) / 8;
            for (int k = 0; k < kNumRbClasses; k++)
                if (needRT1 <= kRbClasses[k][0] && needCT1 <= kRbClasses[k][1] && needRT2 <= kRbClasses[k][2]) {
                    const int tiles = kRbClasses[k][0] * kRbClasses[k][1] + kRbClasses[k][2];
                    if (tiles < best_tiles) { best_tiles = tiles; P.rb_class = k; layout = lay; }
                    break;
                }
        }
    }

    // sectioned layout (kRbClassSectioned): column index beta just above one tile wide — see k_assemble_rb
    if (force_generic != 1 && !use_r4 && P.nB > 8 && P.nB <= 10 && P.nRR <= 10 && P.nLL <= 16 && P.nLL * P.nRR <= 104 &&
        (P.nB - 8) * P.nLL <= 24 && rl * rr <= 16 && p <= 8) {
        const int cur = P.rb_class >= 0 ? kRbClasses[P.rb_class][0] * kRbClasses[P.rb_class][1] + kRbClasses[P.rb_class][2] : (1 << 30);
        if (20 < cur) P.rb_class = kRbClassSectioned;
    }

    if (P.rb_class == kRbClassSectioned) {
        constexpr int RT1 = 13, RT2 = 2, RT3 = 3, CT4 = 2;
        n_tiles = RT1 + RT2 + RT3 + CT4;
        P.NJ = n_tiles; P.n_groups = 1; P.n_jobs = n_tiles;
        tab.assign((size_t)RT1 * 16 + 8 + RT2 * 16 + 8 + RT3 * 16 + 8 + 16 + CT4 * 8, 0);      // 0 = the ZERO function
        uint16_t* rows1t = tab.data();
        uint16_t* cols1t = rows1t + RT1 * 16;
        uint16_t* rows2t = cols1t + 8;
        uint16_t* cols2t = rows2t + RT2 * 16;
        uint16_t* rows3t = cols2t + 8;
        uint16_t* cols3t = rows3t + RT3 * 16;
        uint16_t* rows4t = cols3t + 8;
        uint16_t* cols4t = rows4t + 16;
        const int T2 = RT1, T3 = T2 + RT2, T4 = T3 + RT3;
        const int nRlo = std::min(P.nRR, 8), nRx = P.nRR - nRlo;
        for (int beta = 0; beta < P.nB; beta++)
            for (int lam = 0; lam < P.nLL; lam++)
                for (int rho = 0; rho < P.nRR; rho++) {
                    int slot;
                    if (beta < 8) {                       // section 1: rows (lambda, rho), column beta
                        const int r = lam * P.nRR + rho;
                        rows1t[r * 2] = (uint16_t)(P.fLL + lam); rows1t[r * 2 + 1] = (uint16_t)(P.fRR + rho);
                        cols1t[beta] = (uint16_t)(P.fRow1 + beta);
                        slot = (r / 8) * 64 + (r % 8) * 8 + beta;
                    } else if (rho < 8) {                 // section 3: rows (beta', lambda), column rho
                        const int r = (beta - 8) * P.nLL + lam;
                        rows3t[r * 2] = (uint16_t)(P.fRow1 + beta); rows3t[r * 2 + 1] = (uint16_t)(P.fLL + lam);
                        cols3t[rho] = (uint16_t)(P.fRR + rho);
                        slot = (T3 + r / 8) * 64 + (r % 8) * 8 + rho;
                    } else {                              // section 4: rows (beta', rho'), columns lambda
                        const int r = (beta - 8) * nRx + (rho - 8);
                        rows4t[r * 2] = (uint16_t)(P.fRow1 + beta); rows4t[r * 2 + 1] = (uint16_t)(P.fRR + rho);
                        cols4t[lam] = (uint16_t)(P.fLL + lam);
                        slot = (T4 + lam / 8) * 64 + r * 8 + lam % 8;
                    }
                    P.h_gather[((size_t)beta * P.nLL + lam) * P.nRR + rho] = slot;
                }
        for (int m = 0; m < p; m++)
            for (int a = 0; a < rl; a++)
                for (int b = 0; b < rr; b++) {
                    const int r = a * rr + b;
                    rows2t[r * 2] = (uint16_t)(P.fL + a); rows2t[r * 2 + 1] = (uint16_t)(P.fR + b);
                    cols2t[m] = (uint16_t)(P.fRow2 + m);
                    P.h_gather[base2 + ((size_t)m * rl + a) * rr + b] = (T2 + r / 8) * 64 + (r % 8) * 8 + m;
                }
        P.tab_entries = (int)tab.size();
        tab_smem = tab.size() * sizeof(uint16_t);
    } else if (use_r4) {
        // k_assemble_r4: see the tile layout in its header comment
        P.rb_class = kRbClassR4;
        n_tiles = kR4Tiles;
        P.NJ = n_tiles; P.n_groups = 1; P.n_jobs = n_tiles;
        for (int beta = 0; beta < P.nB; beta++)
            for (int lam = 0; lam < P.nLL; lam++)
                for (int rho = 0; rho < P.nRR; rho++) {
                    int tile, row;
                    if (rho < 8) { tile = lam; row = rho; }
                    else if (lam < 8) { tile = 10 + (rho - 8); row = lam; }
                    else { tile = 12; row = (lam - 8) * 2 + (rho - 8); }
                    P.h_gather[((size_t)beta * P.nLL + lam) * P.nRR + rho] = tile * 64 + row * 8 + beta;
                }
        for (int m = 0; m < p; m++)
            for (int a = 0; a < rl; a++)
                for (int b = 0; b < rr; b++)
                    P.h_gather[base2 + ((size_t)m * rl + a) * rr + b] = 13 * 64 + (a * 2 + (m >> 1)) * 8 + (b * 2 + (m & 1));
        P.tab_entries = 0;
        tab_smem = 0;
    } else if (P.rb_class >= 0) {
        const int RT1 = kRbClasses[P.rb_class][0], CT1 = kRbClasses[P.rb_class][1], RT2 = kRbClasses[P.rb_class][2];
        n_tiles = RT1 * CT1 + RT2;
        P.NJ = n_tiles; P.n_groups = 1; P.n_jobs = n_tiles;
        tab.assign((size_t)RT1 * 16 + CT1 * 8 + RT2 * 16 + 8, 0);
        uint16_t* t = tab.data();
        uint16_t* rows1t = t;
        uint16_t* cols1t = rows1t + RT1 * 16;
        uint16_t* rows2t = cols1t + CT1 * 8;
        uint16_t* cols2t = rows2t + RT2 * 16;
        // row index of an M1 / M2 entry and its column, per layout
        auto row1 = [&](int beta, int lam, int rho) { return layout == 0 ? beta * P.nLL + lam : lam * P.nRR + rho; };
        auto col1 = [&](int beta, int lam, int rho) { (void)lam; return layout == 0 ? rho : beta; };
        auto row2 = [&](int m, int a, int b) { return layout == 0 ? m * rl + a : a * rr + b; };
        auto col2 = [&](int m, int a, int b) { (void)a; return layout == 0 ? b : m; };
        for (int beta = 0; beta < P.nB; beta++)
            for (int lam = 0; lam < P.nLL; lam++)
                for (int rho = 0; rho < P.nRR; rho++) {
                    const int r = row1(beta, lam, rho), c = col1(beta, lam, rho);
                    rows1t[r * 2] = (uint16_t)(layout == 0 ? P.fRow1 + beta : P.fLL + lam);
                    rows1t[r * 2 + 1] = (uint16_t)(layout == 0 ? P.fLL + lam : P.fRR + rho);
                    cols1t[c] = (uint16_t)(layout == 0 ? P.fRR + rho : P.fRow1 + beta);
                    P.h_gather[((size_t)beta * P.nLL + lam) * P.nRR + rho] = ((r / 8) * CT1 + c / 8) * 64 + (r % 8) * 8 + c % 8;
                }
        for (int m = 0; m < p; m++)
            for (int a = 0; a < rl; a++)
                for (int b = 0; b < rr; b++) {
                    const int r = row2(m, a, b), c = col2(m, a, b);
                    rows2t[r * 2] = (uint16_t)(layout == 0 ? P.fRow2 + m : P.fL + a);
                    rows2t[r * 2 + 1] = (uint16_t)(layout == 0 ? P.fL + a : P.fR + b);
                    cols2t[c] = (uint16_t)(layout == 0 ? P.fR + b : P.fRow2 + m);
                    P.h_gather[base2 + ((size_t)m * rl + a) * rr + b] = (RT1 * CT1 + r / 8) * 64 + (r % 8) * 8 + c;
                }
        P.tab_entries = (int)tab.size();
        tab_smem = tab.size() * sizeof(uint16_t);
    } else {
        // ---- generic layout: rows beta, columns (lambda, rho) pairs; M2 rows m, columns (a, b) ----
        struct Col { int f1, f2, key1, key2; };
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
        n_tiles = P.n_groups * P.NJ;
        tab.assign((size_t)n_tiles * 8 * 3, 0);
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
        P.tab_entries = 0;
        tab_smem = (size_t)P.NJ * 8 * sizeof(unsigned long long);
    }
    P.partial_stride = (int64_t)n_tiles * 64 + 1;
    P.h_gather[P.n_mom - 1] = (int32_t)(P.partial_stride - 1);

    const int tiles_per_cta = P.rb_class >= 0 ? n_tiles : P.NJ;
    P.warp_doubles = std::max(P.F * kStride + 16, tiles_per_cta * 64 + 1);    // +16: look-ahead reads of the pipelined loop
    if (use_r4) P.warp_doubles = std::max(kR4WarpDoubles, kR4Tiles * 64 + 2);
    P.warps = use_r4 ? kR4Warps : kAsmWarps;
    P.core_off = (int)((size_t)P.warps * P.warp_doubles + (((tab_smem + 15) & ~(size_t)15) / sizeof(double)));
    P.smem_bytes = ((size_t)P.core_off + (size_t)8 * kMaxP * 8) * sizeof(double);      // + the neighbouring core of a fused interface update
    if (P.smem_bytes > 200 * 1024) return fail_msg(2, "assembly: core shape (%d,%d,%d) needs %zu B of shared memory", rl, p, rr, P.smem_bytes);
    const int64_t nchunks = (N + 31) / 32;
    const int64_t want = (nchunks + P.warps - 1) / P.warps;
    int per_sm = 1;
    { const int rc = dispatch(OP_OCCUPANCY, P, nullptr, nullptr, &per_sm); if (rc) return rc; }
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;
    P.grid_x = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count * per_sm));

    {   // expansion map of the solve kernel: entry of A / b -> canonical moment
        auto pair_index = [](int a, int b, int r) { if (a > b) std::swap(a, b); return a * r - (a * (a - 1)) / 2 + (b - a); };
        const int n = P.n;
        std::vector<uint16_t> emap((size_t)n * n + n);
        for (int row = 0; row < n; row++) {
            const int a1 = row / (p * rr), m1 = (row / rr) % p, b1 = row % rr;
            for (int col = 0; col < n; col++) {
                const int a2 = col / (p * rr), m2 = (col / rr) % p, b2 = col % rr;
                const int lam = pair_index(a1, a2, rl), rho = pair_index(b1, b2, rr);
                const int beta = mono ? (m1 + m2) : pair_index(m1, m2, p);
                emap[(size_t)row * n + col] = (uint16_t)P.h_gather[(beta * P.nLL + lam) * P.nRR + rho];
            }
            emap[(size_t)n * n + row] = (uint16_t)P.h_gather[base2 + (m1 * rl + a1) * rr + b1];
        }
        if (P.partial_stride > 65535) return fail_msg(2, "assembly: %lld partial slots exceed the 16-bit expansion map", (long long)P.partial_stride);
        BSDE_CUDA_OK(cudaMalloc(&P.d_expand, emap.size() * sizeof(uint16_t)));
        BSDE_CUDA_OK(cudaMemcpy(P.d_expand, emap.data(), emap.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    BSDE_CUDA_OK(cudaMalloc(&P.d_tab, tab.size() * sizeof(uint16_t)));
    BSDE_CUDA_OK(cudaMalloc(&P.d_gather, P.h_gather.size() * sizeof(int32_t)));
    BSDE_CUDA_OK(cudaMemcpy(P.d_tab, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    BSDE_CUDA_OK(cudaMemcpy(P.d_gather, P.h_gather.data(), P.h_gather.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return 0;
}

bool asm_fuse_supported(const AsmPlan& P, const Inputs& in, const FusedInterface& fi) {
    if (in.kind == BASIS_PHI) return false;
    if (fi.r_in < 1 || fi.r_in > 8) return false;
    if (P.rb_class == kRbClassR4 && fi.r_in > 4) return false;     // that kernel holds interfaces of at most 4 components
    return true;
}

void asm_plan_free(AsmPlan& P) {
    if (P.d_tab) cudaFree(P.d_tab);
    if (P.d_gather) cudaFree(P.d_gather);
    if (P.d_expand) cudaFree(P.d_expand);
    P.d_tab = nullptr; P.d_gather = nullptr; P.d_expand = nullptr;
}

int launch_assembly(const AsmPlan& P, const Inputs& in, int j, const double* L, const double* R, const double* y,
                    double* partial, double* moments, cudaStream_t st, cudaEvent_t mid, const PeerExchange* px,
                    const FusedInterface* fi, int moment_groups) {
    AsmArgs a;
    a.in = in; a.j = j; a.L = L; a.R = R; a.y = y;
    a.fuse = 0; a.fj = 0; a.fr_in = 0; a.fcore = nullptr; a.f_in = nullptr; a.f_out = nullptr; a.core_off = P.core_off;
    if (fi && fi->side) {
        if (in.kind == BASIS_PHI) return fail_msg(2, "assembly: the fused interface update needs a fused basis (x inputs)");
        if (fi->r_in < 1 || fi->r_in > 8 || !fi->core || !fi->out) return fail_msg(2, "assembly: bad fused interface update");
        a.fuse = fi->side; a.fj = fi->j; a.fr_in = fi->r_in; a.fcore = fi->core; a.f_in = fi->in; a.f_out = fi->out;
        if (fi->side == 1) a.L = fi->out; else a.R = fi->out;     // (documentation only: the kernel forms them itself)
    }
    a.ldL = a.fuse == 1 ? a.f_in : L; a.nldL = a.fuse == 1 ? a.fr_in : P.rl;
    a.ldR = a.fuse == 2 ? a.f_in : R; a.nldR = a.fuse == 2 ? a.fr_in : P.rr;
    a.rl = P.rl; a.rr = P.rr; a.mono = P.mono; a.nB = P.nB;
    a.F = P.F; a.fLL = P.fLL; a.fRR = P.fRR; a.fRow1 = P.fRow1; a.fL = P.fL; a.fR = P.fR; a.fRow2 = P.fRow2;
    a.warp_doubles = P.warp_doubles;
    a.tab = P.d_tab; a.tab_entries = P.tab_entries;
    a.partial = partial; a.partial_stride = P.partial_stride;
    const int rc = dispatch(OP_LAUNCH, P, &a, st, nullptr);
    if (rc) return rc;
    if (mid) BSDE_CUDA_OK(cudaEventRecord(mid, st));
    const int red_blocks = (int)((P.partial_stride + 31) / 32);
    PeerExchange ex;
    if (px && px->world > 1) {
        if (red_blocks * moment_groups > kXBlocks || red_blocks * moment_groups * 32 > kXSlots)
            return fail_msg(2, "assembly: %d x %d reduce blocks exceed the peer-exchange buffer", red_blocks, moment_groups);
        ex = *px;
    }
    if (moment_groups < 1 || moment_groups > kRedGroups) return fail_msg(2, "assembly: %d moment groups", moment_groups);
    if (g_pdl_launch) {
        BSDE_CUDA_OK(launch_chain(k_moment_reduce, dim3(red_blocks, moment_groups), dim3(32 * kRedWarps), 0, st, true, partial,
                                  P.grid_x, P.partial_stride, (int)P.partial_stride, moments, ex));
        return 0;
    }
    k_moment_reduce<<<dim3(red_blocks, moment_groups), 32 * kRedWarps, 0, st>>>(partial, P.grid_x, P.partial_stride,
                                                                               (int)P.partial_stride, moments, ex);
    BSDE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace bsde
