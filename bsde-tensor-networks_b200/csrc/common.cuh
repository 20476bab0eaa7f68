// Shared types and device helpers for the B200 (sm_100a) tensor-train regression kernels.
// All arithmetic is IEEE float64.  Layouts are sample-contiguous ("SoA"):
//   inputs  X    [d][N]        (asset-major)            or  PHI [d][p][N]
//   left    L_j  [r_j][N],  right R_j [r_{j+1}][N]      (interfaces, one slab per core)
//   cores   C_j  [r_j][p][r_{j+1}] row-major — identical to the reference's (r_j, m, r_{j+1}) C order
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace bsde {

constexpr int kMaxP = 8;        // basis functions per asset supported by the fused kernels
constexpr int kMaxR = 16;       // TT rank supported by the fused kernels
constexpr int kMaxN = 512;      // unknowns of one core (r*p*r): every shape the assembly supports (ranks <= 8, p <= 8); cores whose
                                // system does not fit the solver's shared memory (n > ~140) are solved in a global-memory workspace

enum BasisKind : int { BASIS_PHI = 0, BASIS_POLY = 1, BASIS_LEGENDRE = 2 };

// Where the per-sample basis values come from.  BASIS_PHI: explicit values phi[j][m][s] (and dphi for
// gradients), the reference's calling convention (lists of (N,p) arrays).  BASIS_POLY / BASIS_LEGENDRE:
// fused evaluation from x[j][s] (reference: core/calculus/basis.py:15-44).
struct Inputs {
    int kind;
    int p;
    int d;
    int64_t N;              // samples on this GPU (also the leading stride)
    const double* data;     // X [d][N] or PHI [d][p][N]
    const double* ddata;    // dPHI [d][p][N] for BASIS_PHI gradients, else null
    const double* d2data = nullptr;   // d2PHI [d][p][N] for BASIS_PHI Hessians, else null
    const double* coef;     // BASIS_LEGENDRE: [3][p][p] descending-power polyval coefficients (deriv order, i, k)
};

// ---- programmatic dependent launch (the three kernels of a micro-step inside a captured sweep) -------------
// A kernel launched with the programmatic-serialisation attribute becomes resident while its predecessor in the stream
// still runs and blocks in pdl_wait() until that predecessor has completed and its writes are visible; without the
// attribute both instructions do nothing.  Every kernel of the chain waits before it releases its own dependant, so
// at most two consecutive kernels are resident and completion stays transitive along the chain.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_pdl_launch;      // set by api.cu around the launches of a captured sweep (option "pdl")

// kernel<<<grid, block, smem, st>>>(args...), with the programmatic-serialisation attribute when `pdl` is set
template <class... KArgs, class... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- error-free transformations -----------------------------------------------------------------
__device__ __forceinline__ void two_prod(double a, double b, double& p, double& e) {
    p = __dmul_rn(a, b);
    e = __fma_rn(a, b, -p);
}
__device__ __forceinline__ void fast_two_sum(double a, double b, double& s, double& e) {
    s = __dadd_rn(a, b);
    e = __dadd_rn(b, -__dadd_rn(s, -a));      // b - (s - a)
}

// Monomials x^0..x^{PB-1}, each rounded once from a double-double running product, so that they agree
// with numpy's `x**i` (libm pow, basis.py:21) to the last bit in all but vanishingly rare cases.
template <int PB>
__device__ __forceinline__ void monomials_cr(double x, double (&out)[PB]) {
    out[0] = 1.0;
    if (PB > 1) out[1] = x;
    double h = x, l = 0.0;
#pragma unroll
    for (int k = 2; k < PB; k++) {
        double p, e;
        two_prod(h, x, p, e);
        double lo = __fma_rn(l, x, e);
        fast_two_sum(p, lo, h, l);
        out[k] = h;
    }
}

// Fused bases: values (deriv = 0), first (1) or second (2) derivatives at the point x.  Entries m >= in.p are zero.
template <int PB>
__device__ __forceinline__ void basis_from_x(const Inputs& in, double x, int deriv, double (&phi)[PB]) {
    if (in.kind == BASIS_POLY) {
        double mono[PB];
        monomials_cr<PB>(x, mono);
        if (deriv == 0) {
#pragma unroll
            for (int m = 0; m < PB; m++) phi[m] = (m < in.p) ? mono[m] : 0.0;
        } else if (deriv == 1) {        // i * x**(i-1), basis.py:24
            phi[0] = 0.0;
#pragma unroll
            for (int m = 1; m < PB; m++) phi[m] = (m < in.p) ? __dmul_rn((double)m, mono[m - 1]) : 0.0;
        } else {                        // i*(i-1) * x**(i-2), basis.py:27
            phi[0] = 0.0;
            if (PB > 1) phi[1] = 0.0;
#pragma unroll
            for (int m = 2; m < PB; m++) phi[m] = (m < in.p) ? __dmul_rn((double)(m * (m - 1)), mono[m - 2]) : 0.0;
        }
        return;
    }
    // BASIS_LEGENDRE: numpy.polyval order — y = 0; for c in coeffs (descending): y = y*x + c   (basis.py:38-44).
    // Two rows of coefficients are fetched at a time with every load issued before the first use, and the Horner chains of
    // the two rows are interleaved: with the loads inside a run-time loop (round 1) every step of every chain waited for its
    // own coefficient — ~95 us per assembly launch at the C3 shape for 64 multiply-adds per sample (ncu, round 2).  Same
    // operations in the same order: bit-identical.
    const double* cf = in.coef + (int64_t)deriv * in.p * in.p;
    const int p = in.p;
#pragma unroll
    for (int m0 = 0; m0 < PB; m0 += 2) {
        double c0[PB], c1[PB];
#pragma unroll
        for (int k = 0; k < PB; k++) {
            c0[k] = (m0 < p && k < p) ? __ldg(cf + m0 * p + k) : 0.0;
            c1[k] = (m0 + 1 < PB && m0 + 1 < p && k < p) ? __ldg(cf + (m0 + 1) * p + k) : 0.0;
        }
        double y0 = 0.0, y1 = 0.0;
#pragma unroll
        for (int k = 0; k < PB; k++)
            if (k < p) {
                y0 = __dadd_rn(__dmul_rn(y0, x), c0[k]);
                y1 = __dadd_rn(__dmul_rn(y1, x), c1[k]);
            }
        phi[m0] = (m0 < p) ? y0 : 0.0;
        if (m0 + 1 < PB) phi[m0 + 1] = (m0 + 1 < p) ? y1 : 0.0;
    }
}

// Basis values (deriv = 0), first (1) or second (2) derivatives for asset j, sample s.
// Entries m >= in.p are set to zero.
template <int PB>
__device__ __forceinline__ void basis_values(const Inputs& in, int j, int64_t s, int deriv, double (&phi)[PB]) {
    if (in.kind == BASIS_PHI) {
        const double* src = (deriv == 0 ? in.data : (deriv == 1 ? in.ddata : in.d2data)) + ((int64_t)j * in.p) * in.N + s;
#pragma unroll
        for (int m = 0; m < PB; m++) phi[m] = (m < in.p) ? __ldg(src + (int64_t)m * in.N) : 0.0;
        return;
    }
    basis_from_x<PB>(in, __ldg(in.data + (int64_t)j * in.N + s), deriv, phi);
}

#define BSDE_CUDA_OK(expr)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) return ::bsde::fail_cuda(_e, #expr, __FILE__, __LINE__);       \
    } while (0)

int fail_cuda(cudaError_t e, const char* what, const char* file, int line);
int fail_msg(int code, const char* fmt, ...);

}  // namespace bsde
