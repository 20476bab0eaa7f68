// Truncated minimum-norm solve of a symmetric positive semi-definite system inside one CTA — the part of
// solver(A, b, 'lstsq') (reference: core/optimisation/solve.py:11, numpy.linalg.lstsq(rcond=None)) that the Cholesky fast
// path cannot serve: normal equations whose spectrum reaches the truncation threshold eps * n * lambda_max.
//
//   x = sum_{lambda_i > eps n lambda_max}  v_i (v_i . b) / lambda_i          (A = V diag(lambda) V^T)
//
// Method (Veselic-Hari): diagonally pivoted Cholesky A ~ L L^T (L: n x r, stopped when the largest remaining pivot is
// rounding noise), then ONE-sided Jacobi on the columns of L:  L W = G with orthogonal columns g_i, so that
// lambda_i = |g_i|^2, v_i = g_i / |g_i| and
//
//   x = sum_i g_i (g_i . b) / lambda_i^2 .
//
// Why this and not a two-sided Jacobi on A (round 1, ~1.3 ms at n = 64): L^T L is one LR step closer to diagonal than
// L L^T, the sweep count drops from 13-17 to 5-7; a round needs ONE block barrier (column pairs are disjoint) instead of
// two; the rotations come from three dot products that every lane of the pair's lane group forms itself, so there is no
// single-warp rotation phase between barriers; the right singular vectors W are never needed.  Backward error of the
// factorisation is O(eps |A|), the same class as LAPACK's own singular values (measured on ALS systems: the difference
// to numpy.linalg.lstsq is the difference between numpy.linalg.lstsq and numpy.linalg.eigh).
//
// Rows are never swapped: pivot row j_k is only marked, the k-th column of L lives on the rows not yet selected; the
// Jacobi phase does not care about the order of the rows, and x comes out in the original order.
#pragma once
#include <cfloat>
#include <climits>
#include <cuda_runtime.h>

namespace bsde {

#ifdef VH_PROFILE
#define VH_T(idx) do { if (threadIdx.x == 0 && S.clk) { const long long _n = clock64(); S.clk[idx] += _n - vh_t; vh_t = _n; } } while (0)
#else
#define VH_T(idx) do { } while (0)
#endif

struct VhScratch {
    long long* clk;  // VH_PROFILE only: phase cycle counters of thread 0
    double* G;       // r columns of pitch ldg (column-major), ldg even
    int ldg;
    double* d;       // n   remaining diagonal of the Schur complement
    double* lam;     // n   squared column norms
    double* coef;    // n
    double* red;     // 8
    int* flags;      // 8
    unsigned short* tab;   // (m - 1) * m / 2 round-robin pairs, m = n rounded up to even
    long long* keys;       // [2][32] per-warp pivot candidates of the Cholesky phase (double-buffered by step parity)
};

__host__ __device__ inline int vh_ldg(int n) { return ((n + 1) & ~1) + 8; }
__host__ __device__ inline size_t vh_scratch_doubles(int n) {
    const size_t m = (size_t)((n + 1) & ~1);
    return (size_t)n * vh_ldg(n) + 3 * (size_t)n + 8 + 8 + 64 + ((m - 1) * (m / 2) + 3) / 4 + 1;
}

__device__ __forceinline__ VhScratch vh_carve(double* base, int n) {
    VhScratch S;
    S.clk = nullptr;
    S.ldg = vh_ldg(n);
    S.G = base;
    S.d = base + (size_t)n * S.ldg;
    S.lam = S.d + n;
    S.coef = S.lam + n;
    S.red = S.coef + n;
    S.flags = reinterpret_cast<int*>(S.red + 8);
    S.keys = reinterpret_cast<long long*>(S.red + 16);
    S.tab = reinterpret_cast<unsigned short*>(S.red + 16 + 64);
    return S;
}

// 1 / sqrt(x) for normal positive x: hardware seed (MUFU.RSQ64H, ~2^-22) and one third-order step
// y (1 + e/2 + 3 e^2/8), e = 1 - x y^2: relative error ~e^3, i.e. correctly rounded up to the final rounding.
// 50 cycles of dependent latency instead of the 75 of the library routine (no special-case branches).
__device__ __forceinline__ double vh_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

__device__ __forceinline__ void vh_rr_pair(int k, int step, int m, int& p, int& q) {   // round-robin tournament
    if (k == 0) { p = m - 1; q = step; }
    else { p = (step + k) % (m - 1); q = (step - k + m - 1) % (m - 1); }
    if (p > q) { const int t = p; p = q; q = t; }
}

// T    : threads of the CTA (blockDim.x), a multiple of 64
// LP   : lanes per column pair in the Jacobi phase (power of two, <= 32); blockDim.x / LP pairs are rotated concurrently
// UMAX : 16-byte row chunks per lane and column, >= ceil(ceil(n / 2) / LP)
// AFn  : double operator()(int i, int j) — entry of the symmetric matrix
// Returns the numerical rank (>= 0), or -1 when A is not positive semi-definite to working accuracy (nothing written).
// All threads of the CTA must call; uses __syncthreads().
template <int T, int LP, int UMAX, class AFn>
__device__ int vh_lstsq(AFn Aat, int n, const double* __restrict__ bvec, double* __restrict__ xout, VhScratch S, int* sweeps_out) {
    static_assert(T % 64 == 0 && T <= 1024 && (LP & (LP - 1)) == 0 && LP <= 32, "thread layout");
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = T >> 5;
    const int ldg = S.ldg;
    double* G = S.G;
    double* d = S.d;

#ifdef VH_PROFILE
    long long vh_t = clock64();
#endif
    // ---------------- diagonally pivoted Cholesky (left-looking, rows unpermuted) ----------------
    for (int i = tid; i < n; i += T) d[i] = Aat(i, i);
    for (int e = tid; e < n * ldg; e += T) G[e] = 0.0;
    __syncthreads();
    double dmax0 = 0.0;
    {
        bool bad = false;
        for (int i = lane; i < n; i += 32) { const double v = d[i]; dmax0 = fmax(dmax0, v); bad = bad || !isfinite(v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax0 = fmax(dmax0, __shfl_xor_sync(full, dmax0, o));
        if (__any_sync(full, bad)) {          // non-finite input: the answer is not a number (numpy raises LinAlgError)
            for (int i = tid; i < n; i += T) xout[i] = nan("");
            __syncthreads();
            return 0;
        }
    }
    // pivots at the level of the rounding noise of the Schur complement are not columns of A any more
    const double tolpiv = (DBL_EPSILON / 16.0) * dmax0;
    // threads per row in the column update: TPR lanes (stride RPW inside the warp) share one row
    constexpr int TPR = T / 64 > 32 ? 32 : T / 64;
    constexpr int RPW = 32 / TPR;
    const int rowlocal = lane % RPW, sub = lane / RPW;
    // One 64-bit key per row for the pivot search: the diagonal entry with the row index in its 8 lowest mantissa bits
    // (pivot ties within 2^-44 go to the key order; the pivot VALUE is read back exactly).  The keys of the rows a warp has
    // just updated are reduced by that warp BEFORE the step's barrier (s_key[warp]); after it every warp only combines the
    // nwarps keys, so the dependent chain between two steps is: barrier, one shared-memory load, log2(nwarps) shuffles.
    constexpr int KIT = (2 * LP * UMAX + TPR - 1) / TPR;       // column-update iterations per lane (rows capacity / TPR)
    static_assert(nwarps <= 32, "one key per warp");
    // (systems of more than 256 rows: 10 index bits, ties within 2^-42)
    const long long kmask = n <= 256 ? 0xffll : 0x3ffll;
    auto make_key = [kmask](double v, int i) -> long long {
        return v > 0.0 ? ((__double_as_longlong(v) & ~kmask) | (kmask - (long long)i)) : LLONG_MIN;
    };
    int r = 0;
#ifndef VH_FORCE_LEFT_LOOKING
#define VH_FORCE_LEFT_LOOKING 0     // 1: always the left-looking factorisation (A/B measurement)
#endif
    if (T == 256 && n <= 64 && !VH_FORCE_LEFT_LOOKING) {
        // ---- right-looking, register-resident: the Schur complement lives in 4 x 4 blocks on a 16 x 16 thread grid ----
        // Per step: every warp finds the pivot among the 64 diagonal entries (its own register copy, two warp-wide integer
        // maxima), the 16 threads that hold column jk publish it, ONE barrier, every thread scales its 4 + 4 entries of the
        // column and updates its 16 entries.  ~700 cycles per step against ~1500 of the left-looking form below (dot products over the k finished columns from shared memory,
        // two shuffle reductions, per-warp pivot keys), which stays for other thread counts and n > 64.
        const int ty = tid >> 4, tx = tid & 15;
        double Sr[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int i = 4 * ty + a, j = 4 * tx + b;
                Sr[a][b] = (i < n && j < n) ? Aat(i, j) : 0.0;
            }
        // the pivot column of the step, double-buffered by step parity (S.coef and S.lam are free until the final phase), so
        // that one barrier per step is enough; every warp keeps its own copy of the diagonal (lane l: rows l, l + 32) and
        // updates it from the published column with the very operations of the diagonal blocks — no second exchange
        double* colbuf[2] = {S.coef, S.lam};
        unsigned long long sel = 0ull;                 // rows already selected
        double dd0 = lane < n ? d[lane] : -DBL_MAX, dd1 = lane + 32 < n ? d[lane + 32] : -DBL_MAX;
        for (int k = 0; k < n; k++) {
            VH_T(8);
            // pivot = largest remaining diagonal entry (ties: lowest row).  The 64-bit keys (value bits with the row in the
            // lowest 8) are compared as two 32-bit halves, each with one warp-wide integer maximum (REDUX)
            int jk;
            {
                const long long k0 = make_key(dd0, lane), k1 = make_key(dd1, lane + 32);
                const long long km = k0 > k1 ? k0 : k1;
                if (__all_sync(full, km == LLONG_MIN)) break;                 // nothing positive left (uniform over the CTA)
                const unsigned hi = km == LLONG_MIN ? 0u : (unsigned)(km >> 32), lo = (unsigned)km;
                const unsigned himax = __reduce_max_sync(full, hi);
                const unsigned lomax = __reduce_max_sync(full, (km != LLONG_MIN && hi == himax) ? lo : 0u);
                jk = 255 - (int)(lomax & 0xffu);
            }
            VH_T(9);
            double* cb_ = colbuf[k & 1];
            if (tx == (jk >> 2)) {
                const int cb = jk & 3;
#pragma unroll
                for (int a = 0; a < 4; a++)
                    if (4 * ty + a < n) cb_[4 * ty + a] = cb == 0 ? Sr[a][0] : (cb == 1 ? Sr[a][1] : (cb == 2 ? Sr[a][2] : Sr[a][3]));
            }
            __syncthreads();
            VH_T(10);
            const double bv = cb_[jk];                 // the pivot itself: entry jk of its own column
            if (!(bv > tolpiv)) break;
            const double inv = vh_rsqrt(bv), lkk = bv * inv;
            sel |= 1ull << jk;
            double li[4], lj[4];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int i = 4 * ty + a, j = 4 * tx + a;
                li[a] = (i < n && !((sel >> i) & 1ull)) ? cb_[i] * inv : 0.0;
                lj[a] = (j < n && !((sel >> j) & 1ull)) ? cb_[j] * inv : 0.0;
            }
            {   // this warp's copy of the diagonal
                const double l0 = lane < n ? cb_[lane] * inv : 0.0, l1 = lane + 32 < n ? cb_[lane + 32] * inv : 0.0;
                dd0 = (lane >= n || ((sel >> lane) & 1ull)) ? -DBL_MAX : fma(-l0, l0, dd0);
                dd1 = (lane + 32 >= n || ((sel >> (lane + 32)) & 1ull)) ? -DBL_MAX : fma(-l1, l1, dd1);
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) Sr[a][b] = fma(-li[a], lj[b], Sr[a][b]);
            VH_T(11);
            if (tx == 0) {
                double* colk = G + (size_t)k * ldg;
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const int i = 4 * ty + a;
                    if (i < n) colk[i] = (i == jk) ? lkk : li[a];
                }
            }
            VH_T(12);
            r = k + 1;
        }
        __syncthreads();
        if (warp == 0) {                               // what the final phase and the caller read: selected rows marked in d
            if (lane < n) d[lane] = dd0;
            if (lane + 32 < n) d[lane + 32] = dd1;
        }
        if (r < n) {       // positive semi-definite to working accuracy?  what is left of the Schur complement must be noise
            if (tid == 0) S.flags[2] = 0;
            __syncthreads();
            bool bad = false;
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int i = 4 * ty + a, j = 4 * tx + b;
                    if (i < n && j <= i && !((sel >> i) & 1ull) && !((sel >> j) & 1ull) && !(fabs(Sr[a][b]) <= 1e-10 * dmax0)) bad = true;
                }
            if (bad) S.flags[2] = 1;
            __syncthreads();
            if (S.flags[2]) return -1;
        }
    } else {
    {   // keys of the initial diagonal
        long long key = LLONG_MIN;
        for (int i0 = warp * RPW; i0 < n; i0 += nwarps * RPW) {
            const int i = i0 + rowlocal;
            if (sub == 0 && i < n) { const long long kv = make_key(d[i], i); key = kv > key ? kv : key; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const long long ok = __shfl_xor_sync(full, key, o); key = ok > key ? ok : key; }
        if (lane == 0) S.keys[warp] = key;
    }
    __syncthreads();
    for (int k = 0; k < n; k++) {
        VH_T(8);
        long long key = lane < nwarps ? S.keys[(k & 1) * 32 + lane] : LLONG_MIN;
#pragma unroll
        for (int o = (nwarps > 16 ? 16 : (nwarps > 8 ? 8 : (nwarps > 4 ? 4 : (nwarps > 2 ? 2 : 1)))); o > 0; o >>= 1) {
            const long long ok = __shfl_xor_sync(full, key, o);
            key = ok > key ? ok : key;
        }
        key = __shfl_sync(full, key, 0);
        if (key == LLONG_MIN) break;                   // nothing positive left (uniform over the CTA)
        const int jk = (int)(kmask - (key & kmask));
        const double bv = d[jk];
        if (!(bv > tolpiv)) break;
        VH_T(9);
        const double inv = vh_rsqrt(bv), lkk = bv * inv;
        double* colk = G + (size_t)k * ldg;
        long long mykey = LLONG_MIN;
        for (int i0 = warp * RPW; i0 < n; i0 += nwarps * RPW) {
            const int i = i0 + rowlocal;
            const bool live = i < n && i != jk && d[i] > -DBL_MAX;          // not selected yet
            // row i of the factor times row jk, columns t = sub, sub + TPR, ... < k: all loads issued up front, four chains
            double pa[4] = {0.0, 0.0, 0.0, 0.0};
            const double* gi = G + (live ? i : jk) + (size_t)sub * ldg;
            const double* gj = G + jk + (size_t)sub * ldg;
#pragma unroll
            for (int u = 0; u < KIT; u++) {
                const int t = sub + u * TPR;
                if (t < k) pa[u & 3] = fma(gi[(size_t)u * TPR * ldg], gj[(size_t)u * TPR * ldg], pa[u & 3]);
            }
            double part = (pa[0] + pa[1]) + (pa[2] + pa[3]);
            VH_T(10);
#pragma unroll
            for (int o = RPW; o < 32; o <<= 1) part += __shfl_xor_sync(full, part, o);
            VH_T(11);
            if (sub == 0 && i < n) {
                if (live) {
                    const double v = (Aat(i, jk) - part) * inv;
                    colk[i] = v;
                    const double dn = fma(-v, v, d[i]);
                    d[i] = dn;
                    const long long kv = make_key(dn, i);
                    mykey = kv > mykey ? kv : mykey;
                } else if (i == jk) {
                    colk[i] = lkk;
                    d[i] = -DBL_MAX;                   // selected: never a pivot again, no entries in later columns
                }
            }
        }
        // this warp's best remaining pivot (lanes with sub == 0 hold the rows' keys)
#pragma unroll
        for (int o = RPW >> 1; o > 0; o >>= 1) { const long long ok = __shfl_xor_sync(full, mykey, o); mykey = ok > mykey ? ok : mykey; }
        if (lane == 0) S.keys[((k + 1) & 1) * 32 + warp] = mykey;
        VH_T(12);
        __syncthreads();
        VH_T(13);
        r = k + 1;
    }
    // positive semi-definite to working accuracy?  (Gram matrices always are; an explicit system may not be)
    if (r < n) {
        if (tid == 0) S.flags[2] = 0;
        __syncthreads();
        const int rest = n * n;
        bool bad = false;
        for (int e = tid; e < rest; e += T) {
            const int i = e / n, j = e - i * n;
            if (j > i || !(d[i] > -DBL_MAX) || !(d[j] > -DBL_MAX)) continue;
            double s = Aat(i, j);
            for (int t = 0; t < r; t++) s = fma(-G[(size_t)t * ldg + i], G[(size_t)t * ldg + j], s);
            if (!(fabs(s) <= 1e-10 * dmax0)) bad = true;
        }
        if (bad) S.flags[2] = 1;
        __syncthreads();
        if (S.flags[2]) return -1;
    }
    }
    if (r == 0) {                                      // A = 0: the minimum-norm solution is 0
        for (int i = tid; i < n; i += T) xout[i] = 0.0;
        __syncthreads();
        return 0;
    }

    VH_T(0);
    // ---------------- one-sided Jacobi on the r columns ----------------
    const int m = (r + 1) & ~1, NP = m >> 1;
    constexpr int PG = T / LP;
    const int pg = tid / LP, pl = tid % LP;
    const int nchunk = (n + 1) >> 1;                   // 16-byte chunks per column (rows beyond n are zero)
    const double negl = 1e-2 * DBL_EPSILON * (double)n * dmax0;   // both columns below: dropped whatever they mix into
    const double tol2 = DBL_EPSILON * DBL_EPSILON;               // rotate when cos^2 > tol2
    const double thr2 = 1e-14;                                    // another sweep when some cos^2 > thr2 (cos 1e-7: the
                                                                  // sweep leaves cos^2 ~ 1e-28 behind, quadratic convergence)
    // ... unless BOTH columns of the pair lie below a quarter of the smallest possible truncation cutoff
    // (eps n lambda_max >= eps n dmax0): such columns are dropped whatever their mutual angle (cos^2 <= thr2_sub = 1e-4
    // bounds what any group of them could add up to far below the cutoff), and their entries are rounding noise of the
    // large columns, so their angles never settle — on the heavily rank-deficient systems of small sample sets (the
    // reference's benchmark point N = 2000, r = p = 5: numerical rank 26..66 of 125) a handful of such pairs kept the
    // sweeps going up to the limit of 30 (7 with this rule; same ranks, solutions equal to 1e-11 in the fit seminorm:
    // bench/micro/jacobi_studies/sweeps_r5.py).  They are still rotated like every other pair.
    const double sub_cut = 0.25 * DBL_EPSILON * (double)n * dmax0;
    const double thr2_sub = 1e-4;
    int sweep = 0;
    // Jacobi rotation that orthogonalises two columns with squared norms al, be and inner product ga; w = be - al, g = 2 ga,
    // hyp = sqrt(w^2 + g^2), a = |w| + hyp:  tan = sign(w) g / a, so (cos, sin) = (a, sign(w) g) / sqrt(a^2 + g^2).
    // The normalisation is done to full precision (one reciprocal square root with its correction step: the rotation is
    // orthogonal to rounding), hyp only to the 2^-22 of the hardware seed: the angle is then off by 2^-22 relative, i.e. the
    // rotation leaves 2^-22 of the inner product behind instead of nothing — far below what the next sweep removes, and
    // the second correction step (60 cycles of the dependent chain of every round) is gone.  w and g are first scaled by a
    // power of two near 1 / max(al, be), so that the squares neither overflow nor vanish.
    auto rotation = [](double al, double be, double ga, double& c, double& sn) {
        const int ex = (__double2hiint(fmax(al, be)) >> 20) & 0x7ff;
        const double scale = __hiloint2double((2046 - ex) << 20, 0);      // 2^(1023 - ex)
        const double w = (be - al) * scale, g = 2.0 * ga * scale;
        const double qq = fma(w, w, g * g);
        double seed;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(seed) : "d"(qq));
        const double a = fma(qq, seed, fabs(w));             // |w| + hyp
        const double rn = vh_rsqrt(fma(a, a, g * g));
        c = a * rn;
        sn = copysign(g * rn, g * w);                         // sign(w) g: w = 0 counts as positive
    };
#ifndef VH_FORCE_ROUND_ROBIN
#define VH_FORCE_ROUND_ROBIN 0     // 1: always the round-robin tournament (A/B measurement)
#endif
    if (r > 1 && (r >> 1) <= PG && !VH_FORCE_ROUND_ROBIN) {
        // ---- odd-even (Brent-Luk) ordering, one column per lane group resident in registers ----
        // Group i owns slot 2i+1 for the whole phase (K, in registers).  Even rounds pair it with slot 2i, odd rounds with
        // slot 2i+2; after the rotation the two columns trade places (the one from shared memory stays in the registers,
        // the old K goes to the slot), so that every column travels through the array and meets every other one once in r
        // rounds — the systolic-array ordering.  A round reads ONE column per pair from shared memory and writes one
        // (the round-robin tournament below moves both: 64 KB per round at n = 64, the floor of that formulation), touches
        // only the slot between two neighbouring groups, and needs one barrier.
        const int ng = r >> 1;
        const bool mine = pg < ng;
        double2 kc[UMAX];
        {
            const double2* ck = reinterpret_cast<const double2*>(G + (size_t)(mine ? 2 * pg + 1 : 0) * ldg);
#pragma unroll
            for (int u = 0; u < UMAX; u++) {
                const int c = pl + LP * u;
                kc[u] = (mine && c < nchunk) ? ck[c] : make_double2(0.0, 0.0);
            }
        }
        if (tid == 0) { S.flags[0] = 0; S.flags[1] = 0; }
        __syncthreads();
        // the slot this group works on in even / odd rounds
        const bool act_odd = mine && 2 * pg + 2 < r;
        double2* const cx_even = reinterpret_cast<double2*>(G + (size_t)(mine ? 2 * pg : 0) * ldg);
        double2* const cx_odd = reinterpret_cast<double2*>(G + (size_t)(act_odd ? 2 * pg + 2 : 0) * ldg);
        for (; sweep < 30; sweep++) {
            if (tid == 0) S.flags[(sweep + 1) & 1] = 0;
            for (int t = 0; t < r; t++) {
                const bool odd = (t & 1) != 0;
                const bool act = odd ? act_odd : mine;
                VH_T(1);
                // (every lane runs the same instruction stream with the full shuffle mask — idle groups work on zeros —
                // so that the four groups of a warp never serialise)
                double2* cx = odd ? cx_odd : cx_even;
                double2 xc[UMAX];
#pragma unroll
                for (int u = 0; u < UMAX; u++) {
                    const int c = pl + LP * u;
                    xc[u] = (act && c < nchunk) ? cx[c] : make_double2(0.0, 0.0);
                }
                double al0 = 0.0, be0 = 0.0, ga0 = 0.0, al1 = 0.0, be1 = 0.0, ga1 = 0.0;      // al: x.x, be: k.k
#pragma unroll
                for (int u = 0; u < UMAX; u++) {
                    al0 = fma(xc[u].x, xc[u].x, al0); al1 = fma(xc[u].y, xc[u].y, al1);
                    be0 = fma(kc[u].x, kc[u].x, be0); be1 = fma(kc[u].y, kc[u].y, be1);
                    ga0 = fma(xc[u].x, kc[u].x, ga0); ga1 = fma(xc[u].y, kc[u].y, ga1);
                }
                double al = al0 + al1, be = be0 + be1, ga = ga0 + ga1;
                VH_T(2);
#pragma unroll
                for (int o = LP >> 1; o > 0; o >>= 1) {
                    al += __shfl_xor_sync(full, al, o);
                    be += __shfl_xor_sync(full, be, o);
                    ga += __shfl_xor_sync(full, ga, o);
                }
                VH_T(3);
                // even round: x is the left column, K the right one; odd round: the mirror image.  Rotated columns
                // left' = c left - s right, right' = s left + c right; then the two trade places: the registers keep what was
                // read from the slot, the slot receives the old K.
                double c = 1.0, sn = 0.0;
                if (act && !(al <= negl && be <= negl)) {
                    const double g2 = ga * ga, ab = al * be;
                    if (g2 > ((al <= sub_cut && be <= sub_cut) ? thr2_sub : thr2) * ab && pl == 0) S.flags[sweep & 1] = 1;
                    if (g2 > tol2 * ab) rotation(odd ? be : al, odd ? al : be, ga, c, sn);
                }
                VH_T(4);
                const double sx = odd ? -sn : sn;
#pragma unroll
                for (int u = 0; u < UMAX; u++) {
                    const int cidx = pl + LP * u;
                    const double2 x = xc[u], k = kc[u];
                    // even: x' = c x - s k, k' = s x + c k;  odd: k' = c k - s x, x' = s k + c x
                    const double2 xn = make_double2(fma(c, x.x, -sx * k.x), fma(c, x.y, -sx * k.y));
                    const double2 kn = make_double2(fma(sx, x.x, c * k.x), fma(sx, x.y, c * k.y));
                    if (act) {
                        kc[u] = xn;
                        if (cidx < nchunk) cx[cidx] = kn;
                    }
                }
                VH_T(5);
                __syncthreads();
                VH_T(6);
            }
            const int need = S.flags[sweep & 1];
            __syncthreads();                           // the flag is cleared again at the start of the next sweep
            if (!need) { sweep++; break; }
        }
        if (mine) {
            double2* ck = reinterpret_cast<double2*>(G + (size_t)(2 * pg + 1) * ldg);
#pragma unroll
            for (int u = 0; u < UMAX; u++) {
                const int c = pl + LP * u;
                if (c < nchunk) ck[c] = kc[u];
            }
        }
        __syncthreads();
    } else if (r > 1) {
        // the tournament's pairs, once per solve: tab[step * NP + k] = p | q << 8 (q = 255: padding of an odd r)
        // (more than 254 columns do not fit the 8-bit fields: the pairs are then computed on the fly)
        unsigned short* tab = S.tab;
        const bool use_tab = r <= 254;
        if (use_tab)
            for (int e = tid; e < (m - 1) * NP; e += T) {
                int p, q;
                vh_rr_pair(e % NP, e / NP, m, p, q);
                tab[e] = (unsigned short)(p | ((q < r ? q : 255) << 8));
            }
        if (tid == 0) { S.flags[0] = 0; S.flags[1] = 0; }
        __syncthreads();
        for (; sweep < 30; sweep++) {
            if (tid == 0) S.flags[(sweep + 1) & 1] = 0;
            for (int step = 0; step < m - 1; step++) {
                for (int k0 = 0; k0 < NP; k0 += PG) {
                    const int k = k0 + pg;
                    int p = 0, q = 255;
                    bool valid = false;
                    if (k < NP) {
                        if (use_tab) { const unsigned pq = tab[step * NP + k]; p = pq & 255; q = pq >> 8; valid = q != 255; }
                        else { vh_rr_pair(k, step, m, p, q); valid = q < r; }
                    }
                    VH_T(1);
                    double2 xp[UMAX], xq[UMAX];
                    double2* cp = reinterpret_cast<double2*>(G + (size_t)p * ldg);
                    double2* cq = reinterpret_cast<double2*>(G + (size_t)(valid ? q : 0) * ldg);
#pragma unroll
                    for (int u = 0; u < UMAX; u++) {
                        const int c = pl + LP * u;
                        if (valid && c < nchunk) { xp[u] = cp[c]; xq[u] = cq[c]; }
                        else { xp[u] = make_double2(0.0, 0.0); xq[u] = make_double2(0.0, 0.0); }
                    }
                    double al0 = 0.0, be0 = 0.0, ga0 = 0.0, al1 = 0.0, be1 = 0.0, ga1 = 0.0;      // two chains per sum
#pragma unroll
                    for (int u = 0; u < UMAX; u++) {
                        al0 = fma(xp[u].x, xp[u].x, al0); al1 = fma(xp[u].y, xp[u].y, al1);
                        be0 = fma(xq[u].x, xq[u].x, be0); be1 = fma(xq[u].y, xq[u].y, be1);
                        ga0 = fma(xp[u].x, xq[u].x, ga0); ga1 = fma(xp[u].y, xq[u].y, ga1);
                    }
                    double al = al0 + al1, be = be0 + be1, ga = ga0 + ga1;
                    VH_T(2);
#pragma unroll
                    for (int o = LP >> 1; o > 0; o >>= 1) {
                        al += __shfl_xor_sync(full, al, o);
                        be += __shfl_xor_sync(full, be, o);
                        ga += __shfl_xor_sync(full, ga, o);
                    }
                    VH_T(3);
                    if (!valid || (al <= negl && be <= negl)) continue;       // no shuffles below this line
                    const double g2 = ga * ga, ab = al * be;
                    if (g2 > ((al <= sub_cut && be <= sub_cut) ? thr2_sub : thr2) * ab && pl == 0) S.flags[sweep & 1] = 1;
                    if (g2 > tol2 * ab) {
                        double c, s;
                        rotation(al, be, ga, c, s);
                        VH_T(4);
#pragma unroll
                        for (int u = 0; u < UMAX; u++) {
                            const int cidx = pl + LP * u;
                            if (cidx < nchunk) {
                                const double2 a = xp[u], b = xq[u];
                                cp[cidx] = make_double2(fma(c, a.x, -s * b.x), fma(c, a.y, -s * b.y));
                                cq[cidx] = make_double2(fma(s, a.x, c * b.x), fma(s, a.y, c * b.y));
                            }
                        }
                        VH_T(5);
                    }
                }
                __syncthreads();
                VH_T(6);
            }
            const int need = S.flags[sweep & 1];
            __syncthreads();                           // the flag is cleared again at the start of the next sweep
            if (!need) { sweep++; break; }
        }
    }
    if (sweeps_out && tid == 0) *sweeps_out = sweep;

    // ---------------- eigenvalues, truncation, minimum-norm sum ----------------
    for (int i0 = 0; i0 < r; i0 += PG) {
        const int i = i0 + pg;
        const bool valid = i < r;
        const double2* ci = reinterpret_cast<const double2*>(G + (size_t)(valid ? i : 0) * ldg);
        double ss = 0.0, pb = 0.0;
#pragma unroll
        for (int u = 0; u < UMAX; u++) {
            const int c = pl + LP * u;
            if (valid && c < nchunk) {
                const double2 g = ci[c];
                const double b0 = bvec[2 * c], b1 = (2 * c + 1 < n) ? bvec[2 * c + 1] : 0.0;
                ss = fma(g.x, g.x, ss); ss = fma(g.y, g.y, ss);
                pb = fma(g.x, b0, pb); pb = fma(g.y, b1, pb);
            }
        }
#pragma unroll
        for (int o = LP >> 1; o > 0; o >>= 1) { ss += __shfl_xor_sync(full, ss, o); pb += __shfl_xor_sync(full, pb, o); }
        if (valid && pl == 0) { S.lam[i] = ss; S.coef[i] = pb; }
    }
    __syncthreads();
    double lmax = 0.0;
    for (int i = lane; i < r; i += 32) lmax = fmax(lmax, S.lam[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(full, lmax, o));
    const double cutoff = DBL_EPSILON * (double)n * lmax;
    int rank = 0;
    double lmin = DBL_MAX;                              // smallest eigenvalue that is kept
    for (int i = lane; i < r; i += 32) {
        const double l = S.lam[i];
        if (l > cutoff) { rank++; lmin = fmin(lmin, l); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { rank += __shfl_xor_sync(full, rank, o); lmin = fmin(lmin, __shfl_xor_sync(full, lmin, o)); }
    if (tid == 0) S.red[4] = cutoff > 0.0 ? lmin / cutoff : 0.0;     // margin of the rank decision, for the caller
    __syncthreads();
    for (int i = tid; i < r; i += T) {
        const double l = S.lam[i];
        S.coef[i] = (l > cutoff) ? S.coef[i] / (l * l) : 0.0;
    }
    __syncthreads();
    for (int i0 = warp * RPW; i0 < n; i0 += nwarps * RPW) {
        const int i = i0 + rowlocal;
        double acc = 0.0;
        if (i < n)
            for (int t = sub; t < r; t += TPR) acc = fma(G[(size_t)t * ldg + i], S.coef[t], acc);
#pragma unroll
        for (int o = RPW; o < 32; o <<= 1) acc += __shfl_xor_sync(full, acc, o);
        if (sub == 0 && i < n) xout[i] = acc;
    }
    __syncthreads();
    return rank;
}

}  // namespace bsde
