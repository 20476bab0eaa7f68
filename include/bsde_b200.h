/* bsde_b200 — C ABI of the B200 (sm_100a) tensor-train regression backend.
 *
 * Drop-in boundary for the sample-batched TT regression of antoine311200/BSDE-tensor-networks.  Every entry
 * point below replaces one Python function of the reference (paths relative to src/bsde_solver/ of the
 * reference); the reference has no FFI of its own — its "GPU backend" is `xp = cupy` — so these are the symbols a
 * ctypes binding inside the reference binds (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; bsde_last_error() returns the message
 *     (thread-local, valid until the next failing call on that thread);
 *   - pointers suffixed _dev are CUDA device pointers on the context's device, _host are host pointers;
 *     the library never frees caller memory;
 *   - all floating point data is IEEE float64; sample counts are int64_t;
 *   - device arrays are SAMPLE-CONTIGUOUS:  X [d][N],  PHI [d][p][N],  Z [d][N],  paths [steps+1][d][N];
 *   - a tensor train travels as `ranks` (d+1 ints, ranks[0] = ranks[d] = 1) plus `cores_host`, the cores
 *     concatenated in order, core j being (ranks[j], p, ranks[j+1]) in C order — exactly the memory of the
 *     reference's TensorCore with indices (r_j, m_{j+1}, r_{j+1}) (core/tensor/tensor_train.py:21-25);
 *   - one context per process and GPU; calls on one context are not re-entrant; work is issued on the
 *     context's stream and the calls that return host data synchronise it.
 */
#ifndef BSDE_B200_H
#define BSDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BSDE_API __attribute__((visibility("default")))
#else
#define BSDE_API
#endif

typedef struct bsde_ctx bsde_ctx;

/* basis_kind */
#define BSDE_BASIS_PHI 0      /* explicit basis values PHI[d][p][N] (the reference's `phis` lists)        */
#define BSDE_BASIS_POLY 1     /* monomials x^0..x^(p-1) evaluated on the fly    (core/calculus/basis.py:15-27) */
#define BSDE_BASIS_LEGENDRE 2 /* Legendre by polyval on scipy's coefficients     (core/calculus/basis.py:29-44) */

/* solver_id — core/optimisation/solve.py:6-19 */
#define BSDE_SOLVER_LSTSQ 0
#define BSDE_SOLVER_LU 1
#define BSDE_SOLVER_QR 2
#define BSDE_SOLVER_SOLVE 3

/* model_id — bsde.py */
#define BSDE_MODEL_BLACKSCHOLES 0 /* params = {r, sigma}   bsde.py:34-55  */
#define BSDE_MODEL_HJB 1          /* params = {sigma}      bsde.py:80-102 */

BSDE_API const char* bsde_last_error(void);
BSDE_API int bsde_version(void);

/* ---- context ------------------------------------------------------------------------------------------- */
BSDE_API int bsde_ctx_create(int device, bsde_ctx** out);
BSDE_API int bsde_ctx_destroy(bsde_ctx* ctx);
/* run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); NULL = the context's own */
BSDE_API int bsde_ctx_set_stream(bsde_ctx* ctx, void* cuda_stream);
BSDE_API int bsde_ctx_sync(bsde_ctx* ctx);
/* kernels launched by this context so far (bench.py's gpu_launches) */
BSDE_API int bsde_ctx_launch_count(bsde_ctx* ctx, int64_t* out);
/* options: "use_graph" (CUDA graph per ALS sweep, default 1); "profile" (1 = time every eager kernel launch with
 * CUDA events by class and disable graphs; setting it resets the counters); measurement switches with the product path
 * as default, listed in INTEGRATION.md §4: "fuse_interface", "fast_solve", "gate_tier2", "svd_onesided", "red_groups",
 * "r4_full", "asm_generic", "solve_stop" */
BSDE_API int bsde_ctx_set_option(bsde_ctx* ctx, const char* name, int64_t value);
/* stats: "<class>_ms" / "<class>_n" for class in assembly, reduce, allreduce, solve, interface (device time and
 * launch count accumulated while "profile" is on); "solve_phase0".."solve_phase3" (SM cycles spent by the solve
 * kernel in moment expansion / solve / QR / factor push, accumulated while "profile" is on); "sm_count";
 * "workspace_bytes" */
BSDE_API int bsde_ctx_get_stat(bsde_ctx* ctx, const char* name, double* out);

/* sample-sharded data parallelism: one process per GPU, NCCL all-reduce of the per-core moments.
 * bsde_comm_unique_id fills 128 bytes on rank 0; broadcast them (torch.distributed / MPI) and call bsde_comm_init. */
BSDE_API int bsde_comm_unique_id(void* out128);
BSDE_API int bsde_comm_init(bsde_ctx* ctx, const void* id128, int rank, int world);

/* Optional, after bsde_comm_init on a single NVLink node (world <= 8): fuse the all-reduce of the per-core moments into
 * the reduction kernel over peer memory instead of calling NCCL.  bsde_comm_p2p_export fills 128 bytes (two CUDA IPC
 * handles) per rank; gather them from all ranks in rank order (world x 128 bytes) and call bsde_comm_p2p_open with
 * enable = 1 on every rank (all ranks must take the same decision; enable = 0 keeps NCCL). */
BSDE_API int bsde_comm_p2p_export(bsde_ctx* ctx, void* out128);
BSDE_API int bsde_comm_p2p_open(bsde_ctx* ctx, const void* all_handles, int enable);

/* plain device memory helpers so a C caller needs nothing but this library */
BSDE_API int bsde_malloc(bsde_ctx* ctx, int64_t bytes, void** out_dev);
BSDE_API int bsde_free(bsde_ctx* ctx, void* ptr_dev);
BSDE_API int bsde_memcpy_h2d(bsde_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes);
BSDE_API int bsde_memcpy_d2h(bsde_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);
/* (rows, cols) C-order -> (cols, rows) C-order on the device, e.g. X (N,d) -> X [d][N] */
BSDE_API int bsde_transpose(bsde_ctx* ctx, const double* src_dev, int64_t rows, int64_t cols, double* dst_dev);

/* ---- the hot path: ALS   (replaces ALS, core/optimisation/als.py:26-96) ---------------------------------
 * inputs_dev : X [d][N] (POLY/LEGENDRE) or PHI [d][p][N] (PHI)
 * coef_host  : LEGENDRE only: [3][p][p] polyval coefficients (derivative order, function, descending power), else NULL
 * cores_host : in  = the cores after `tt.randomize()` (als.py:37; drawn by the caller from its own RNG stream)
 *              out = the cores after n_iter sweeps
 * info_host  : optional int[4]: {cholesky solves, jacobi-svd solves, lu solves, minimum numerical rank}
 */
BSDE_API int bsde_fit_als(bsde_ctx* ctx, int basis_kind, int d, int p, int64_t N, const double* inputs_dev,
                 const double* coef_host, const double* y_dev, int n_iter, const int* ranks, int solver_id,
                 double* cores_host, int* info_host);
/* same with host buffers (inputs_host in the device layouts above); copies are part of the call */
BSDE_API int bsde_fit_als_host(bsde_ctx* ctx, int basis_kind, int d, int p, int64_t N, const double* inputs_host,
                      const double* coef_host, const double* y_host, int n_iter, const int* ranks, int solver_id,
                      double* cores_host, int* info_host);

/* same, with one host block per asset: asset_ptrs_host[j] points at the N values x_j (POLY / LEGENDRE) or at asset j's p
 * basis rows [p][N] (PHI) — the reference's `phis` list is d separate (N, p) F-ordered arrays (core/calculus/basis.py:21),
 * whose transposes are exactly such blocks, so the list is consumed where it lies, without a stacked host copy */
BSDE_API int bsde_fit_als_host_assets(bsde_ctx* ctx, int basis_kind, int d, int p, int64_t N, const double* const* asset_ptrs_host,
                             const double* coef_host, const double* y_host, int n_iter, const int* ranks, int solver_id,
                             double* cores_host, int* info_host);

/* ---- SALSA   (replaces SALSA, core/optimisation/als.py:99-323) -------------------------------------------
 * ranks          : in = current bond ranks of cores_host; out = ranks after adaptation (d+1 ints)
 * init_ranks     : the ranks the TensorTrain object was constructed with (the reference's stale tt.ranks, als.py:122-123)
 * cores_host     : in/out, cores concatenated for the ranks they hold; cores_capacity = doubles available, at least
 *                  sum_j cap_j * p * cap_{j+1} with cap_k = max(ranks[k], min(init_ranks[k-1]*p, max_rank, 8))
 * state_out_host : optional double[8]: {eta, omega, min_sv, rank growth stopped by the kernel limit of 8 (0/1),
 *                  cholesky solves, jacobi-svd solves, lu solves, unused}
 */
BSDE_API int bsde_fit_salsa(bsde_ctx* ctx, int basis_kind, int d, int p, int64_t N, const double* inputs_dev,
                   const double* coef_host, const double* y_dev, int n_iter, int* ranks, const int* init_ranks,
                   double omega, double freq_omega, double min_sv, int max_rank, int solver_id, int do_reg,
                   double* cores_host, int64_t cores_capacity, double* state_out_host);

/* ---- evaluation and gradient   (replace fast_contract utils.py:19-27, multi_derivative derivative.py:43-85) */
BSDE_API int bsde_tt_eval(bsde_ctx* ctx, int basis_kind, int d, int p, int64_t N, const double* inputs_dev,
                 const double* coef_host, const int* ranks, const double* cores_host, double* y_out_dev);
/* dphi_dev: PHI mode only (derivative values [d][p][N]); z_out_dev: [d][N] */
BSDE_API int bsde_tt_grad(bsde_ctx* ctx, int basis_kind, int d, int p, int64_t N, const double* inputs_dev,
                 const double* dphi_dev, const double* coef_host, const int* ranks, const double* cores_host,
                 double* z_out_dev);

/* Hessian of the train at every sample (replaces hessian, core/calculus/hessian.py:6-26): h_out_dev [d][d][N], both
 * triangles.  dphi_dev / d2phi_dev: PHI mode only (first / second derivative values [d][p][N]).  The reference contracts
 * the whole network once per index pair, O(d^3) work; here one right-interface pass and one row kernel per asset, O(d^2). */
BSDE_API int bsde_tt_hessian(bsde_ctx* ctx, int basis_kind, int d, int p, int64_t N, const double* inputs_dev, const double* dphi_dev,
                    const double* d2phi_dev, const double* coef_host, const int* ranks, const double* cores_host,
                    double* h_out_dev);

/* ---- paths and targets   (replace generate_trajectories stochastic/path.py:6-34 and the target lines of
 *      src/scripts/pipeline.py:129-133, src/scripts/pipeline_explicit.py:97-102) ---------------------------- */
/* x_out_dev [steps+1][d][N]; xi_dev [steps+1][d][N] is read when replay != 0 (host-supplied increments),
 * otherwise increments come from Philox4x32-10 keyed by (seed; sample_offset + sample, step, asset) and are
 * written to xi_dev when it is non-NULL.  x0_host [d] is the common starting point of all samples (every reference script
 * tiles one X0); x0_host == NULL: the caller has filled slab 0 of x_out_dev ([d][N]) with one starting point per sample
 * (path.py:6 takes a general (N, d) X0). */
BSDE_API int bsde_paths_simulate(bsde_ctx* ctx, int model_id, const double* params, int d, int64_t N, int steps, double T,
                        const double* x0_host, uint64_t seed, int64_t sample_offset, int replay, double* xi_dev,
                        double* x_out_dev);
BSDE_API int bsde_terminal(bsde_ctx* ctx, int model_id, const double* params, int d, int64_t N, const double* x_dev,
                  double* g_out_dev);
/* out = h(x, y, z) * dt + y_next - sqrt(dt) * sum_j z_j xi_j   (last term only when xi_dev != NULL);
 * z = grad_dev, multiplied by the model's diagonal sigma(x) first when apply_sigma != 0 */
BSDE_API int bsde_target(bsde_ctx* ctx, int model_id, const double* params, int d, int64_t N, const double* x_dev,
                const double* y_dev, const double* y_next_dev, const double* grad_dev, const double* xi_dev,
                double dt, int apply_sigma, double* out_dev);
/* mean over samples of a device vector (deterministic) */
BSDE_API int bsde_mean(bsde_ctx* ctx, const double* v_dev, int64_t N, double* out_host);

/* ---- verification hooks --------------------------------------------------------------------------------- */
/* Normal equations of core j for the given train (interfaces built from scratch): A (n*n) and b (n) to host. */
BSDE_API int bsde_normal_eq(bsde_ctx* ctx, int basis_kind, int d, int p, int64_t N, const double* inputs_dev,
                   const double* coef_host, const double* y_dev, const int* ranks, const double* cores_host, int j,
                   double* A_host, double* b_host);
/* solver(A, b, method) of solve.py on the device for an arbitrary symmetric system (n <= 100) */
BSDE_API int bsde_solve(bsde_ctx* ctx, int n, const double* A_host, const double* b_host, int solver_id, double* x_host,
               int* info_host);
/* TensorTrain.orthonormalize (tensor_train.py:34-87) on the device; mode 0 = left, 1 = right */
BSDE_API int bsde_orthonormalize(bsde_ctx* ctx, int d, int p, const int* ranks, double* cores_host, int mode, int start, int stop);
/* the same for a train whose cores have their own basis sizes shape[j] (TensorTrain(shape, ranks) accepts a ragged
 * shape, tensor_train.py:18-32); cores_host holds core j as (ranks[j], shape[j], ranks[j+1]) */
BSDE_API int bsde_orthonormalize_shape(bsde_ctx* ctx, int d, const int* shape, const int* ranks, double* cores_host, int mode,
                              int start, int stop);
/* One SVD move of SALSA's sweep on a pair of host cores (teacher-forced parity of als.py:240-277 and of rank_adaptation,
 * als.py:185-211).  side 0 = the move on the bond LEFT of core j: core_a = core_{j-1} (ra, p, k), core_b = core_j (k, p, rb);
 * SVD of left_unfold(core_a) = U S Vt, core_a <- U, core_b <- (S * Vt) @ right_unfold(core_b), one more rank when
 * `expand` (the host's expand_rank && k < max_ranks[j-1]) and S[-1] > min_sv.  side 1 = the move on the bond RIGHT of
 * core j: core_a = core_j (ra, p, k), core_b = core_{j+1} (k, p, rb); core_a <- U S, core_b <- Vt @ right_unfold(core_b).
 * Both buffers must have room for rank k + 1.  reg_out_host (>= k + 1 doubles, filled when do_reg): gamma resp. theta =
 * nan_to_num(1 / max(S, min_sv)).  rank_out: the bond rank afterwards; adapted_out: 1 when it grew. */
BSDE_API int bsde_salsa_move(bsde_ctx* ctx, int side, int ra, int p, int k, int rb, double min_sv, int expand, int do_reg,
                    double* core_a_host, double* core_b_host, double* reg_out_host, int* rank_out, int* adapted_out);
/* basis values / derivatives as the fused kernels compute them: out_dev [p][N] for one input vector x_dev [N] */
BSDE_API int bsde_basis_eval(bsde_ctx* ctx, int basis_kind, int p, int64_t N, const double* x_dev, const double* coef_host,
                    int deriv, double* out_dev);

#ifdef __cplusplus
}
#endif
#endif /* BSDE_B200_H */
