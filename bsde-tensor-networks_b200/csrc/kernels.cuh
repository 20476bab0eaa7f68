// Host-callable launchers of the kernels (internal to the library; the public surface is include/bsde_b200.h).
#pragma once
#include "common.cuh"
#include <vector>

namespace bsde {

// ---- interface.cu --------------------------------------------------------------------------------
int launch_interface_left(const Inputs& in, int j, const double* core, int r_in, int r_out, const double* Lin,
                          double* Lout, cudaStream_t st);
int launch_interface_right(const Inputs& in, int j, const double* core, int r_out, int r_in, const double* Rin,
                           double* Rout, cudaStream_t st);
int launch_tt_eval(const Inputs& in, const double* cores, const int64_t* core_off, const int* ranks_dev, int max_rank,
                   double* out, cudaStream_t st);
int launch_grad_step(const Inputs& in, int i, const double* core, int rl, int rr, const double* Lin, const double* Rin,
                     double* Lout, double* z, cudaStream_t st);
int launch_hessian_row(const Inputs& in, int i, const double* cores, const int64_t* core_off, const int* ranks_dev,
                       const double* const* rbond_dev, const double* Lin, int max_rank, double* out, cudaStream_t st);
int residual_grid(int64_t N);
int launch_residual(const Inputs& in, int j, const double* core, int rl, int rr, const double* Lin, const double* Rin,
                    const double* y, double* pred_out, double* partial, double* out_sum, cudaStream_t st);
int launch_sum_partials(const double* partial, int n_partials, int64_t stride, int n_elems, double* out, cudaStream_t st);

// ---- assembly.cu ---------------------------------------------------------------------------------
// Plan of the moment GEMMs for one core shape (r_l, p, r_r, monomial?) — see assembly.cu.
struct AsmPlan {
    int rl = 0, p = 0, rr = 0, mono = 0;
    int nLL = 0, nRR = 0, nB = 0;
    int F = 0;                                    // staged per-sample functions
    int fLL = 0, fRR = 0, fRow1 = 0, fL = 0, fR = 0, fRow2 = 0;   // first index of each family (0 = ZERO)
    int rb_class = -1;                            // >= 0: register-blocked kernel of that tile-count class; -1: generic
    int n_jobs = 0, NJ = 0, n_groups = 0;         // 8x8 output tiles; tiles per CTA; blockIdx.y groups
    int n_mom = 0;                                // canonical moments: nB*nLL*nRR + p*rl*rr + 1
    int n = 0;                                    // unknowns r_l*p*r_r
    int grid_x = 0;                               // CTAs along the sample axis (= number of partial rows)
    int64_t partial_stride = 0;                   // doubles per partial row = tiles*64 + 1
    int warp_doubles = 0;                         // shared-memory doubles per warp
    int warps = 4;                                // warps per CTA of the kernel this plan launches
    int tab_entries = 0;
    int core_off = 0;                             // shared-memory offset (doubles) of the fused update's neighbouring core
    size_t smem_bytes = 0;
    uint16_t* d_tab = nullptr;                    // fragment-source table (layout depends on the variant)
    int32_t* d_gather = nullptr;                  // [n_mom] canonical moment -> slot in a partial row
    uint16_t* d_expand = nullptr;                 // [n*n + n] entry of A (row-major) / b -> slot of a partial row
    std::vector<int32_t> h_gather;
};

// One-shot all-reduce of the assembly sums over NVLink peer memory, fused into k_moment_reduce (assembly.cu).
// Every rank owns an exchange buffer that all ranks of the node map through CUDA IPC.
constexpr int kXSlots = 4096;     // doubles per (parity, source rank)
constexpr int kXBlocks = 160;     // reduce-kernel blocks (32 slots each) that can take part
constexpr int kXMaxWorld = 8;
struct PeerExchange {
    int world = 1, rank = 0;
    double* data[kXMaxWorld] = {};                 // data[p]  : rank p's buffer  [2][world][kXSlots], empty slots = all-ones bits
    unsigned long long* flags[kXMaxWorld] = {};    // (round-1 protocol; the data words are their own flags now)
    unsigned long long* step = nullptr;            // local    : [kXBlocks] exchanges done by each block (parity of the buffer)
    int* error = nullptr;                          // local    : set to 1 when a wait timed out (the sums are then NaN)
    long long timeout_cycles = 20000000000ll;      // ~10 s of SM clock; option "px_timeout_ms"
};

// Interface update fused into the assembly of the NEXT micro-step of an ALS half-sweep: the kernel forms the interface
// on the side the sweep comes from out of the neighbouring bond, core and asset instead of reading it, and writes it to
// `out` (it is needed again by the later micro-steps).  Bit-identical to k_interface_left / k_interface_right.
struct FusedInterface {
    int side = 0;                 // 0 none; 1 left: L_j from core j-1; 2 right: R_j from core j+1
    int j = 0;                    // asset of the neighbouring core
    int r_in = 1;                 // components of `in`
    const double* core = nullptr; // neighbouring core, (r_in, p, r_l) for side 1, (r_r, p, r_in) for side 2
    const double* in = nullptr;   // neighbouring bond, null = ones(1)
    double* out = nullptr;        // receives the interface (r_l resp. r_r components)
};

int asm_plan_build(AsmPlan& plan, int rl, int p, int rr, int mono, int64_t N, int sm_count, int force_generic = 0);
void asm_plan_free(AsmPlan& plan);
bool asm_fuse_supported(const AsmPlan& plan, const Inputs& in, const FusedInterface& fi);
// Launches the assembly kernel and the partial reduction; `moments` (moment_groups rows of partial_stride doubles, device)
// receives [M1 (nB,nLL,nRR) | M2 (p,rl,rr) | y.y] as moment_groups partial sums (MicroSolveArgs::moment_rows adds them).
constexpr int kRedGroups = 6;
extern int g_asm_r4_full;
int launch_assembly(const AsmPlan& plan, const Inputs& in, int j, const double* L, const double* R, const double* y,
                    double* partial, double* moments, cudaStream_t st, cudaEvent_t mid = nullptr,
                    const PeerExchange* px = nullptr, const FusedInterface* fi = nullptr, int moment_groups = 1);

// ---- dense.cu ------------------------------------------------------------------------------------
enum SolverId : int { SOLVER_LSTSQ = 0, SOLVER_LU = 1, SOLVER_QR = 2, SOLVER_SOLVE = 3 };

struct MicroSolveArgs {
    const double* moments;   // summed partial-row slots of the assembly kernel (AsmPlan::partial_stride of them)
    int n_slots = 0;
    const uint16_t* expand_map = nullptr;   // AsmPlan::d_expand
    int rl, p, rr, mono, nLL, nRR, nB;
    int direction;           // +1: left half-sweep (QR, push S right); -1: right half-sweep (push S^T left); 0: no QR
    int solver;
    double* core_j;          // (rl,p,rr) out
    double* core_nb;         // neighbour core: j+1 (rr, p, r_nb) for +1 ; j-1 (r_nb, p, rl) for -1
    int r_nb;                // the neighbour's far rank
    double ridge;            // added to the diagonal, 0 for ALS
    const double* reg_diag;  // optional n-vector added to the diagonal, may be null
    // SALSA (als.py:143-174): A += eta^2 (Gamma^2 (x) I (x) I) [reg_left] + eta^2 (Theta-term, flat layout) [reg_right]
    // + eta I, with eta read from device memory; the two regulariser terms only when gamma AND theta are given
    const double* eta_ptr = nullptr;
    const double* gamma = nullptr;
    const double* theta = nullptr;
    int reg_left = 0, reg_right = 0;
    long long* clk = nullptr;           // optional: 4 phase-cycle accumulators (expand, solve, QR, push)
    double* dbg_A;           // optional n*n + n (A | b) dump for parity tests
    int* info;               // accumulated: [0] cholesky solves, [1] jacobi-svd solves, [2] lu solves, [3] min numerical rank
    int moment_rows = 1;                // rows of `moments` to add (fixed order) while staging; row pitch = n_slots
    int tier2 = 1;                      // second-tier acceptance test of k_micro_solve_fast (option "gate_tier2")
    int force_onesided = 0;             // A/B measurement: symmetric systems also take the one-sided Jacobi SVD (option "svd_onesided")
    int stop_after = 0;                 // measurement only: k_micro_solve_fast returns after phase 1 (load), 2 (factor), 3 (solve)
    int* fast_flag = nullptr;           // device int: enables k_micro_solve_fast (written 0 = done, 1 = fallback needed)
    int* hint = nullptr;                // device int of this core: != 0 = its last solve truncated (k_micro_solve_fast goes straight to vh_lstsq)
    const double* A_direct = nullptr;   // explicit n x n system (then rl = n, p = rr = 1 and moments is unused)
    const double* b_direct = nullptr;
    double* big_ws = nullptr;           // device workspace of micro_solve_workspace_bytes() for cores beyond the shared-memory solver
};
// Bytes of global-memory workspace k_micro_solve needs for a core of n unknowns whose moments fill n_slots slots
// (0: the system fits the shared memory of the one CTA).  The caller provides it as MicroSolveArgs::big_ws.
size_t micro_solve_workspace_bytes(int n, int n_slots);
// *extra_launches: kernels launched in front of k_micro_solve (the fast kernel, see dense.cu)
int launch_micro_solve(const MicroSolveArgs& a, cudaStream_t st, int* extra_launches = nullptr);
// pvec_dev: optional per-core basis sizes (device, d ints) for trains with a ragged `shape` (tensor_train.py:18-32)
int launch_orthonormalize_right(double* cores, const int64_t* core_off_dev, const int* ranks_dev, int p, int d,
                                int start, int stop, cudaStream_t st, const int* pvec_dev = nullptr);
int launch_orthonormalize_left(double* cores, const int64_t* core_off_dev, const int* ranks_dev, int p, int d,
                               int start, int stop, cudaStream_t st, const int* pvec_dev = nullptr);


// ---- model.cu ------------------------------------------------------------------------------------
int launch_paths(int model, double sigma0, double sqrt_dt, int d, int64_t N, int steps, const double* x0_dev,
                 uint64_t seed, int64_t sample_offset, int replay, double* xi, double* x, cudaStream_t st);
int launch_terminal(int model, int d, int64_t N, const double* x, double* out, cudaStream_t st);
int launch_target(int model, double r, double sigma0, double dt, int d, int64_t N, const double* x, const double* y,
                  const double* y_next, const double* grad, const double* xi, int apply_sigma, double* out,
                  cudaStream_t st);
int launch_transpose(const double* src, int64_t rows, int64_t cols, double* dst, cudaStream_t st);
int launch_basis_eval(const Inputs& in, int deriv, double* out, cudaStream_t st);
int mean_grid(int64_t N);
int launch_block_sums(const double* v, int64_t N, double* partial, cudaStream_t st);


// ---- salsa.cu ------------------------------------------------------------------------------------
int launch_salsa_left(double* core_prev, double* core_cur, int rp, int p, int k, int rn, const double* min_sv, int expand,
                      int do_reg, double* gamma, int* rank_out, int* adapted_out, cudaStream_t st);
int launch_salsa_right(double* core_cur, double* core_next, int rl, int p, int k, int rn, const double* min_sv, int do_reg,
                       double* theta, cudaStream_t st);
int launch_salsa_scalars(double* state, const double* res_sum, double n_samples, double freq_omega, int mode, cudaStream_t st);

}  // namespace bsde
