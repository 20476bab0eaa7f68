// C ABI of the library (include/bsde_b200.h): context, workspaces, and the host-side sequencing of the kernels.
// The only host work per fit is enqueueing launches (one CUDA graph per ALS sweep); all arithmetic runs on the GPU.
#include "../../include/bsde_b200.h"
#include "kernels.cuh"

#include <dlfcn.h>
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace bsde {

static thread_local std::string g_err;

int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_err = buf;
    cudaGetLastError();
    return 1;
}

int fail_msg(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// ---- NCCL, bound at run time so that the library loads (and single-GPU runs work) without it --------------
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId_ { char internal[128]; };
struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId_*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;

static int nccl_load() {
    if (g_nccl.lib) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return fail_msg(3, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
#define BSDE_SYM(field, name)                                                        \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                               \
    if (!g_nccl.field) return fail_msg(3, "NCCL symbol %s missing", name)
    BSDE_SYM(GetUniqueId, "ncclGetUniqueId");
    BSDE_SYM(CommInitRank, "ncclCommInitRank");
    BSDE_SYM(AllReduce, "ncclAllReduce");
    BSDE_SYM(CommDestroy, "ncclCommDestroy");
    BSDE_SYM(GetErrorString, "ncclGetErrorString");
#undef BSDE_SYM
    return 0;
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail_cuda(e, "cudaMalloc(workspace)", __FILE__, __LINE__); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

}  // namespace bsde

using namespace bsde;

struct bsde_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own = nullptr, stream = nullptr;
    int64_t launches = 0;
    int64_t use_graph = 1;
    DevBuf bond, partial, moments, cores, meta, coef, tmp, in_copy, y_copy, scalars, dbg, clk, hints;
    DevBuf solve_ws;              // global-memory workspace of k_micro_solve for cores beyond its shared memory (dense.cu)
    std::map<std::tuple<int, int, int, int, int64_t>, AsmPlan> plans;
    int64_t asm_generic = 0;   // option: force the generic assembly kernel (A/B measurements, tests)
    int64_t solve_stop = 0;       // measurement only (MicroSolveArgs::stop_after)
    int64_t red_groups = 3;       // option: row groups of the moment reduction on a single GPU (1 = one level; measured best: 3)
    int64_t solve_hints = 1;      // option: per-core "last solve truncated" hints (MicroSolveArgs::hint)
    int64_t gate_tier2 = 1;       // option: second-tier acceptance test in k_micro_solve_fast (MicroSolveArgs::tier2)
    int64_t svd_onesided = 0;     // A/B measurement: one-sided Jacobi SVD for symmetric systems too (MicroSolveArgs::force_onesided)
    int64_t fast_solve = 1;       // option: register-resident fast kernel in front of k_micro_solve (dense.cu)
    int64_t fuse_interface = 1;   // option: ALS forms the trailing interface inside the next assembly kernel (als_sweep)
    int64_t pdl = 1;              // option: kernels of a captured sweep launched with programmatic dependent launch (common.cuh)
    int64_t graph_edges = 0, graph_pdl_edges = 0;      // stats: edges of the last captured ALS sweep, programmatic ones among them
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // per-kernel-class device timing (option "profile"): event pairs recorded around eager launches
    int64_t profile = 0;
    enum { P_ASSEMBLY = 0, P_REDUCE, P_ALLREDUCE, P_SOLVE, P_INTERFACE, P_CLASSES };
    double prof_ms[P_CLASSES] = {0, 0, 0, 0, 0};
    int64_t prof_n[P_CLASSES] = {0, 0, 0, 0, 0};
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
    bool capturing = false;
    // host-buffer fits (bsde_fit_als_host): the inputs arrive asset by asset on a copy stream, in the order the right
    // interfaces are built (d-1 .. 1, then asset 0 and y), so that build_right overlaps the tail of the transfer
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> in_ev;      // in_ev[j]: asset j has arrived; in_ev[d]: y has arrived
    int in_ev_count = 0;                 // > 0: the current fit must wait on them
    // fused NVLink all-reduce (bsde_comm_p2p_*): buffers of this rank and the peers' mappings
    PeerExchange px;
    bool px_on = false;
    double* px_data_local = nullptr;
    unsigned long long* px_flags_local = nullptr;
    void* px_opened[2 * kXMaxWorld] = {};
};

namespace {

struct Guard {   // select the context's device for the duration of a call
    int prev = -1;
    explicit Guard(const bsde_ctx* c) { cudaGetDevice(&prev); if (prev != c->device) cudaSetDevice(c->device); else prev = -1; }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define BSDE_TRY(expr)            \
    do {                          \
        const int _rc = (expr);   \
        if (_rc) return _rc;      \
    } while (0)

#define BSDE_CHECK_CTX(c) \
    if (!(c)) return fail_msg(2, "null context")

// Called at the start of a fit only: a plan must never be dropped (and rebuilt with cudaMalloc) between the first sweep
// of a fit and the capture of its sweep graph.
void plans_trim(bsde_ctx* c) {
    if (c->plans.size() <= 48) return;
    for (auto& kv : c->plans) asm_plan_free(kv.second);
    c->plans.clear();
}

int get_plan(bsde_ctx* c, int rl, int p, int rr, int mono, int64_t N, const AsmPlan** out) {
    auto key = std::make_tuple(rl, p, rr, mono, N);
    auto it = c->plans.find(key);
    if (it == c->plans.end()) {
        AsmPlan P;
        BSDE_TRY(asm_plan_build(P, rl, p, rr, mono, N, c->sm_count, (int)c->asm_generic));
        it = c->plans.emplace(key, std::move(P)).first;
    }
    *out = &it->second;
    return 0;
}

// One tensor train on the device: cores at fixed slots of `slot` doubles (capacity for rank growth).
struct Train {
    int d = 0, p = 0;
    std::vector<int> ranks;        // d+1
    std::vector<int> cap;          // d+1 capacity ranks (>= ranks)
    std::vector<int64_t> off;      // d offsets (doubles) into `cores`
    double* cores = nullptr;       // device
    int64_t* off_dev = nullptr;
    int* ranks_dev = nullptr;
    int64_t total = 0;
    int max_rank() const { int m = 1; for (int r : ranks) m = r > m ? r : m; return m; }
    double* core(int j) const { return cores + off[j]; }
    int64_t size(int j) const { return (int64_t)ranks[j] * p * ranks[j + 1]; }
};

int check_ranks(int d, int p, const int* ranks) {
    if (d < 1) return fail_msg(2, "tensor train needs at least one core (d = %d)", d);
    if (p < 1 || p > kMaxP) return fail_msg(2, "basis size p = %d outside 1..%d", p, kMaxP);
    if (!ranks) return fail_msg(2, "ranks is null");
    if (ranks[0] != 1 || ranks[d] != 1) return fail_msg(2, "boundary ranks must be 1 (got %d, %d)", ranks[0], ranks[d]);
    for (int j = 0; j <= d; j++)
        if (ranks[j] < 1 || ranks[j] > kMaxR) return fail_msg(2, "rank %d = %d outside 1..%d", j, ranks[j], kMaxR);
    return 0;
}

// Lay the train out in c->cores / c->meta and upload the packed host cores.
int train_upload(bsde_ctx* c, Train& T, int d, int p, const int* ranks, const int* cap_ranks, const double* cores_host) {
    T.d = d; T.p = p;
    T.ranks.assign(ranks, ranks + d + 1);
    T.cap.assign(cap_ranks ? cap_ranks : ranks, (cap_ranks ? cap_ranks : ranks) + d + 1);
    T.off.resize(d);
    int64_t o = 0;
    for (int j = 0; j < d; j++) { T.off[j] = o; o += (int64_t)T.cap[j] * p * T.cap[j + 1]; }
    T.total = o;
    BSDE_TRY(c->cores.ensure(sizeof(double) * (size_t)o));
    BSDE_TRY(c->meta.ensure(sizeof(int64_t) * d + sizeof(int) * (d + 1) + 64));
    T.cores = c->cores.as<double>();
    T.off_dev = c->meta.as<int64_t>();
    T.ranks_dev = (int*)(T.off_dev + d);
    BSDE_CUDA_OK(cudaMemcpyAsync(T.off_dev, T.off.data(), sizeof(int64_t) * d, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaMemcpyAsync(T.ranks_dev, T.ranks.data(), sizeof(int) * (d + 1), cudaMemcpyHostToDevice, c->stream));
    if (cores_host) {
        bool packed = true;
        for (int j = 0; j <= d; j++) packed = packed && T.cap[j] == T.ranks[j];
        if (packed) {
            BSDE_CUDA_OK(cudaMemcpyAsync(T.cores, cores_host, sizeof(double) * (size_t)o, cudaMemcpyHostToDevice, c->stream));
        } else {
            int64_t ho = 0;
            for (int j = 0; j < d; j++) {
                BSDE_CUDA_OK(cudaMemcpyAsync(T.core(j), cores_host + ho, sizeof(double) * (size_t)T.size(j), cudaMemcpyHostToDevice, c->stream));
                ho += T.size(j);
            }
        }
    }
    // the host vectors above must outlive the async copies
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int train_download(bsde_ctx* c, const Train& T, double* cores_host) {
    int64_t ho = 0;
    for (int j = 0; j < T.d; j++) {
        BSDE_CUDA_OK(cudaMemcpyAsync(cores_host + ho, T.core(j), sizeof(double) * (size_t)T.size(j), cudaMemcpyDeviceToHost, c->stream));
        ho += T.size(j);
    }
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int make_inputs(bsde_ctx* c, Inputs& in, int basis_kind, int d, int p, int64_t N, const double* data, const double* ddata,
                const double* coef_host) {
    if (basis_kind < 0 || basis_kind > 2) return fail_msg(2, "unknown basis_kind %d", basis_kind);
    if (N < 1) return fail_msg(2, "sample count N = %lld must be positive", (long long)N);
    if (!data) return fail_msg(2, "inputs pointer is null");
    in.kind = basis_kind; in.p = p; in.d = d; in.N = N; in.data = data; in.ddata = ddata; in.coef = nullptr;
    if (basis_kind == BASIS_LEGENDRE) {
        if (!coef_host) return fail_msg(2, "Legendre basis needs coef_host ([3][p][p])");
        BSDE_TRY(c->coef.ensure(sizeof(double) * 3 * p * p));
        BSDE_CUDA_OK(cudaMemcpyAsync(c->coef.p, coef_host, sizeof(double) * 3 * p * p, cudaMemcpyHostToDevice, c->stream));
        BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
        in.coef = c->coef.as<double>();
    }
    return 0;
}

// bond[k], k = 1..d-1: the interface living on the bond between core k-1 and core k (cap[k] slabs of N doubles).
// It holds L_k (left interface entering core k) or R_{k-1} (right interface of core k-1) — never both at once.
struct Bonds {
    std::vector<double*> b;
    double* at(int k) const { return b[k]; }
};

int bonds_alloc(bsde_ctx* c, const Train& T, int64_t N, Bonds& B) {
    size_t tot = 0;
    for (int k = 1; k < T.d; k++) tot += (size_t)T.cap[k] * N;
    BSDE_TRY(c->bond.ensure(sizeof(double) * (tot ? tot : 1)));
    B.b.assign(T.d + 1, nullptr);
    double* q = c->bond.as<double>();
    for (int k = 1; k < T.d; k++) { B.b[k] = q; q += (size_t)T.cap[k] * N; }
    return 0;
}

// R_{j-1} for j = d-1 .. stop  (bond[j] <- core_j applied to bond[j+1])
int build_right(bsde_ctx* c, const Train& T, const Inputs& in, const Bonds& B, int stop) {
    for (int j = T.d - 1; j >= stop && j >= 1; j--) {
        if (c->in_ev_count > j) BSDE_CUDA_OK(cudaStreamWaitEvent(c->stream, c->in_ev[j], 0));
        BSDE_TRY(launch_interface_right(in, j, T.core(j), T.ranks[j], T.ranks[j + 1], j == T.d - 1 ? nullptr : B.at(j + 1),
                                        B.at(j), c->stream));
        c->launches++;
    }
    return 0;
}

// L_{j+1} for j = 0 .. upto-1   (bond[j+1] <- bond[j] applied to core_j)
int build_left(bsde_ctx* c, const Train& T, const Inputs& in, const Bonds& B, int upto) {
    for (int j = 0; j < upto && j < T.d - 1; j++) {
        BSDE_TRY(launch_interface_left(in, j, T.core(j), T.ranks[j], T.ranks[j + 1], j == 0 ? nullptr : B.at(j), B.at(j + 1),
                                       c->stream));
        c->launches++;
    }
    return 0;
}

// ---- profiling helpers: no-ops unless option "profile" is set and we are not capturing a graph ----
cudaEvent_t prof_mark(bsde_ctx* c) {
    if (!c->profile || c->capturing) return nullptr;
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        c->ev_pool.push_back(e);
    }
    cudaEvent_t e = c->ev_pool[c->ev_used++];
    cudaEventRecord(e, c->stream);
    return e;
}
cudaEvent_t prof_take(bsde_ctx* c) {     // an event to be recorded by someone else
    if (!c->profile || c->capturing) return nullptr;
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        c->ev_pool.push_back(e);
    }
    return c->ev_pool[c->ev_used++];
}
void prof_span(bsde_ctx* c, int cls, cudaEvent_t a, cudaEvent_t b) {
    if (a && b) c->spans.push_back({cls, a, b});
}
void prof_collect(bsde_ctx* c) {         // after a stream synchronise
    for (const auto& s : c->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { c->prof_ms[s.cls] += ms; c->prof_n[s.cls]++; }
    }
    c->spans.clear();
    c->ev_used = 0;
}

int allreduce_moments(bsde_ctx* c, double* moments, int n) {
    if (c->world <= 1) return 0;
    const int rc = g_nccl.AllReduce(moments, moments, (size_t)n, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, c->stream);
    if (rc) return fail_msg(3, "ncclAllReduce failed: %s", g_nccl.GetErrorString(rc));
    c->launches++;
    return 0;
}

struct MicroCtx {
    const Train* T;
    const Inputs* in;
    const Bonds* B;
    const double* y;
    int solver;
    int mono;
    int* info;
    int* hints = nullptr;      // one int per core (MicroSolveArgs::hint), zeroed at the start of the fit; ALS only
};

// Programmatic dependent launch inside a captured sweep (single GPU: the sharded sweeps, whose reduction kernel waits on
// its peers, keep the plain launches they were validated with).
static int pdl_allowed(const bsde_ctx* c) { return (c->pdl && c->world == 1) ? 1 : 0; }

// assembly + (all-reduce) + solve/QR of core j; direction as MicroSolveArgs
struct SalsaReg {            // als.py:143-174, all device pointers
    const double* eta = nullptr;
    const double* gamma = nullptr;
    const double* theta = nullptr;
    int reg_left = 0, reg_right = 0;
};

int micro_step(bsde_ctx* c, const MicroCtx& m, int j, int direction, double ridge, const double* reg_diag, double* dbg,
               const SalsaReg* sr = nullptr, const FusedInterface* fi = nullptr) {
    const Train& T = *m.T;
    const int rl = T.ranks[j], rr = T.ranks[j + 1];
    const AsmPlan* P;
    BSDE_TRY(get_plan(c, rl, T.p, rr, m.mono, m.in->N, &P));
    BSDE_TRY(c->partial.ensure(sizeof(double) * (size_t)P->grid_x * P->partial_stride));
    BSDE_TRY(c->moments.ensure(sizeof(double) * (size_t)P->partial_stride * kRedGroups));
    // the partial rows are summed in `groups` row groups, added up by the solve kernel while staging.  Sharded runs
    // all-reduce every group row on its own (fused exchange: one flag per (group, block); NCCL: groups * stride doubles),
    // so the replicas still add identical numbers in identical order.
    int groups = P->grid_x < 8 * kRedGroups ? 1 : (int)c->red_groups;
    if (c->world > 1 && c->px_on && (int64_t)groups * ((P->partial_stride + 31) / 32) > kXBlocks) groups = 1;
    if (fi && fi->side && !asm_fuse_supported(*P, *m.in, *fi)) {      // this shape cannot fuse: separate interface kernel
        cudaEvent_t i0 = prof_mark(c);
        if (fi->side == 1) BSDE_TRY(launch_interface_left(*m.in, fi->j, fi->core, fi->r_in, rl, fi->in, fi->out, c->stream));
        else BSDE_TRY(launch_interface_right(*m.in, fi->j, fi->core, rr, fi->r_in, fi->in, fi->out, c->stream));
        c->launches++;
        prof_span(c, bsde_ctx::P_INTERFACE, i0, prof_mark(c));
        fi = nullptr;
    }
    cudaEvent_t e0 = prof_mark(c), e1 = prof_take(c);
    BSDE_TRY(launch_assembly(*P, *m.in, j, j == 0 ? nullptr : m.B->at(j), j == T.d - 1 ? nullptr : m.B->at(j + 1), m.y,
                             c->partial.as<double>(), c->moments.as<double>(), c->stream, e1, c->px_on ? &c->px : nullptr, fi, groups));
    c->launches += 2;
    cudaEvent_t e2 = prof_mark(c);
    prof_span(c, bsde_ctx::P_ASSEMBLY, e0, e1);
    prof_span(c, bsde_ctx::P_REDUCE, e1, e2);
    cudaEvent_t e3 = e2;
    if (c->world > 1 && !c->px_on) {          // NCCL path; with the peer exchange the all-reduce is inside the reduce kernel
        BSDE_TRY(allreduce_moments(c, c->moments.as<double>(), (int)P->partial_stride * groups));
        e3 = prof_mark(c);
        prof_span(c, bsde_ctx::P_ALLREDUCE, e2, e3);
    }
    MicroSolveArgs a{};
    a.moments = c->moments.as<double>();
    a.expand_map = P->d_expand;
    a.n_slots = (int)P->partial_stride;
    a.moment_rows = groups;
    a.rl = rl; a.p = T.p; a.rr = rr; a.mono = m.mono; a.nLL = P->nLL; a.nRR = P->nRR; a.nB = P->nB;
    a.direction = direction; a.solver = m.solver;
    a.core_j = T.core(j);
    a.core_nb = nullptr; a.r_nb = 1;
    if (direction > 0) { a.core_nb = T.core(j + 1); a.r_nb = T.ranks[j + 2]; }
    if (direction < 0) { a.core_nb = T.core(j - 1); a.r_nb = T.ranks[j - 1]; }
    a.ridge = ridge; a.reg_diag = reg_diag; a.dbg_A = dbg; a.info = m.info;
    if (c->profile && !c->capturing) {
        if (!c->clk.p) { if (c->clk.ensure(256) == 0) cudaMemset(c->clk.p, 0, 256); }
        a.clk = c->clk.as<long long>();
    }
    if (sr) { a.eta_ptr = sr->eta; a.gamma = sr->gamma; a.theta = sr->theta; a.reg_left = sr->reg_left; a.reg_right = sr->reg_right; }
    if (c->fast_solve && c->scalars.p) a.fast_flag = c->scalars.as<int>() + 24;
    if (m.hints && !sr && c->solve_hints) a.hint = m.hints + j;
    a.stop_after = (int)c->solve_stop;
    a.force_onesided = (int)c->svd_onesided;
    a.tier2 = (int)c->gate_tier2;
    if (const size_t ws = micro_solve_workspace_bytes(rl * T.p * rr, a.n_slots)) {      // cores beyond the solver's shared memory
        BSDE_TRY(c->solve_ws.ensure(ws));
        a.big_ws = c->solve_ws.as<double>();
    }
    int extra = 0;
    BSDE_TRY(launch_micro_solve(a, c->stream, &extra));
    c->launches += 1 + extra;
    prof_span(c, bsde_ctx::P_SOLVE, e3, prof_mark(c));
    return 0;
}

// one ALS sweep: als.py:61-94.
// fused = 0: every micro-step is followed by its own interface kernel (explicit basis values, option "fuse_interface" 0).
// fused = 1: the interface a micro-step leaves behind is formed inside the assembly kernel of the NEXT micro-step
//            (FusedInterface); `first` = this is the first sweep of the fit (the right interfaces were built up front,
//            so core 0 has nothing to fuse).  After the last micro-step (j = 1) no interface is needed any more.
int als_sweep(bsde_ctx* c, const MicroCtx& m, int fused = 0, int first = 1) {
    const Train& T = *m.T;
    const int d = T.d;
    auto left_of = [&](int j) {          // L_j from core j-1
        FusedInterface f;
        f.side = 1; f.j = j - 1; f.r_in = T.ranks[j - 1]; f.core = T.core(j - 1);
        f.in = j - 1 == 0 ? nullptr : m.B->at(j - 1); f.out = m.B->at(j);
        return f;
    };
    auto right_of = [&](int j) {         // R_j from core j+1
        FusedInterface f;
        f.side = 2; f.j = j + 1; f.r_in = T.ranks[j + 2]; f.core = T.core(j + 1);
        f.in = j + 1 == d - 1 ? nullptr : m.B->at(j + 2); f.out = m.B->at(j + 1);
        return f;
    };
    for (int j = 0; j < d - 1; j++) {
        if (fused) {
            FusedInterface f;
            if (j > 0) f = left_of(j);
            else if (!first) f = right_of(0);
            BSDE_TRY(micro_step(c, m, j, +1, 0.0, nullptr, nullptr, nullptr, &f));
            continue;
        }
        BSDE_TRY(micro_step(c, m, j, +1, 0.0, nullptr, nullptr));
        cudaEvent_t e0 = prof_mark(c);
        BSDE_TRY(launch_interface_left(*m.in, j, T.core(j), T.ranks[j], T.ranks[j + 1], j == 0 ? nullptr : m.B->at(j),
                                       m.B->at(j + 1), c->stream));
        c->launches++;
        prof_span(c, bsde_ctx::P_INTERFACE, e0, prof_mark(c));
    }
    for (int j = d - 1; j >= 1; j--) {
        if (fused) {
            const FusedInterface f = j == d - 1 ? left_of(j) : right_of(j);
            BSDE_TRY(micro_step(c, m, j, -1, 0.0, nullptr, nullptr, nullptr, &f));
            continue;
        }
        BSDE_TRY(micro_step(c, m, j, -1, 0.0, nullptr, nullptr));
        cudaEvent_t e0 = prof_mark(c);
        BSDE_TRY(launch_interface_right(*m.in, j, T.core(j), T.ranks[j], T.ranks[j + 1], j == d - 1 ? nullptr : m.B->at(j + 1),
                                        m.B->at(j), c->stream));
        c->launches++;
        prof_span(c, bsde_ctx::P_INTERFACE, e0, prof_mark(c));
    }
    return 0;
}

// Ranks are not synchronised when a fit starts (lazy module loads, plan allocation, a slow host): before the first fused
// exchange of a fit every rank therefore joins a one-double ncclAllReduce on the fit's stream, so that no kernel starts
// waiting for a peer that has not reached the fit yet.
int px_handshake(bsde_ctx* c) {
    if (!c->px_on || c->world <= 1) return 0;
    BSDE_TRY(c->scalars.ensure(1024));
    double* token = c->scalars.as<double>() + 120;
    BSDE_CUDA_OK(cudaMemsetAsync(token, 0, sizeof(double), c->stream));
    return allreduce_moments(c, token, 1);
}

// After a synchronised fit: did a wait for a peer time out on ANY rank?  (The rank that timed out continued on NaN sums; its
// peers may have carried on with good ones: every rank must fail, or the replicas silently diverge.)
int px_check(bsde_ctx* c) {
    if (!c->px_on) return 0;
    int err = 0;
    BSDE_CUDA_OK(cudaMemcpy(&err, c->px.error, sizeof err, cudaMemcpyDeviceToHost));
    double* token = c->scalars.as<double>() + 120;
    const double mine = err ? 1.0 : 0.0;
    BSDE_CUDA_OK(cudaMemcpyAsync(token, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    BSDE_TRY(allreduce_moments(c, token, 1));
    double any = 0.0;
    BSDE_CUDA_OK(cudaMemcpyAsync(&any, token, sizeof any, cudaMemcpyDeviceToHost, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (any != 0.0) {
        const int zero = 0;
        cudaMemcpy(c->px.error, &zero, sizeof zero, cudaMemcpyHostToDevice);
        return fail_msg(3, "fused all-reduce: a wait for a peer timed out on %s (rank %d of %d); the fit's sums were poisoned with NaN",
                        err ? "this rank" : "another rank", c->rank, c->world);
    }
    return 0;
}

int info_reset(bsde_ctx* c, int** info) {
    BSDE_TRY(c->scalars.ensure(1024));      // one size for every user of this buffer: pointers into it must stay valid
    *info = c->scalars.as<int>() + 16;
    const int init[4] = {0, 0, 0, INT_MAX};
    BSDE_CUDA_OK(cudaMemcpyAsync(*info, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int fit_als(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* inputs_dev, const double* coef_host,
            const double* y_dev, int n_iter, const int* ranks, int solver_id, double* cores_host, int* info_host) {
    BSDE_TRY(check_ranks(d, p, ranks));
    plans_trim(c);
    if (!y_dev || !cores_host) return fail_msg(2, "fit_als: null y or cores pointer");
    if (n_iter < 0) return fail_msg(2, "fit_als: n_iter = %d", n_iter);
    if (solver_id < 0 || solver_id > 3) return fail_msg(2, "Unknown method %d", solver_id);
    for (int j = 0; j < d; j++) {
        if (ranks[j] * p * ranks[j + 1] > kMaxN) return fail_msg(2, "core %d has %d unknowns (limit %d)", j, ranks[j] * p * ranks[j + 1], kMaxN);
        if (ranks[j] > 8 || ranks[j + 1] > 8) return fail_msg(2, "ALS supports ranks <= 8 (core %d is %d x %d)", j, ranks[j], ranks[j + 1]);
    }
    Inputs in;
    BSDE_TRY(make_inputs(c, in, basis_kind, d, p, N, inputs_dev, nullptr, coef_host));
    Train T;
    BSDE_TRY(train_upload(c, T, d, p, ranks, nullptr, cores_host));
    int* info;
    BSDE_TRY(info_reset(c, &info));
    BSDE_TRY(px_handshake(c));
    if (d >= 2) {
        // als.py:39 — tt.orthonormalize(mode="right", start=1)
        BSDE_TRY(launch_orthonormalize_right(T.cores, T.off_dev, T.ranks_dev, p, d, 1, d - 1, c->stream));
        c->launches++;
        Bonds B;
        BSDE_TRY(bonds_alloc(c, T, N, B));
        BSDE_TRY(build_right(c, T, in, B, 1));
        for (int j = 0; j < c->in_ev_count; j++) BSDE_CUDA_OK(cudaStreamWaitEvent(c->stream, c->in_ev[j], 0));
        c->in_ev_count = 0;
        MicroCtx m{&T, &in, &B, y_dev, solver_id, basis_kind == BASIS_POLY ? 1 : 0, info};
        BSDE_TRY(c->hints.ensure(sizeof(int) * (size_t)d));
        BSDE_CUDA_OK(cudaMemsetAsync(c->hints.p, 0, sizeof(int) * (size_t)d, c->stream));
        m.hints = c->hints.as<int>();
        int it = 0;
        const int fused = (c->fuse_interface && basis_kind != BASIS_PHI) ? 1 : 0;
        if (n_iter > 0) { BSDE_TRY(als_sweep(c, m, fused, 1)); it = 1; }     // also builds the assembly plans / workspaces
        if (it < n_iter) {
            if (c->use_graph && !c->profile && n_iter - it >= 2) {
                cudaGraph_t graph = nullptr;
                cudaGraphExec_t exec = nullptr;
                const int64_t before = c->launches;
                BSDE_CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                c->capturing = true;
                g_pdl_launch = pdl_allowed(c);
                const int rc = als_sweep(c, m, fused, 0);
                g_pdl_launch = 0;
                c->capturing = false;
                cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (e != cudaSuccess) return fail_cuda(e, "cudaStreamEndCapture", __FILE__, __LINE__);
                {   // how many of the kernel -> kernel edges of the sweep are programmatic (stats "graph_edges", "graph_pdl_edges")
                    size_t ne = 0;
                    c->graph_edges = c->graph_pdl_edges = 0;
                    if (cudaGraphGetEdges_v2(graph, nullptr, nullptr, nullptr, &ne) == cudaSuccess && ne > 0) {
                        std::vector<cudaGraphNode_t> from(ne), to(ne);
                        std::vector<cudaGraphEdgeData> ed(ne);
                        if (cudaGraphGetEdges_v2(graph, from.data(), to.data(), ed.data(), &ne) == cudaSuccess) {
                            c->graph_edges = (int64_t)ne;
                            for (size_t i = 0; i < ne; i++) c->graph_pdl_edges += ed[i].type == cudaGraphDependencyTypeProgrammatic;
                        }
                    }
                    (void)cudaGetLastError();
                }
                const int64_t per_sweep = c->launches - before;
                c->launches = before;
                e = cudaGraphInstantiate(&exec, graph, 0);
                if (e != cudaSuccess) { cudaGraphDestroy(graph); return fail_cuda(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
                for (; it < n_iter; it++) {
                    e = cudaGraphLaunch(exec, c->stream);
                    if (e != cudaSuccess) break;
                    c->launches += per_sweep;
                }
                cudaGraphExecDestroy(exec);
                cudaGraphDestroy(graph);
                if (e != cudaSuccess) return fail_cuda(e, "cudaGraphLaunch", __FILE__, __LINE__);
            } else {
                for (; it < n_iter; it++) BSDE_TRY(als_sweep(c, m, fused, 0));
            }
        }
    }
    BSDE_TRY(train_download(c, T, cores_host));
    BSDE_TRY(px_check(c));
    prof_collect(c);
    if (info_host) {
        BSDE_CUDA_OK(cudaMemcpy(info_host, info, sizeof(int) * 4, cudaMemcpyDeviceToHost));
        if (info_host[3] == INT_MAX) info_host[3] = 0;
    }
    return 0;
}


// SALSA (als.py:99-323).  The host sequences the launches and mirrors the integer bookkeeping (sweep counters, bond
// ranks); every number — SVDs, rank-adaptation test S[-1] > min_sv, regulariser, eta / omega / min_sv — is computed on
// the device.  The only device->host traffic during a fit is the bond rank after a possible adaptation (one int).
int fit_salsa(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* inputs_dev, const double* coef_host,
              const double* y_dev, int n_iter, int* ranks, const int* init_ranks, double omega, double freq_omega, double min_sv,
              int max_rank, int solver_id, int do_reg, double* cores_host, int64_t cores_capacity, double* state_out_host) {
    BSDE_TRY(check_ranks(d, p, ranks));
    plans_trim(c);
    if (!init_ranks) return fail_msg(2, "fit_salsa: init_ranks is null");
    if (!y_dev || !cores_host) return fail_msg(2, "fit_salsa: null y or cores pointer");
    if (solver_id < 0 || solver_id > 3) return fail_msg(2, "Unknown method %d", solver_id);
    if (n_iter < 0) return fail_msg(2, "fit_salsa: n_iter = %d", n_iter);
    const int kRankLimit = 8;
    // als.py:122-123 — caps from the ranks the train was CONSTRUCTED with; the bond left of core j reads max_ranks[j-1]
    std::vector<int> max_ranks(d > 1 ? d - 1 : 1, 1);
    for (int i = 0; i < d - 1; i++) max_ranks[i] = std::min(init_ranks[i] * p, max_rank);
    if (d > 1) max_ranks[d - 2] = std::min(init_ranks[d] * p, max_rank);
    std::vector<int> cap(d + 1, 1);
    for (int k = 1; k < d; k++) {
        if (ranks[k] > kRankLimit) return fail_msg(2, "SALSA supports ranks <= %d (bond %d has %d)", kRankLimit, k, ranks[k]);
        cap[k] = std::max(ranks[k], std::min(max_ranks[k - 1], kRankLimit));
    }
    int64_t need = 0;
    for (int j = 0; j < d; j++) need += (int64_t)cap[j] * p * cap[j + 1];
    if (cores_capacity < need) return fail_msg(2, "fit_salsa: cores buffer holds %lld doubles, rank growth may need %lld", (long long)cores_capacity, (long long)need);

    Inputs in;
    BSDE_TRY(make_inputs(c, in, basis_kind, d, p, N, inputs_dev, nullptr, coef_host));
    Train T;
    BSDE_TRY(train_upload(c, T, d, p, ranks, cap.data(), cores_host));
    int* info;
    BSDE_TRY(info_reset(c, &info));
    BSDE_TRY(px_handshake(c));
    // device scalars: state = {eta, omega, min_sv}, residual sum, gamma[8], theta[8]; ints: adapted flag, new rank
    BSDE_TRY(c->scalars.ensure(1024));
    double* state = c->scalars.as<double>() + 32;
    double* res_sum = state + 4;
    double* gamma = state + 8;
    double* theta = state + 24;
    int* flags = c->scalars.as<int>() + 8;           // [0] adapted, [1] rank
    const double init_state[3] = {1e-6, omega, min_sv};     // als.py:121
    BSDE_CUDA_OK(cudaMemcpyAsync(state, init_state, sizeof init_state, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    double n_total = (double)N * c->world;           // shards are equal-sized up to one sample; exact count all-reduced below
    if (c->world > 1) {
        double cnt = (double)N;
        BSDE_CUDA_OK(cudaMemcpyAsync(res_sum, &cnt, sizeof cnt, cudaMemcpyHostToDevice, c->stream));
        BSDE_TRY(allreduce_moments(c, res_sum, 1));
        BSDE_CUDA_OK(cudaMemcpyAsync(&cnt, res_sum, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
        BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
        n_total = cnt;
    }

    Bonds B;
    BSDE_TRY(bonds_alloc(c, T, N, B));
    const int rgrid = residual_grid(N);
    BSDE_TRY(c->tmp.ensure(sizeof(double) * (rgrid + 8)));
    MicroCtx m{&T, &in, &B, y_dev, solver_id, basis_kind == BASIS_POLY ? 1 : 0, info};

    bool expand_rank = false, has_adapted = false;
    const int warmup = 2;
    int last_adapt = 0, rank_limited = 0;
    // one sweep, als.py:222-298 (everything a sweep launches; the host decisions about expand_rank stay in the loop below)
    auto sweep = [&](int it) -> int {
        if (d >= 2) {
            // (a captured sweep has fixed ranks: the device copy is refreshed once, before the capture)
            if (!c->capturing)
                BSDE_CUDA_OK(cudaMemcpyAsync(T.ranks_dev, T.ranks.data(), sizeof(int) * (d + 1), cudaMemcpyHostToDevice, c->stream));
            BSDE_TRY(launch_orthonormalize_right(T.cores, T.off_dev, T.ranks_dev, p, d, 1, d - 1, c->stream));   // als.py:234
            c->launches++;
            BSDE_TRY(build_right(c, T, in, B, 2));          // bond[k] = R_{k-1} for k >= 2 (R_0 is rebuilt at j = 0)
        }
        for (int j = 0; j < d; j++) {
            SalsaReg sr;
            FusedInterface fi;
            sr.eta = state;
            if (j != 0) {
                const int k = T.ranks[j];
                const bool may_grow = expand_rank && k < max_ranks[j - 1];
                if (may_grow && k >= kRankLimit) rank_limited = 1;
                const int expand = (may_grow && k < kRankLimit) ? 1 : 0;
                if (expand) BSDE_CUDA_OK(cudaMemsetAsync(flags, 0, 2 * sizeof(int), c->stream));
                BSDE_TRY(launch_salsa_left(T.core(j - 1), T.core(j), T.ranks[j - 1], p, k, T.ranks[j + 1], state + 2, expand, do_reg,
                                           gamma, flags + 1, flags, c->stream));
                c->launches++;
                if (expand) {
                    int h[2];
                    BSDE_CUDA_OK(cudaMemcpyAsync(h, flags, sizeof h, cudaMemcpyDeviceToHost, c->stream));
                    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
                    if (h[0]) { has_adapted = true; T.ranks[j] = h[1]; }
                }
                if (do_reg) { sr.gamma = gamma; sr.reg_left = 1; }
                // L_j from the final core_{j-1}: formed inside this micro-step's assembly kernel (FusedInterface) where the
                // shape allows it, by its own kernel otherwise (micro_step decides)
                fi.side = 1; fi.j = j - 1; fi.r_in = T.ranks[j - 1]; fi.core = T.core(j - 1);
                fi.in = j - 1 == 0 ? nullptr : B.at(j - 1); fi.out = B.at(j);
                if (!(c->fuse_interface && basis_kind != BASIS_PHI)) {
                    BSDE_TRY(launch_interface_left(in, j - 1, T.core(j - 1), T.ranks[j - 1], T.ranks[j], fi.in, fi.out, c->stream));
                    c->launches++;
                    fi.side = 0;
                }
            }
            if (j != d - 1) {
                BSDE_TRY(launch_salsa_right(T.core(j), T.core(j + 1), T.ranks[j], p, T.ranks[j + 1], T.ranks[j + 2], state + 2, do_reg,
                                            theta, c->stream));
                c->launches++;
                if (do_reg) { sr.theta = theta; sr.reg_right = 1; }
                // R_j from the moved core_{j+1}
                BSDE_TRY(launch_interface_right(in, j + 1, T.core(j + 1), T.ranks[j + 1], T.ranks[j + 2],
                                                j + 1 == d - 1 ? nullptr : B.at(j + 2), B.at(j + 1), c->stream));
                c->launches++;
            }
            const double* Lj = j == 0 ? nullptr : B.at(j);
            const double* Rj = j == d - 1 ? nullptr : B.at(j + 1);
            if (j == 0 && it == 0) {                                                      // als.py:168-172
                BSDE_TRY(launch_residual(in, j, T.core(j), T.ranks[j], T.ranks[j + 1], Lj, Rj, y_dev, nullptr, c->tmp.as<double>(),
                                         res_sum, c->stream));
                c->launches += 2;
                BSDE_TRY(allreduce_moments(c, res_sum, 1));
                BSDE_TRY(launch_salsa_scalars(state, res_sum, n_total, freq_omega, 0, c->stream));
                c->launches++;
            }
            BSDE_TRY(micro_step(c, m, j, 0, 0.0, nullptr, nullptr, &sr, fi.side ? &fi : nullptr));
        }
        // als.py:288-298 — residual of the last micro-step's design matrix with the new last core
        {
            const int j = d - 1;
            BSDE_TRY(launch_residual(in, j, T.core(j), T.ranks[j], T.ranks[j + 1], j == 0 ? nullptr : B.at(j), nullptr, y_dev, nullptr,
                                     c->tmp.as<double>(), res_sum, c->stream));
            c->launches += 2;
            BSDE_TRY(allreduce_moments(c, res_sum, 1));
            BSDE_TRY(launch_salsa_scalars(state, res_sum, n_total, freq_omega, 1, c->stream));
            c->launches++;
        }
            return 0;
    };
    // A sweep in which no bond may grow launches a fixed sequence of kernels (no rank read-back): it is captured once
    // into a CUDA graph and replayed for as long as that stays true and the ranks do not change — after the adaptations
    // have saturated that is every remaining sweep of the fit (measured at C4, round 2: 285 us per micro-step of eager
    // launches against ~150 us of kernel time).
    cudaGraph_t sgraph = nullptr;
    cudaGraphExec_t sexec = nullptr;
    int64_t sgraph_launches = 0;
    auto drop_graph = [&]() {
        if (sexec) cudaGraphExecDestroy(sexec);
        if (sgraph) cudaGraphDestroy(sgraph);
        sexec = nullptr; sgraph = nullptr;
    };
    int rc_fit = 0;
    for (int it = 0; it < n_iter && !rc_fit; it++) {
        if (it >= last_adapt + warmup) expand_rank = true;                               // als.py:224
        if (has_adapted) { expand_rank = false; has_adapted = false; last_adapt = it; }   // als.py:225
        bool may_grow_any = false;
        if (expand_rank)
            for (int j = 1; j < d; j++) {
                if (T.ranks[j] < max_ranks[j - 1] && T.ranks[j] >= kRankLimit) rank_limited = 1;
                if (T.ranks[j] < max_ranks[j - 1] && T.ranks[j] < kRankLimit) may_grow_any = true;
            }
        const bool fixed = it > 0 && !may_grow_any && d >= 2 && c->use_graph && !c->profile;
        if (!fixed) {
            drop_graph();                                   // the ranks may change in this sweep
            rc_fit = sweep(it);
            continue;
        }
        if (!sexec) {
            // nothing may allocate inside the capture: plans and workspaces of every core shape of the current ranks first
            // (a bond that grew in the last eager sweep changed the shape of the core to its left after that core's turn)
            BSDE_CUDA_OK(cudaMemcpyAsync(T.ranks_dev, T.ranks.data(), sizeof(int) * (d + 1), cudaMemcpyHostToDevice, c->stream));
            BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
            for (int j = 0; j < d; j++) {
                const AsmPlan* P;
                BSDE_TRY(get_plan(c, T.ranks[j], p, T.ranks[j + 1], m.mono, N, &P));
                BSDE_TRY(c->partial.ensure(sizeof(double) * (size_t)P->grid_x * P->partial_stride));
                BSDE_TRY(c->moments.ensure(sizeof(double) * (size_t)P->partial_stride * kRedGroups));
                BSDE_TRY(c->solve_ws.ensure(micro_solve_workspace_bytes(T.ranks[j] * p * T.ranks[j + 1], (int)P->partial_stride)));
            }
            const int64_t before = c->launches;
            BSDE_CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            c->capturing = true;
            g_pdl_launch = pdl_allowed(c);
            const int rc = sweep(it);
            g_pdl_launch = 0;
            c->capturing = false;
            cudaError_t e = cudaStreamEndCapture(c->stream, &sgraph);
            if (rc) { drop_graph(); return rc; }
            if (e != cudaSuccess) { drop_graph(); return fail_cuda(e, "cudaStreamEndCapture", __FILE__, __LINE__); }
            sgraph_launches = c->launches - before;
            c->launches = before;
            e = cudaGraphInstantiate(&sexec, sgraph, 0);
            if (e != cudaSuccess) { drop_graph(); return fail_cuda(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
        }
        const cudaError_t e = cudaGraphLaunch(sexec, c->stream);
        if (e != cudaSuccess) { drop_graph(); return fail_cuda(e, "cudaGraphLaunch", __FILE__, __LINE__); }
        c->launches += sgraph_launches;
    }
    if (sexec) {                                            // the graph reads host and device buffers of this fit
        const cudaError_t e = cudaStreamSynchronize(c->stream);
        drop_graph();
        if (e != cudaSuccess) return fail_cuda(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    if (rc_fit) return rc_fit;
    BSDE_TRY(train_download(c, T, cores_host));
    BSDE_TRY(px_check(c));
    for (int k = 0; k <= d; k++) ranks[k] = T.ranks[k];
    if (state_out_host) {
        BSDE_CUDA_OK(cudaMemcpy(state_out_host, state, sizeof(double) * 3, cudaMemcpyDeviceToHost));
        state_out_host[3] = rank_limited;
        int hinfo[4];
        BSDE_CUDA_OK(cudaMemcpy(hinfo, info, sizeof hinfo, cudaMemcpyDeviceToHost));
        state_out_host[4] = hinfo[0]; state_out_host[5] = hinfo[1]; state_out_host[6] = hinfo[2];
    }
    return 0;
}

}  // namespace

// ================================================ C ABI ================================================
extern "C" {

const char* bsde_last_error(void) { return g_err.c_str(); }
int bsde_version(void) { return 100; }

int bsde_ctx_create(int device, bsde_ctx** out) {
    if (!out) return fail_msg(2, "bsde_ctx_create: out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail_msg(4, "no CUDA device available (%s) — this backend has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail_msg(2, "device %d out of range (0..%d)", device, count - 1);
    BSDE_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    BSDE_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail_msg(4, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    bsde_ctx* c = new bsde_ctx();
    if (const char* e = getenv("BSDE_B200_PDL")) c->pdl = atoi(e);      // A/B switch of option "pdl" for whole test runs
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    BSDE_CUDA_OK(cudaStreamCreateWithFlags(&c->own, cudaStreamNonBlocking));
    c->stream = c->own;
    *out = c;
    return 0;
}

int bsde_ctx_destroy(bsde_ctx* c) {
    if (!c) return 0;
    Guard g(c);
    cudaStreamSynchronize(c->stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (auto& kv : c->plans) asm_plan_free(kv.second);
    for (DevBuf* b : {&c->bond, &c->partial, &c->moments, &c->cores, &c->meta, &c->coef, &c->tmp, &c->in_copy, &c->y_copy, &c->scalars, &c->dbg, &c->hints, &c->solve_ws})
        b->release();
    for (void* q : c->px_opened) if (q) cudaIpcCloseMemHandle(q);
    if (c->px_data_local) cudaFree(c->px_data_local);
    if (c->px_flags_local) cudaFree(c->px_flags_local);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->in_ev) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->own);
    delete c;
    return 0;
}

int bsde_ctx_set_stream(bsde_ctx* c, void* cuda_stream) {
    BSDE_CHECK_CTX(c);
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own;
    return 0;
}

int bsde_ctx_sync(bsde_ctx* c) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int bsde_ctx_launch_count(bsde_ctx* c, int64_t* out) {
    BSDE_CHECK_CTX(c);
    if (out) *out = c->launches;
    return 0;
}

int bsde_ctx_set_option(bsde_ctx* c, const char* name, int64_t value) {
    BSDE_CHECK_CTX(c);
    if (name && !strcmp(name, "use_graph")) { c->use_graph = value; return 0; }
    if (name && !strcmp(name, "px_timeout_ms")) {          // how long the fused all-reduce waits for a peer before it gives up
        if (value < 1) return fail_msg(2, "px_timeout_ms must be positive");
        c->px.timeout_cycles = (long long)value * 2000000ll;
        return 0;
    }
    if (name && !strcmp(name, "fuse_interface")) { c->fuse_interface = value; return 0; }
    if (name && !strcmp(name, "fast_solve")) { c->fast_solve = value; return 0; }
    if (name && !strcmp(name, "pdl")) { c->pdl = value; return 0; }
    if (name && !strcmp(name, "solve_stop")) { c->solve_stop = value; return 0; }
    if (name && !strcmp(name, "svd_onesided")) { c->svd_onesided = value; return 0; }
    if (name && !strcmp(name, "gate_tier2")) { c->gate_tier2 = value; return 0; }
    if (name && !strcmp(name, "solve_hints")) { c->solve_hints = value; return 0; }
    if (name && !strcmp(name, "r4_full")) { g_asm_r4_full = value ? 1 : 0; return 0; }
    if (name && !strcmp(name, "red_groups")) { if (value < 1 || value > kRedGroups) return fail_msg(2, "red_groups outside 1..%d", kRedGroups); c->red_groups = value; return 0; }
    if (name && !strcmp(name, "asm_generic")) {
        c->asm_generic = value;
        for (auto& kv : c->plans) asm_plan_free(kv.second);
        c->plans.clear();
        return 0;
    }
    if (name && !strcmp(name, "profile")) {      // per-kernel-class event timing of eager launches; resets the counters
        c->profile = value;
        for (int k = 0; k < bsde_ctx::P_CLASSES; k++) { c->prof_ms[k] = 0.0; c->prof_n[k] = 0; }
        if (c->clk.p) cudaMemset(c->clk.p, 0, 256);
        return 0;
    }
    return fail_msg(2, "unknown option '%s'", name ? name : "(null)");
}

int bsde_ctx_get_stat(bsde_ctx* c, const char* name, double* out) {
    BSDE_CHECK_CTX(c);
    if (!name || !out) return fail_msg(2, "bsde_ctx_get_stat: null argument");
    static const char* cls[bsde_ctx::P_CLASSES] = {"assembly", "reduce", "allreduce", "solve", "interface"};
    for (int k = 0; k < bsde_ctx::P_CLASSES; k++) {
        const size_t n = strlen(cls[k]);
        if (!strncmp(name, cls[k], n)) {
            if (!strcmp(name + n, "_ms")) { *out = c->prof_ms[k]; return 0; }
            if (!strcmp(name + n, "_n")) { *out = (double)c->prof_n[k]; return 0; }
        }
    }
    if (!strncmp(name, "solve_phase", 11) && name[11] >= '0' && name[11] <= '9') {
        long long v[32] = {};
        const int k = atoi(name + 11);
        if (k < 0 || k >= 32) return fail_msg(2, "unknown stat '%s'", name);
        if (c->clk.p) BSDE_CUDA_OK(cudaMemcpy(v, c->clk.p, sizeof v, cudaMemcpyDeviceToHost));
        *out = (double)v[k];
        return 0;
    }
    if (!strcmp(name, "sm_count")) { *out = c->sm_count; return 0; }
    if (!strcmp(name, "graph_edges")) { *out = (double)c->graph_edges; return 0; }
    if (!strcmp(name, "graph_pdl_edges")) { *out = (double)c->graph_pdl_edges; return 0; }
    if (!strcmp(name, "workspace_bytes")) {
        double t = 0;
        for (const DevBuf* b : {&c->bond, &c->partial, &c->moments, &c->cores, &c->meta, &c->coef, &c->tmp, &c->in_copy, &c->y_copy, &c->scalars, &c->dbg, &c->solve_ws})
            t += (double)b->cap;
        *out = t;
        return 0;
    }
    return fail_msg(2, "unknown stat '%s'", name);
}

int bsde_comm_unique_id(void* out128) {
    if (!out128) return fail_msg(2, "bsde_comm_unique_id: out is null");
    BSDE_TRY(nccl_load());
    ncclUniqueId_ id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc) return fail_msg(3, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString(rc));
    memcpy(out128, &id, 128);
    return 0;
}

int bsde_comm_init(bsde_ctx* c, const void* id128, int rank, int world) {
    BSDE_CHECK_CTX(c);
    if (world < 1 || rank < 0 || rank >= world) return fail_msg(2, "bsde_comm_init: rank %d of %d", rank, world);
    if (world == 1) { c->rank = 0; c->world = 1; return 0; }
    if (!id128) return fail_msg(2, "bsde_comm_init: id is null");
    BSDE_TRY(nccl_load());
    Guard g(c);
    ncclUniqueId_ id;
    memcpy(&id, id128, 128);
    const int rc = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (rc) return fail_msg(3, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(rc));
    c->rank = rank; c->world = world;
    return 0;
}

int bsde_comm_p2p_export(bsde_ctx* c, void* out128) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    if (!out128) return fail_msg(2, "bsde_comm_p2p_export: out is null");
    if (c->world < 2 || c->world > kXMaxWorld) return fail_msg(2, "bsde_comm_p2p_export: call bsde_comm_init first (world 2..%d)", kXMaxWorld);
    if (!c->px_data_local) {
        const size_t nd = (size_t)2 * kXMaxWorld * kXSlots * sizeof(double);
        const size_t nf = (size_t)2 * kXMaxWorld * kXBlocks * sizeof(unsigned long long) + kXBlocks * sizeof(unsigned long long) + 64;
        BSDE_CUDA_OK(cudaMalloc(&c->px_data_local, nd));
        BSDE_CUDA_OK(cudaMalloc(&c->px_flags_local, nf));
        BSDE_CUDA_OK(cudaMemset(c->px_data_local, 0xFF, nd));        // every slot empty
        BSDE_CUDA_OK(cudaMemset(c->px_flags_local, 0, nf));
    }
    cudaIpcMemHandle_t h[2];
    BSDE_CUDA_OK(cudaIpcGetMemHandle(&h[0], c->px_data_local));
    BSDE_CUDA_OK(cudaIpcGetMemHandle(&h[1], c->px_flags_local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out128, h, 128);
    return 0;
}

int bsde_comm_p2p_open(bsde_ctx* c, const void* all_handles, int enable) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    if (!enable) { c->px_on = false; return 0; }
    if (!all_handles || !c->px_data_local) return fail_msg(2, "bsde_comm_p2p_open: export first");
    const int W = c->world;
    c->px.world = W; c->px.rank = c->rank;
    for (int p = 0; p < W; p++) {
        if (p == c->rank) { c->px.data[p] = c->px_data_local; c->px.flags[p] = c->px_flags_local; continue; }
        cudaIpcMemHandle_t h[2];
        memcpy(h, (const char*)all_handles + (size_t)p * 128, 128);
        void *pd = nullptr, *pf = nullptr;
        BSDE_CUDA_OK(cudaIpcOpenMemHandle(&pd, h[0], cudaIpcMemLazyEnablePeerAccess));
        BSDE_CUDA_OK(cudaIpcOpenMemHandle(&pf, h[1], cudaIpcMemLazyEnablePeerAccess));
        c->px_opened[2 * p] = pd; c->px_opened[2 * p + 1] = pf;
        c->px.data[p] = (double*)pd; c->px.flags[p] = (unsigned long long*)pf;
    }
    c->px.step = c->px_flags_local + (size_t)2 * kXMaxWorld * kXBlocks;
    c->px.error = (int*)(c->px.step + kXBlocks);
    c->px_on = true;
    return 0;
}

int bsde_malloc(bsde_ctx* c, int64_t bytes, void** out_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    if (!out_dev || bytes < 0) return fail_msg(2, "bsde_malloc: bad arguments");
    BSDE_CUDA_OK(cudaMalloc(out_dev, (size_t)(bytes ? bytes : 1)));
    return 0;
}

int bsde_free(bsde_ctx* c, void* ptr_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    BSDE_CUDA_OK(cudaFree(ptr_dev));
    return 0;
}

int bsde_memcpy_h2d(bsde_ctx* c, void* dst_dev, const void* src_host, int64_t bytes) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_CUDA_OK(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int bsde_memcpy_d2h(bsde_ctx* c, void* dst_host, const void* src_dev, int64_t bytes) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_CUDA_OK(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int bsde_transpose(bsde_ctx* c, const double* src_dev, int64_t rows, int64_t cols, double* dst_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_TRY(launch_transpose(src_dev, rows, cols, dst_dev, c->stream));
    c->launches++;
    return 0;
}

int bsde_fit_als(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* inputs_dev, const double* coef_host,
                 const double* y_dev, int n_iter, const int* ranks, int solver_id, double* cores_host, int* info_host) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    return fit_als(c, basis_kind, d, p, N, inputs_dev, coef_host, y_dev, n_iter, ranks, solver_id, cores_host, info_host);
}

// Host-buffer fits: asset j's block (N values of x_j, or its p basis rows [p][N]) is read from asset_ptrs[j].  The blocks
// arrive asset by asset on a copy stream in the order the right interfaces are built (d-1 .. 1, then asset 0 and y), so
// that build_right overlaps the tail of the transfer.
static int fit_als_host_blocks(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* const* asset_ptrs,
                               const double* coef_host, const double* y_host, int n_iter, const int* ranks, int solver_id,
                               double* cores_host, int* info_host) {
    if (!asset_ptrs || !y_host) return fail_msg(2, "fit_als_host: null input pointer");
    if (N < 1 || d < 1 || p < 1) return fail_msg(2, "fit_als_host: bad sizes");
    if (basis_kind < 0 || basis_kind > 2) return fail_msg(2, "unknown basis_kind %d", basis_kind);
    for (int j = 0; j < d; j++)
        if (!asset_ptrs[j]) return fail_msg(2, "fit_als_host: asset %d has a null block", j);
    const size_t chunk = sizeof(double) * (size_t)(basis_kind == BSDE_BASIS_PHI ? p : 1) * (size_t)N;   // one asset
    BSDE_TRY(c->in_copy.ensure(chunk * (size_t)d));
    BSDE_TRY(c->y_copy.ensure(sizeof(double) * (size_t)N));
    if (d >= 2 && c->world == 1) {
        if (!c->copy_stream) BSDE_CUDA_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        while ((int)c->in_ev.size() < d + 2) {
            cudaEvent_t e;
            BSDE_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->in_ev.push_back(e);
        }
        // the copy stream starts after whatever the context's stream still does with the staging buffers
        BSDE_CUDA_OK(cudaEventRecord(c->in_ev[d + 1], c->stream));
        BSDE_CUDA_OK(cudaStreamWaitEvent(c->copy_stream, c->in_ev[d + 1], 0));
        for (int j = d - 1; j >= 0; j--) {
            BSDE_CUDA_OK(cudaMemcpyAsync((char*)c->in_copy.p + (size_t)j * chunk, asset_ptrs[j], chunk, cudaMemcpyHostToDevice, c->copy_stream));
            BSDE_CUDA_OK(cudaEventRecord(c->in_ev[j], c->copy_stream));
        }
        BSDE_CUDA_OK(cudaMemcpyAsync(c->y_copy.p, y_host, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, c->copy_stream));
        BSDE_CUDA_OK(cudaEventRecord(c->in_ev[d], c->copy_stream));
        c->in_ev_count = d + 1;
    } else {
        for (int j = 0; j < d; j++)
            BSDE_CUDA_OK(cudaMemcpyAsync((char*)c->in_copy.p + (size_t)j * chunk, asset_ptrs[j], chunk, cudaMemcpyHostToDevice, c->stream));
        BSDE_CUDA_OK(cudaMemcpyAsync(c->y_copy.p, y_host, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, c->stream));
    }
    const int rc = fit_als(c, basis_kind, d, p, N, c->in_copy.as<double>(), coef_host, c->y_copy.as<double>(), n_iter, ranks, solver_id,
                           cores_host, info_host);
    if (c->in_ev_count) {            // the fit returned before consuming the events (error path): do not leave copies in flight
        cudaStreamSynchronize(c->copy_stream);
        c->in_ev_count = 0;
    }
    return rc;
}

int bsde_fit_als_host(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* inputs_host,
                      const double* coef_host, const double* y_host, int n_iter, const int* ranks, int solver_id,
                      double* cores_host, int* info_host) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    if (!inputs_host) return fail_msg(2, "fit_als_host: null input pointer");
    if (N < 1 || d < 1 || p < 1) return fail_msg(2, "fit_als_host: bad sizes");
    const size_t chunk = (size_t)(basis_kind == BSDE_BASIS_PHI ? p : 1) * (size_t)N;
    std::vector<const double*> ptrs((size_t)d);
    for (int j = 0; j < d; j++) ptrs[j] = inputs_host + (size_t)j * chunk;
    return fit_als_host_blocks(c, basis_kind, d, p, N, ptrs.data(), coef_host, y_host, n_iter, ranks, solver_id, cores_host, info_host);
}

int bsde_fit_als_host_assets(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* const* asset_ptrs_host,
                             const double* coef_host, const double* y_host, int n_iter, const int* ranks, int solver_id,
                             double* cores_host, int* info_host) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    return fit_als_host_blocks(c, basis_kind, d, p, N, asset_ptrs_host, coef_host, y_host, n_iter, ranks, solver_id, cores_host, info_host);
}

int bsde_fit_salsa(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* inputs_dev, const double* coef_host,
                   const double* y_dev, int n_iter, int* ranks, const int* init_ranks, double omega, double freq_omega,
                   double min_sv, int max_rank, int solver_id, int do_reg, double* cores_host, int64_t cores_capacity,
                   double* state_out_host) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    return fit_salsa(c, basis_kind, d, p, N, inputs_dev, coef_host, y_dev, n_iter, ranks, init_ranks, omega, freq_omega, min_sv,
                     max_rank, solver_id, do_reg, cores_host, cores_capacity, state_out_host);
}

int bsde_tt_eval(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* inputs_dev, const double* coef_host,
                 const int* ranks, const double* cores_host, double* y_out_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_TRY(check_ranks(d, p, ranks));
    if (!cores_host || !y_out_dev) return fail_msg(2, "tt_eval: null pointer");
    Inputs in;
    BSDE_TRY(make_inputs(c, in, basis_kind, d, p, N, inputs_dev, nullptr, coef_host));
    Train T;
    BSDE_TRY(train_upload(c, T, d, p, ranks, nullptr, cores_host));
    BSDE_TRY(launch_tt_eval(in, T.cores, T.off_dev, T.ranks_dev, T.max_rank(), y_out_dev, c->stream));
    c->launches++;
    return 0;
}

int bsde_tt_grad(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* inputs_dev, const double* dphi_dev,
                 const double* coef_host, const int* ranks, const double* cores_host, double* z_out_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_TRY(check_ranks(d, p, ranks));
    if (!cores_host || !z_out_dev) return fail_msg(2, "tt_grad: null pointer");
    if (basis_kind == BSDE_BASIS_PHI && !dphi_dev) return fail_msg(2, "tt_grad: explicit basis values need dphi_dev");
    Inputs in;
    BSDE_TRY(make_inputs(c, in, basis_kind, d, p, N, inputs_dev, dphi_dev, coef_host));
    Train T;
    BSDE_TRY(train_upload(c, T, d, p, ranks, nullptr, cores_host));
    Bonds B;
    BSDE_TRY(bonds_alloc(c, T, N, B));
    BSDE_TRY(build_right(c, T, in, B, 1));           // bond[k] = R_{k-1}
    const int rmax = T.max_rank();
    BSDE_TRY(c->tmp.ensure(sizeof(double) * 2 * (size_t)rmax * N));
    double* Lbuf[2] = {c->tmp.as<double>(), c->tmp.as<double>() + (size_t)rmax * N};
    for (int i = 0; i < d; i++) {
        const double* Lin = i == 0 ? nullptr : Lbuf[i & 1];
        double* Lout = i == d - 1 ? nullptr : Lbuf[(i + 1) & 1];
        BSDE_TRY(launch_grad_step(in, i, T.core(i), T.ranks[i], T.ranks[i + 1], Lin, i == d - 1 ? nullptr : B.at(i + 1), Lout,
                                  z_out_dev + (size_t)i * N, c->stream));
        c->launches++;
    }
    return 0;
}

int bsde_tt_hessian(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* inputs_dev, const double* dphi_dev,
                    const double* d2phi_dev, const double* coef_host, const int* ranks, const double* cores_host, double* h_out_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_TRY(check_ranks(d, p, ranks));
    if (!cores_host || !h_out_dev) return fail_msg(2, "tt_hessian: null pointer");
    if (basis_kind == BSDE_BASIS_PHI && (!dphi_dev || !d2phi_dev)) return fail_msg(2, "tt_hessian: explicit basis values need dphi_dev and d2phi_dev");
    Inputs in;
    BSDE_TRY(make_inputs(c, in, basis_kind, d, p, N, inputs_dev, dphi_dev, coef_host));
    in.d2data = d2phi_dev;
    Train T;
    BSDE_TRY(train_upload(c, T, d, p, ranks, nullptr, cores_host));
    if (T.max_rank() > 8) return fail_msg(2, "tt_hessian: ranks <= 8 supported");
    Bonds B;
    BSDE_TRY(bonds_alloc(c, T, N, B));
    BSDE_TRY(build_right(c, T, in, B, 1));           // bond[k] = R_{k-1}
    const int rmax = T.max_rank();
    BSDE_TRY(c->tmp.ensure(sizeof(double) * 2 * (size_t)rmax * N + sizeof(double*) * (size_t)d + 64));
    double* Lbuf[2] = {c->tmp.as<double>(), c->tmp.as<double>() + (size_t)rmax * N};
    const double** rb_dev = (const double**)(c->tmp.as<double>() + 2 * (size_t)rmax * N);
    std::vector<const double*> rb((size_t)d);
    for (int j = 0; j < d; j++) rb[j] = j == d - 1 ? nullptr : B.at(j + 1);
    BSDE_CUDA_OK(cudaMemcpyAsync(rb_dev, rb.data(), sizeof(double*) * (size_t)d, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < d; i++) {
        const double* Lin = i == 0 ? nullptr : Lbuf[i & 1];
        BSDE_TRY(launch_hessian_row(in, i, T.cores, T.off_dev, T.ranks_dev, rb_dev, Lin, rmax, h_out_dev, c->stream));
        c->launches++;
        if (i < d - 1) {      // L_{i+1} = L_i . (C_i . phi_i)
            BSDE_TRY(launch_interface_left(in, i, T.core(i), T.ranks[i], T.ranks[i + 1], Lin, Lbuf[(i + 1) & 1], c->stream));
            c->launches++;
        }
    }
    return 0;
}

static int model_params(int model_id, const double* params, double& r, double& sigma0) {
    if (!params) return fail_msg(2, "model params is null");
    if (model_id == BSDE_MODEL_BLACKSCHOLES) { r = params[0]; sigma0 = params[1]; return 0; }
    if (model_id == BSDE_MODEL_HJB) { r = 0.0; sigma0 = params[0]; return 0; }
    return fail_msg(2, "unknown model_id %d", model_id);
}

int bsde_paths_simulate(bsde_ctx* c, int model_id, const double* params, int d, int64_t N, int steps, double T,
                        const double* x0_host, uint64_t seed, int64_t sample_offset, int replay, double* xi_dev,
                        double* x_out_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    double r, sigma0;
    BSDE_TRY(model_params(model_id, params, r, sigma0));
    if (d < 1 || N < 1 || steps < 1 || !x_out_dev) return fail_msg(2, "paths_simulate: bad arguments");
    if (replay && !xi_dev) return fail_msg(2, "paths_simulate: replay needs xi_dev");
    const double* x0_dev = nullptr;          // x0_host == NULL: the caller filled x_out_dev[0][d][N] with per-sample starting points
    if (x0_host) {
        BSDE_TRY(c->tmp.ensure(sizeof(double) * d));
        BSDE_CUDA_OK(cudaMemcpyAsync(c->tmp.p, x0_host, sizeof(double) * d, cudaMemcpyHostToDevice, c->stream));
        BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
        x0_dev = c->tmp.as<double>();
    }
    const double dt = T / steps;
    BSDE_TRY(launch_paths(model_id, sigma0, sqrt(dt), d, N, steps, x0_dev, seed, sample_offset, replay, xi_dev,
                          x_out_dev, c->stream));
    c->launches++;
    return 0;
}

int bsde_terminal(bsde_ctx* c, int model_id, const double* params, int d, int64_t N, const double* x_dev, double* g_out_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    double r, sigma0;
    BSDE_TRY(model_params(model_id, params, r, sigma0));
    BSDE_TRY(launch_terminal(model_id, d, N, x_dev, g_out_dev, c->stream));
    c->launches++;
    return 0;
}

int bsde_target(bsde_ctx* c, int model_id, const double* params, int d, int64_t N, const double* x_dev, const double* y_dev,
                const double* y_next_dev, const double* grad_dev, const double* xi_dev, double dt, int apply_sigma,
                double* out_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    double r, sigma0;
    BSDE_TRY(model_params(model_id, params, r, sigma0));
    if (!x_dev || !y_dev || !y_next_dev || !grad_dev || !out_dev) return fail_msg(2, "target: null pointer");
    BSDE_TRY(launch_target(model_id, r, sigma0, dt, d, N, x_dev, y_dev, y_next_dev, grad_dev, xi_dev, apply_sigma, out_dev, c->stream));
    c->launches++;
    return 0;
}

int bsde_mean(bsde_ctx* c, const double* v_dev, int64_t N, double* out_host) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    if (!v_dev || !out_host || N < 1) return fail_msg(2, "mean: bad arguments");
    const int nb = mean_grid(N);
    BSDE_TRY(c->tmp.ensure(sizeof(double) * (nb + 1)));
    BSDE_TRY(launch_block_sums(v_dev, N, c->tmp.as<double>(), c->stream));
    BSDE_TRY(launch_sum_partials(c->tmp.as<double>(), nb, 1, 1, c->tmp.as<double>() + nb, c->stream));
    c->launches += 2;
    double s = 0.0;
    BSDE_CUDA_OK(cudaMemcpyAsync(&s, c->tmp.as<double>() + nb, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    *out_host = s / (double)N;
    return 0;
}

int bsde_normal_eq(bsde_ctx* c, int basis_kind, int d, int p, int64_t N, const double* inputs_dev, const double* coef_host,
                   const double* y_dev, const int* ranks, const double* cores_host, int j, double* A_host, double* b_host) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_TRY(check_ranks(d, p, ranks));
    if (j < 0 || j >= d) return fail_msg(2, "normal_eq: core index %d outside 0..%d", j, d - 1);
    if (ranks[j] > 8 || ranks[j + 1] > 8) return fail_msg(2, "normal_eq: ranks <= 8 supported");
    const int n = ranks[j] * p * ranks[j + 1];
    if (n > kMaxN) return fail_msg(2, "normal_eq: n = %d exceeds %d", n, kMaxN);
    Inputs in;
    BSDE_TRY(make_inputs(c, in, basis_kind, d, p, N, inputs_dev, nullptr, coef_host));
    Train T;
    BSDE_TRY(train_upload(c, T, d, p, ranks, nullptr, cores_host));
    Bonds B;
    BSDE_TRY(bonds_alloc(c, T, N, B));
    BSDE_TRY(build_right(c, T, in, B, j + 1));       // bond[k] = R_{k-1} for k > j
    BSDE_TRY(build_left(c, T, in, B, j));            // bond[k] = L_k for k <= j
    int* info;
    BSDE_TRY(info_reset(c, &info));
    BSDE_TRY(c->dbg.ensure(sizeof(double) * ((size_t)n * n + n)));
    MicroCtx m{&T, &in, &B, y_dev, SOLVER_LSTSQ, basis_kind == BASIS_POLY ? 1 : 0, info};
    BSDE_TRY(micro_step(c, m, j, 0, 0.0, nullptr, c->dbg.as<double>()));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (A_host) BSDE_CUDA_OK(cudaMemcpy(A_host, c->dbg.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost));
    if (b_host) BSDE_CUDA_OK(cudaMemcpy(b_host, c->dbg.as<double>() + (size_t)n * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}

int bsde_solve(bsde_ctx* c, int n, const double* A_host, const double* b_host, int solver_id, double* x_host, int* info_host) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    if (n < 1 || n > kMaxN) return fail_msg(2, "solve: n = %d outside 1..%d", n, kMaxN);
    if (solver_id < 0 || solver_id > 3) return fail_msg(2, "Unknown method %d", solver_id);
    if (!A_host || !b_host || !x_host) return fail_msg(2, "solve: null pointer");
    BSDE_TRY(c->dbg.ensure(sizeof(double) * ((size_t)n * n + 2 * n)));
    double* A = c->dbg.as<double>();
    double* b = A + (size_t)n * n;
    double* x = b + n;
    BSDE_CUDA_OK(cudaMemcpyAsync(A, A_host, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaMemcpyAsync(b, b_host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    int* info;
    BSDE_TRY(info_reset(c, &info));
    MicroSolveArgs a{};
    a.rl = n; a.p = 1; a.rr = 1; a.direction = 0; a.solver = solver_id;
    a.core_j = x; a.info = info; a.A_direct = A; a.b_direct = b;
    a.force_onesided = (int)c->svd_onesided;
    if (const size_t ws = micro_solve_workspace_bytes(n, 0)) {      // systems beyond the solver's shared memory
        BSDE_TRY(c->solve_ws.ensure(ws));
        a.big_ws = c->solve_ws.as<double>();
    }
    BSDE_TRY(launch_micro_solve(a, c->stream));
    c->launches++;
    BSDE_CUDA_OK(cudaMemcpyAsync(x_host, x, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (solver_id == SOLVER_LSTSQ && n > 100) {
        // beyond 100 unknowns only the positive semi-definite path exists (the Jacobi fallbacks keep a second n x n matrix
        // in shared memory): an indefinite or non-symmetric system comes back as NaN from the kernel — say so
        bool bad = false;
        for (int i = 0; i < n; i++) bad = bad || std::isnan(x_host[i]);
        if (bad) return fail_msg(2, "solve: 'lstsq' of a system that is not symmetric positive semi-definite (or not finite) is supported up to 100 unknowns, n = %d", n);
    }
    if (info_host) {
        BSDE_CUDA_OK(cudaMemcpy(info_host, info, sizeof(int) * 4, cudaMemcpyDeviceToHost));
        if (info_host[3] == INT_MAX) info_host[3] = 0;
    }
    return 0;
}

int bsde_orthonormalize(bsde_ctx* c, int d, int p, const int* ranks, double* cores_host, int mode, int start, int stop) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    BSDE_TRY(check_ranks(d, p, ranks));
    if (mode != 0 && mode != 1) return fail_msg(2, "orthogonality must be either 'left' or 'right'");
    if (stop == -1) stop = d - 1;
    if (start < 0 || stop > d - 1) return fail_msg(2, "orthonormalize: range %d..%d outside the train", start, stop);
    Train T;
    BSDE_TRY(train_upload(c, T, d, p, ranks, nullptr, cores_host));
    if (mode == 1) BSDE_TRY(launch_orthonormalize_right(T.cores, T.off_dev, T.ranks_dev, p, d, start, stop, c->stream));
    else BSDE_TRY(launch_orthonormalize_left(T.cores, T.off_dev, T.ranks_dev, p, d, start, stop, c->stream));
    c->launches++;
    BSDE_TRY(train_download(c, T, cores_host));
    return 0;
}

int bsde_salsa_move(bsde_ctx* c, int side, int ra, int p, int k, int rb, double min_sv, int expand, int do_reg,
                    double* core_a_host, double* core_b_host, double* reg_out_host, int* rank_out, int* adapted_out) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    if (side != 0 && side != 1) return fail_msg(2, "salsa_move: side must be 0 (left move) or 1 (right move)");
    if (!core_a_host || !core_b_host) return fail_msg(2, "salsa_move: null core pointer");
    if (ra < 1 || p < 1 || p > kMaxP || k < 1 || k > 8 || rb < 1 || ra > 8 || rb > 8) return fail_msg(2, "salsa_move: shape (%d,%d,%d)-(%d) outside the kernels' range", ra, p, k, rb);
    if (side == 0 && expand && k >= 8) return fail_msg(2, "salsa_move: bond rank %d cannot grow beyond 8", k);
    const int kcap = k + 1;                                   // room for one more rank
    const size_t na = (size_t)ra * p * kcap, nb = (size_t)kcap * p * rb;
    BSDE_TRY(c->dbg.ensure(sizeof(double) * (na + nb + 16 + 2) + 64));
    double* a_dev = c->dbg.as<double>();
    double* b_dev = a_dev + na;
    double* reg_dev = b_dev + nb;                             // gamma / theta: up to 9 values
    double* msv_dev = reg_dev + 12;
    int* flags_dev = (int*)(msv_dev + 2);                     // [0] adapted, [1] rank
    BSDE_CUDA_OK(cudaMemsetAsync(a_dev, 0, sizeof(double) * (na + nb + 16 + 2) + 64, c->stream));
    BSDE_CUDA_OK(cudaMemcpyAsync(a_dev, core_a_host, sizeof(double) * (size_t)ra * p * k, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaMemcpyAsync(b_dev, core_b_host, sizeof(double) * (size_t)k * p * rb, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaMemcpyAsync(msv_dev, &min_sv, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int init_flags[2] = {0, k};
    BSDE_CUDA_OK(cudaMemcpyAsync(flags_dev, init_flags, sizeof init_flags, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (side == 0) BSDE_TRY(launch_salsa_left(a_dev, b_dev, ra, p, k, rb, msv_dev, expand ? 1 : 0, do_reg ? 1 : 0, reg_dev, flags_dev + 1, flags_dev, c->stream));
    else BSDE_TRY(launch_salsa_right(a_dev, b_dev, ra, p, k, rb, msv_dev, do_reg ? 1 : 0, reg_dev, c->stream));
    c->launches++;
    int flags[2];
    BSDE_CUDA_OK(cudaMemcpyAsync(flags, flags_dev, sizeof flags, cudaMemcpyDeviceToHost, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    const int knew = (side == 0 && flags[0]) ? flags[1] : k;
    if (knew < k || knew > kcap) return fail_msg(3, "salsa_move: kernel reported bond rank %d", knew);
    BSDE_CUDA_OK(cudaMemcpyAsync(core_a_host, a_dev, sizeof(double) * (size_t)ra * p * knew, cudaMemcpyDeviceToHost, c->stream));
    BSDE_CUDA_OK(cudaMemcpyAsync(core_b_host, b_dev, sizeof(double) * (size_t)knew * p * rb, cudaMemcpyDeviceToHost, c->stream));
    if (reg_out_host && do_reg) BSDE_CUDA_OK(cudaMemcpyAsync(reg_out_host, reg_dev, sizeof(double) * knew, cudaMemcpyDeviceToHost, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (rank_out) *rank_out = knew;
    if (adapted_out) *adapted_out = side == 0 ? flags[0] : 0;
    return 0;
}

int bsde_orthonormalize_shape(bsde_ctx* c, int d, const int* shape, const int* ranks, double* cores_host, int mode, int start, int stop) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    if (!shape || !cores_host) return fail_msg(2, "orthonormalize: null pointer");
    int pmax = 1;
    for (int j = 0; j < d; j++) {
        if (shape[j] < 1 || shape[j] > kMaxP) return fail_msg(2, "basis size shape[%d] = %d outside 1..%d", j, shape[j], kMaxP);
        pmax = std::max(pmax, shape[j]);
    }
    BSDE_TRY(check_ranks(d, pmax, ranks));
    if (mode != 0 && mode != 1) return fail_msg(2, "orthogonality must be either 'left' or 'right'");
    if (stop == -1) stop = d - 1;
    if (start < 0 || stop > d - 1) return fail_msg(2, "orthonormalize: range %d..%d outside the train", start, stop);
    // cores packed back to back with their own basis sizes: offsets, ranks and shape go to the device next to them
    std::vector<int64_t> off((size_t)d);
    int64_t total = 0;
    for (int j = 0; j < d; j++) { off[j] = total; total += (int64_t)ranks[j] * shape[j] * ranks[j + 1]; }
    BSDE_TRY(c->cores.ensure(sizeof(double) * (size_t)total));
    BSDE_TRY(c->meta.ensure(sizeof(int64_t) * d + sizeof(int) * (2 * d + 1) + 64));
    int64_t* off_dev = c->meta.as<int64_t>();
    int* ranks_dev = (int*)(off_dev + d);
    int* shape_dev = ranks_dev + d + 1;
    BSDE_CUDA_OK(cudaMemcpyAsync(off_dev, off.data(), sizeof(int64_t) * d, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaMemcpyAsync(ranks_dev, ranks, sizeof(int) * (d + 1), cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaMemcpyAsync(shape_dev, shape, sizeof(int) * d, cudaMemcpyHostToDevice, c->stream));
    BSDE_CUDA_OK(cudaMemcpyAsync(c->cores.p, cores_host, sizeof(double) * (size_t)total, cudaMemcpyHostToDevice, c->stream));
    if (mode == 1) BSDE_TRY(launch_orthonormalize_right(c->cores.as<double>(), off_dev, ranks_dev, pmax, d, start, stop, c->stream, shape_dev));
    else BSDE_TRY(launch_orthonormalize_left(c->cores.as<double>(), off_dev, ranks_dev, pmax, d, start, stop, c->stream, shape_dev));
    c->launches++;
    BSDE_CUDA_OK(cudaMemcpyAsync(cores_host, c->cores.p, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, c->stream));
    BSDE_CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int bsde_basis_eval(bsde_ctx* c, int basis_kind, int p, int64_t N, const double* x_dev, const double* coef_host, int deriv,
                    double* out_dev) {
    BSDE_CHECK_CTX(c);
    Guard g(c);
    if (basis_kind != BSDE_BASIS_POLY && basis_kind != BSDE_BASIS_LEGENDRE) return fail_msg(2, "basis_eval: basis_kind %d has no evaluator", basis_kind);
    if (deriv < 0 || deriv > 2) return fail_msg(2, "basis_eval: deriv = %d", deriv);
    if (p < 1 || p > kMaxP) return fail_msg(2, "basis size p = %d outside 1..%d", p, kMaxP);
    Inputs in;
    BSDE_TRY(make_inputs(c, in, basis_kind, 1, p, N, x_dev, nullptr, coef_host));
    BSDE_TRY(launch_basis_eval(in, deriv, out_dev, c->stream));
    c->launches++;
    return 0;
}

}  // extern "C"
