// evo_runtime_internal.cuh -- declarations shared by the translation units of libevostencils_b200.so:
// evo_runtime.cu (C-ABI, memory, solver graphs) and evo_dispatch_inst.cu (statement -> kernel launch, compiled once
// per (scalar type, dimension, number of fields) so that the library builds in parallel).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "evo_kernels.cuh"
#include "evo_kernels_star.cuh"
#include "evo_kernels_rbcol.cuh"
#include "evo_kernels_rrcol.cuh"
#include "evo_kernels_warp2d.cuh"
#include "evo_kernels_small.cuh"
#include "evo_kernels_fas.cuh"
#include "evo_kernels_helm.cuh"

using namespace evo;

// records the message evo_last_error() returns and hands back `code` (defined in evo_runtime.cu)
int fail(int code, const char *fmt, ...);
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(e_ == cudaErrorMemoryAllocation ? EVO_ERR_OOM : EVO_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                   \
    } while (0)
#define EV(call)                     \
    do {                             \
        int rc_ = (call);            \
        if (rc_ != EVO_OK) return rc_; \
    } while (0)


// ------------------------------------------------------------------------------------------------
struct evo_problem {
    evo_problem_desc desc;
    int words;
    int sm_count;
    Geom geom[EVO_MAX_LEVELS];
    TransferW R, P;
    void *init_sol[EVO_MAX_FIELDS];  // pristine finest-level SOL (incl. boundary values), padded layout
    void *rhs0[EVO_MAX_FIELDS];      // finest-level RHS, read-only, shared by all cycles
    std::vector<std::pair<void *, size_t>> pool;  // recycled cycle work slabs
    size_t device_bytes = 0;                      // total memory of the device (bound of the pool)
    int slab_world = 1, slab_rank = 0, slab_lc = 0;  // domain decomposition (slab_lc = 0: none)
    int own_g0[EVO_MAX_LEVELS], own_g1[EVO_MAX_LEVELS];   // owned global plane range per distributed level
    struct CycleRes { cudaStream_t stream; cudaEvent_t ev0, ev1; SolveState *h_state; double *h_hist, *d_hist; int hist_cap; };
    std::vector<CycleRes> res_pool;               // streams / events / pinned buffers of destroyed cycles, recycled
    int live_cycles = 0;                          // cycles still referring to this problem
    bool closed = false;                          // evo_problem_destroy called while cycles were alive
};

struct LevelMem {
    void *buf[EVO_BUF_COUNT][EVO_MAX_FIELDS];
    void *slot[EVO_MAX_FIELDS];  // [next] slot of SOL for `with jacobi` statements
    bool swapped[EVO_MAX_FIELDS];
};

struct evo_cycle {
    evo_problem *p;
    std::vector<evo_op> ops;
    OpSten sten[EVO_MAX_LEVELS];
    bool has_sten[EVO_MAX_LEVELS];
    LevelMem lv[EVO_MAX_LEVELS];
    void *slab;
    size_t slab_bytes;                // bytes in use (what a reset clears)
    size_t slab_cap = 0;              // bytes allocated (a recycled slab may be a little larger)
    size_t zero_bytes = 0;            // a reset clears [slab, slab + zero_bytes); the constant tables below lie behind it
    OpSten *d_run_sten = nullptr;     // device copies for the fused runs: sten[level], descriptors of the smoothers (per op),
    SmoothParams *d_run_sp = nullptr; //   restriction / prolongation weights
    TransferW *d_run_rp = nullptr;
    void *krylov[8][EVO_MAX_FIELDS];  // coarsest-level Krylov vectors
    void *scratch[EVO_MAX_FIELDS];    // finest-level scratch field (Richardson)
    void *helm[9];                    // Helmholtz outer solver: x, r, p, ap, s, t, h, rhat, row sums
    helm::HelmState *d_helm;
    OpSten helm_A;                    // un-shifted operator of the finest level
    cudaGraph_t helm_graph;
    cudaGraphExec_t helm_exec;
    double helm_tol;
    int helm_max_iters;
    OpSten helm_A_graph;              // the operator baked into helm_exec (by-value kernel argument)
    SolveState *d_state;
    double *d_hist;
    int hist_cap;
    double *d_partials;
    int n_partials;
    int *d_cg_iters;
    SolveState *h_state;  // pinned
    double *h_hist;       // pinned
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    // captured solver graph (valid for one (tol, max_iters))
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    double graph_tol;
    int graph_max_iters;
    int64_t kernels_per_cycle, kernels_prologue;
    int64_t launch_counter;  // counts kernel launches while enqueueing
    bool coarse_sol_written = false;   // evo_cycle_set_field wrote SOL of a correction level: its boundary layer may be non-zero,
                                       // `SOL = 0` stays a full memset from then on (no folding into the restriction)
    bool fuse_zero = false;  // set around a restriction whose next statement zeroes SOL of the coarse level: a kernel that
                             // also stores those zeros clears the flag (the memset node is then skipped)
    // While a solver-graph body is captured: the final reduction of the convergence norm also does the outer loop's
    // bookkeeping and sets the loop's conditional handles (one launch instead of three).
    struct Finish {
        bool on = false;
        double tol = 0.0;
        int max_iters = 0, mode = 0, n_handles = 0;
        cudaGraphConditionalHandle h[2] = {0, 0};
    } fin;
    bool use_while_graph;
    bool pingpong = false;                               // the WHILE body holds two cycles (see build_solver_graph)
    bool odd_swap[EVO_MAX_LEVELS][EVO_MAX_FIELDS] = {};  // levels whose SOL ends in the [next] slot after one cycle
    bool pristine = false;       // freshly reset: evo_cycle_solve need not reset again
    bool part_no_swap = false;   // partial execution of an out-of-place statement: leave SOL / [next] unexchanged
    int zc_lo = -1, zc_hi = -1;  // plane range override of the statement's destination level (domain decomposition)
    bool own_stream = true;
    bool res_dead_on_entry;  // the cycle overwrites RES@finest before reading it: the solver's own residual
                             // (convergence test) need not be stored
};

template <typename T> static inline Fields<T> fields_of(void *const *p, int nf)
{
    Fields<T> f;
    for (int i = 0; i < EVO_MAX_FIELDS; ++i) f.p[i] = i < nf ? (T *)p[i] : nullptr;
    return f;
}

static inline dim3 row_grid(const Geom &g, int per_thread = 1)
{
    int inner = g.n - 2;
    int threads = (inner + per_thread - 1) / per_thread;
    return dim3((threads + BX - 1) / BX, inner, g.dim == 3 ? g.zhi - g.zlo + 1 : 1);
}


static inline bool slab_level(const evo_problem *p, int level) { return p->slab_lc > 0 && level >= p->slab_lc; }

// 1 / (1 - i k h) of the Robin closure (defined in evo_runtime.cu)
cplx helm_rden(const evo_problem *p, int level);

// statement dispatch, instantiated in evo_dispatch_inst.cu for (double,2,1) (double,2,2) (double,3,1) (double,3,2) (cplx,2,1)
template <typename T, int DIM, int NF> int enqueue_op(evo_cycle *c, const evo_op &op, cudaStream_t s);
template <typename T, int DIM, int NF> int op_residual(evo_cycle *c, int level, bool norm, cudaStream_t s);
template <typename T, int DIM, int NF> int op_restrict(evo_cycle *c, const evo_op &op, cudaStream_t s);
template <typename T, int DIM, int NF> int op_reduce_rows(evo_cycle *c, int ni, cudaStream_t s);
template <typename T, int DIM, int NF> bool run_eligible(const evo_cycle *c, const evo_op &op);
template <typename T, int DIM, int NF> int enqueue_run(evo_cycle *c, const evo_op *ops, int n, cudaStream_t s);
