/*
 * evostencils_b200.h -- C-ABI of the B200-native multigrid fitness-evaluation backend.
 *
 * This library replaces everything *below* the method boundary of the reference's
 * ProgramGenerator (reference: evostencils/code_generation/exastencils.py:39).  In the
 * reference that boundary is crossed by three subprocesses (java code generator :397-403,
 * make :413, the generated solver binary :425-429); here it is crossed by the entry points
 * declared below.  Every entry point cites the reference interface it stands in for.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types in any signature
 *   - all functions return 0 on success, a negative evo_status otherwise;
 *     evo_last_error() returns a thread-local human readable message
 *   - a "level" l is a node-based grid with (2^l + 1)^dim nodes including the Dirichlet
 *     boundary layer, spacing h = 2^-l (reference: exastencils.py:97-103)
 *   - host-side field arrays are dense, x (= i0) fastest, (2^l+1)^dim entries per field,
 *     `scalar_words` doubles per entry (1 real, 2 complex); the padded device layout is private
 */
#ifndef EVOSTENCILS_B200_H
#define EVOSTENCILS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVO_ABI_VERSION 1
#define EVO_MAX_FIELDS 2
#define EVO_MAX_UNKNOWNS 8   /* reference: optimization/program.py:107 (maximum_local_system_size) */
#define EVO_MAX_DIM 3
#define EVO_MAX_LEVELS 16
#define EVO_STENCIL_POINTS 27 /* 3^3; 2-D problems use the 9 entries of the plane dz = 0 */

/* ---------------------------------------------------------------- status codes */
enum evo_status {
    EVO_OK = 0,
    EVO_ERR_INVALID = -1,   /* malformed descriptor / op list                          */
    EVO_ERR_CUDA = -2,      /* CUDA runtime error (message in evo_last_error)          */
    EVO_ERR_NO_DEVICE = -3, /* no CUDA device: the product path has NO CPU fallback    */
    EVO_ERR_UNSUPPORTED = -4,
    EVO_ERR_OOM = -5
};

/* ---------------------------------------------------------------- buffers
 * Field naming of the reference (exastencils.py:14-26, :295-316):
 *   SOL  = <field>@finest, gen_error_<field> below the finest level
 *   RHS  = <rhs_name>            RES = gen_residual_<field>
 *   COR  = gen_error_<field>     (aliases SOL below the finest level; the lowering resolves that)
 *   APX  = FAS restricted fine solution (exastencils_FAS.py:121-136)                     */
enum evo_buffer { EVO_BUF_SOL = 0, EVO_BUF_RHS = 1, EVO_BUF_RES = 2, EVO_BUF_COR = 3, EVO_BUF_APX = 4,
                  EVO_BUF_COUNT = 5,
                  EVO_BUF_NEXT = 100 /* evo_cycle_buffer only: the [next] slot of SOL */ };
#define EVO_PART_NO_SWAP 1

/* ---------------------------------------------------------------- op codes
 * One op == one statement the reference's emitter would print (exastencils.py:684-925);
 * codes >= 32 are fused forms produced only by the optimiser pass, never by the lowering.   */
enum evo_opcode {
    EVO_OP_ZERO = 1,          /* dst@level = 0                                   (:706-709) */
    EVO_OP_COPY = 2,          /* dst@level = src@level                           (:897-911) */
    EVO_OP_RESIDUAL = 3,      /* RES = RHS - A*SOL                               (:837-853) */
    EVO_OP_RICHARDSON = 4,    /* SOL_i += w*(RHS_i - sum_j A_ij SOL_j), i in order (:710-726) */
    EVO_OP_SMOOTH = 5,        /* one `solve locally` statement                   (:769-822) */
    EVO_OP_RESTRICT = 6,      /* dst@(level-1) = R@level * src@level             (:855-873) */
    EVO_OP_PROLONG_ADD = 7,   /* SOL@level += w * (P@(level-1) * src@(level-1))  (:727-743) */
    EVO_OP_PROLONG_SET = 8,   /* dst@level  = P@(level-1) * src@(level-1)        (:868-873) */
    EVO_OP_COARSE_SOLVE = 9,  /* gen_mgCycle@min(): Krylov solve, zero guess     (:874-896) */
    /* FAS (exastencils_FAS.py:99-319) */
    EVO_OP_FAS_RESTRICT_SOL = 10, /* APX@(level-1) = R*SOL@level ; SOL@(level-1) = APX     (:121-136) */
    EVO_OP_FAS_COARSE_RHS = 11,   /* RHS@(level-1) = R*RES@level + N(APX)@(level-1)        (:138-147) */
    EVO_OP_FAS_SUB_APX = 12,      /* SOL@level -= APX@level                                (:173-183) */
    /* fused forms */
    EVO_OP_RESIDUAL_RESTRICT = 32, /* RHS@(level-1) = R*(RHS - A*SOL)@level, RES not stored   */
    EVO_OP_SMOOTH_FUSED = 33       /* `count` identical pointwise sweeps in one pass           */
};

/* smoother update mode of a `solve locally` statement */
enum evo_smooth_mode {
    EVO_SMOOTH_JACOBI = 0,   /* `with jacobi`: read old slot, write new slot, advance (exastencils.py:780-783) */
    EVO_SMOOTH_REDBLACK = 1, /* `color with {(i0+..)%2}`: colour 0 first, in place    (:659-667, :785-790)    */
    EVO_SMOOTH_LEX = 2       /* in place, lexicographic (model-based mode, :64-70)                              */
};

enum evo_smooth_kind {
    EVO_KIND_LINEAR = 0,
    EVO_KIND_FAS_PICARD = 1, /* u += w (f - (A u + N(u)u)) / a00           (exastencils_FAS.py:226-232) */
    EVO_KIND_FAS_NEWTON = 2  /* ... / (a00 + J(u)), `count` inner steps     (:218-231)                   */
};

typedef struct evo_op {
    int32_t code;       /* evo_opcode                                                         */
    int32_t level;      /* level of the written field (fine level for RESTRICT / PROLONG_*)   */
    int32_t dst;        /* evo_buffer                                                         */
    int32_t src;        /* evo_buffer                                                         */
    int32_t mode;       /* evo_smooth_mode                                                    */
    int32_t kind;       /* evo_smooth_kind                                                    */
    int32_t n_unknowns; /* size of the local system of a SMOOTH statement (1..8)              */
    int32_t count;      /* repetitions (fused sweeps / Newton steps) or max Krylov iterations */
    int32_t unk_field[EVO_MAX_UNKNOWNS];           /* field index of each unknown             */
    int32_t unk_off[EVO_MAX_UNKNOWNS][EVO_MAX_DIM];/* node offset of each unknown             */
    double omega;       /* relaxation factor                                                  */
    double tol;         /* Krylov relative residual target                                    */
} evo_op;

/* Rediscretised system operator of one level: coef[i][j][p][w], p = (dz+1)*9+(dy+1)*3+(dx+1),
 * w = 0 real / 1 imaginary part.  A*u at a node = sum_j sum_p coef[i][j][p] * u_j[node+off_p],
 * accumulated in ascending (j, p) order (the oracle and the kernels use the same order).       */
typedef struct evo_level_operator {
    int32_t level;
    int32_t pad_;
    double coef[EVO_MAX_FIELDS][EVO_MAX_FIELDS][EVO_STENCIL_POINTS][2];
} evo_level_operator;

enum evo_problem_kind {
    EVO_PROBLEM_LINEAR = 0,    /* Poisson 2D/3D, LinearElasticity (example_problems/ *.exa2)        */
    EVO_PROBLEM_FAS = 1,       /* -Lap u + gamma u e^u = f  (FAS_2D_Basic_template.exa4:19-34)     */
    EVO_PROBLEM_HELMHOLTZ = 2  /* complex, Robin x-boundaries (Helmholtz/...exa4:25-145)           */
};

typedef struct evo_problem_desc {
    int32_t abi_version;  /* EVO_ABI_VERSION                                             */
    int32_t dim;          /* 2 or 3          (parser.py:114-125, `dimensionality`)       */
    int32_t n_fields;     /* 1 or 2                                                      */
    int32_t scalar_words; /* 1 real fp64, 2 complex fp64                                 */
    int32_t min_level;    /* coarsest level  (knowledge `minLevel`)                      */
    int32_t max_level;    /* finest level    (knowledge `maxLevel`)                      */
    int32_t kind;         /* evo_problem_kind                                            */
    int32_t device;       /* CUDA device ordinal                                         */
    double gamma;         /* FAS nonlinearity factor                                     */
    double k_re, k_im;    /* Helmholtz: Robin factor uses wave number k (exa4:43-108)    */
    double restrict_w[EVO_STENCIL_POINTS]; /* R weights, offset index p as above (reads fine node 2*I+o) */
    double prolong_w[EVO_STENCIL_POINTS];  /* P weights (fine node x gets w[o]*coarse[(x+o)/2], x+o even) */
} evo_problem_desc;

typedef struct evo_solve_params {
    double tol;          /* solver_targetResReduction (Poisson/2D...exa3:3)                     */
    int32_t max_iters;   /* solver_maxNumIts (:4)                                               */
    int32_t samples;     /* evaluation_samples (exastencils.py:417-443): timing repeats          */
    int32_t flags;       /* EVO_SOLVE_* bits                                                    */
    int32_t timeout_ms;  /* > 0: a solve that runs longer is stopped at the next iteration boundary by the device-side
                            loop and reports status 2 -- the reference kills the binary after `evaluation_timeout` and
                            returns (infinity,)*3 (exastencils.py:430-433, :476-483); 0 = no limit                  */
} evo_solve_params;

#define EVO_SOLVE_NO_GRAPH 1   /* debugging: launch kernels directly, no CUDA-graph capture */
#define EVO_SOLVE_KEEP_STATE 2 /* do not reset SOL to the initial guess before solving      */
#define EVO_SOLVE_SOLO_TIMING 4 /* evo_batch_solve: time_ms of every converged member is re-measured with the GPU to itself
                                   (a few iterations, extrapolated), so that the time objective does not depend on what
                                   else was in flight -- the reference times one binary at a time (exastencils.py:417-443) */

typedef struct evo_solve_result {
    int32_t status;       /* 0 ok, 1 non-finite residual encountered, 2 stopped by the timeout_ms watchdog */
    int32_t iterations;   /* outer iterations executed                                          */
    double time_ms;       /* median solve time over `samples` runs, CUDA events (ms)            */
    double time_ms_min;
    double initial_residual;
    double final_residual;
    int64_t kernel_launches; /* kernels launched by one solve (graph nodes x iterations)         */
} evo_solve_result;

typedef struct evo_problem evo_problem;
typedef struct evo_cycle evo_cycle;

/* -- library -------------------------------------------------------------------------------- */
int evo_abi_version(void);
const char *evo_last_error(void);
/* number of visible CUDA devices; <0 on error.  ProgramGenerator.__init__ raises
 * RuntimeError("Compiler not found") when its toolchain is missing (exastencils.py:104-108);
 * the drop-in raises the same way when this returns <= 0.                                      */
int evo_device_count(void);
int evo_device_name(int device, char *buf, size_t len, int *sm_count);
/* Tuning switches (kernel variants / experiments; the defaults are the measured best).  `name` is the name of the
 * environment variable that initialises the switch (EVO_RB_VARIANT, EVO_RB_FUSE2, EVO_RR_VARIANT, ...; DESIGN.md
 * lists them).  A new value applies to solver graphs captured afterwards.  The reference's counterpart are the
 * knowledge-file switches patched per run (exastencils.py:217-266).                                             */
int evo_set_option(const char *name, int value);
int evo_get_option(const char *name, int *value);

/* -- problem = discretisation hierarchy + initial guess / rhs (what InitFields and the field
 *    declarations of the generated program hold; exastencils.py:586-592 generate_storage)      */
int evo_problem_create(const evo_problem_desc *desc, evo_problem **out);
int evo_problem_destroy(evo_problem *p);
/* upload the initial content of buffer `buf` (SOL incl. boundary values, or RHS) of `field` on
 * `level`; n_doubles must be (2^level+1)^dim * scalar_words                                    */
int evo_problem_set_field(evo_problem *p, int level, int buf, int field, const double *host, size_t n_doubles);

/* -- domain decomposition of ONE grid over several GPUs (the reference's generated code can
 *    split a grid into blocks/fragments but all shipped configurations use one, lib/domain_onePatch.knowledge).
 *    z-slabs with nested ownership: level `coarsest_distributed_level` splits its inner planes evenly over the
 *    ranks, a rank owning planes [a, b] of level l owns [2a-1, 2b] of level l+1 (the last rank also 2b+1);
 *    coarser levels are replicated on every rank.  Each slab carries 2 ghost planes per side.  Must be called
 *    before evo_problem_set_field / evo_cycle_build.  3-D real scalar problems only.                       */
int evo_problem_set_slab(evo_problem *p, int rank, int world, int coarsest_distributed_level);
/* ... with `ghost` planes per side instead of 2 (even, <= 16): wide ghost zones let consecutive sweeps recompute the
 * halo redundantly instead of exchanging it after every statement (communication-avoiding schedule, domain.py)  */
int evo_problem_set_slab_ex(evo_problem *p, int rank, int world, int coarsest_distributed_level, int ghost);
/* info[0..7] = zoff (global index of local plane 0), local plane count, first owned local plane, last owned local
 * plane, first owned global plane, last owned global plane, row pitch (entries), plane stride (entries)      */
int evo_problem_slab_info(evo_problem *p, int level, long long info[8]);

/* -- cycle = one lowered individual (replaces generate_cycle_function + java + make,
 *    exastencils.py:318-336, :381-415)                                                          */
int evo_cycle_build(evo_problem *p, const evo_op *ops, int n_ops, const evo_level_operator *operators,
                    int n_operators, evo_cycle **out);
int evo_cycle_destroy(evo_cycle *c);
/* run the op list `repeat` times on the cycle's working hierarchy (parity testing hook)        */
int evo_cycle_reset(evo_cycle *c);
int evo_cycle_apply(evo_cycle *c, int repeat);
int evo_cycle_get_field(evo_cycle *c, int level, int buf, int field, double *host, size_t n_doubles);
int evo_cycle_set_field(evo_cycle *c, int level, int buf, int field, const double *host, size_t n_doubles);
/* RES@finest = RHS - A*SOL and its L2 norm over inner nodes (gen_resNorm of the generated solver) */
int evo_cycle_residual_norm(evo_cycle *c, double *norm);
/* -- host-orchestrated execution (domain decomposition: halo exchanges happen between statements) ----------- */
/* use `cuda_stream` (a cudaStream_t, e.g. torch's current stream) for everything the cycle enqueues         */
int evo_cycle_set_stream(evo_cycle *c, void *cuda_stream);
/* enqueue statements (no synchronisation).  zc_lo/zc_hi >= 0 override the (local) plane range the statement
 * writes: coarse planes of RESTRICT, planes of RESIDUAL / PROLONG_ADD (to include ghost planes whose inputs
 * are valid and save an exchange)                                                                          */
int evo_cycle_exec_ops(evo_cycle *c, const evo_op *ops, int n_ops, int zc_lo, int zc_hi);
/* one statement on the local planes [z_lo, z_hi] (boundary planes first, interior while the halo travels);
 * EVO_PART_NO_SWAP: an out-of-place smoother does not yet exchange SOL and its [next] slot                   */
int evo_cycle_exec_part(evo_cycle *c, const evo_op *op, int z_lo, int z_hi, int flags);
/* current device address of a field (the two slots of SOL swap after out-of-place statements)               */
int evo_cycle_buffer(evo_cycle *c, int level, int buf, int field, void **device_ptr);
/* RES@finest = RHS - A SOL on the owned planes + canonical per-plane sums of |r|^2: device array of
 * `*count` doubles (one per owned plane, ascending)                                                        */
int evo_cycle_residual_plane_sums(evo_cycle *c, double **device_sums, int *count);
/* canonical vecsum (lane-strided + butterfly) of a device array, result on the host (synchronises)           */
int evo_cycle_vecsum(evo_cycle *c, const double *device_vals, int m, double *out);

/* stream-ordered pieces of the above for execution captured in a CUDA graph by the host (domain.py): the sum
 * stays on the device until evo_cycle_read_sum (which synchronises); evo_cycle_swap_slots exchanges SOL and its
 * [next] slot on the host side after an odd number of replays of a captured cycle                          */
int evo_cycle_vecsum_async(evo_cycle *c, const double *device_vals, int m);
int evo_cycle_read_sum(evo_cycle *c, double *out);
int evo_cycle_swap_slots(evo_cycle *c, int level);

/* measurement hook (bench.py roofline, measured-cost surrogate): the kernels of ONE statement `repeat`
 * times as nodes of one CUDA graph on the cycle's stream (what the statement costs inside the solver
 * graph), bracketed by CUDA events; returns the average milliseconds per execution and the
 * number of kernel launches one execution makes.  The reference's counterpart is the per-function
 * timer output of the generated binary (printAllTimers, Helmholtz/...exa4:13-19).                 */
int evo_cycle_profile_op(evo_cycle *c, const evo_op *op, int repeat, double *ms_per_exec, int64_t *launches_per_exec);

/* -- solve = the generated solver's outer loop (a9) run `samples` times (a7: evaluate,
 *    exastencils.py:417-443): res0 = ||f - A u0||; repeat { cycle; res = ||f - A u|| } until
 *    res < tol*res0 or max_iters.  res_hist receives res0, res1, ... (max_iters+1 entries).     */
int evo_cycle_solve(evo_cycle *c, const evo_solve_params *params, evo_solve_result *result, double *res_hist);
/* Helmholtz (complex) problems: the outer PreconditionedBiCGStab@finest of the reference's problem file
 * (example_problems/Helmholtz/2D_FD_Helmholtz_fromL3.exa3:144-200: stop |res| < tol |res0| or max_iters) on the
 * un-shifted operator `A` of the finest level, with the cycle (built on the shifted operator M) applied as
 * right preconditioner `u = 0; f = v; gen_mgCycle()` twice per iteration.  res_hist[k] = |curRes| after k
 * iterations.                                                                                          */
int evo_helmholtz_solve(evo_cycle *c, const evo_level_operator *A, const evo_solve_params *params, evo_solve_result *result,
                        double *res_hist);
/* many individuals in flight on one GPU, one stream each (population evaluation, program.py:491) */
int evo_batch_solve(evo_cycle **cycles, int n_cycles, const evo_solve_params *params, evo_solve_result *results,
                    double *res_hist /* n_cycles * (max_iters+1) */, double *batch_ms);
/* The same, as a pipeline the host drives: evo_cycle_solve_begin captures / instantiates the solver graph and enqueues
 * one complete solve on the cycle's own stream WITHOUT waiting; evo_cycle_solve_end waits for it and returns the result
 * (time_ms = the solve's span on the device, other solves may have been in flight); evo_cycle_solve_retime measures the
 * time of a finished solve again with nothing else in flight (EVO_SOLVE_SOLO_TIMING semantics: call it when the device
 * is idle) and updates result->time_ms.  A sliding window of begin / end calls keeps a constant number of
 * individuals in flight while the host lowers and builds the next ones.                                           */
int evo_cycle_solve_begin(evo_cycle *c, const evo_solve_params *params);
int evo_cycle_solve_end(evo_cycle *c, const evo_solve_params *params, evo_solve_result *result, double *res_hist);
int evo_cycle_solve_retime(evo_cycle *c, const evo_solve_params *params, evo_solve_result *result);

#ifdef __cplusplus
}
#endif
#endif /* EVOSTENCILS_B200_H */
