// evo_kernels_run.cuh -- a maximal run of consecutive statements on small levels executed by ONE kernel launch.
//
// On the latency-bound levels (2-D up to 257^2, 3-D up to 33^3) a statement is a few microseconds of work and a cycle
// is dozens of dependent kernel launches; a generation of 256 individuals is ~1.4 million launches and the GPU's
// launch rate, not its SMs, bounds the evaluation rate (north star: "launch overhead on the latency-bound coarse
// levels is amortised").  Here the host resolves a run of statements (buffers, geometry, stencils) into a table that
// travels as a __grid_constant__ kernel parameter, and one thread-block cluster (1..8 CTAs, hardware cluster barrier
// between statements, colours and repetitions) interprets it.  Per node every statement performs exactly the
// arithmetic of its stand-alone kernel (same device functions: apply_row, local_solve), so results are bit-identical
// whether a statement runs fused or alone.
#pragma once
#include "evo_kernels.cuh"

namespace evo {

constexpr int RUN_MAX_OPS = 32;
constexpr int RUN_MAX_STAGE = 20;        // field arrays a run keeps resident in shared memory
constexpr int RUN_MAX_LEVELS = 6;

struct RunOp {
    int code;                       // evo_opcode
    int li, lj;                     // geometry / stencil table index of the statement's level and of level - 1
    int reps;                       // smoothing repetitions
    int mode;                       // evo_smooth_mode
    unsigned written;               // bit i: field i is written by the smoother (Jacobi: it alternates between its two slots)
    int spi;                        // index of the smoother's descriptor in RunTable::sp
    int nu;                         // size of its local system
    double omega;
    void *a[EVO_MAX_FIELDS];        // smooth: current slot | residual: u | restrict: src (fine) | prolong: src (coarse) | copy: src
    void *b[EVO_MAX_FIELDS];        // smooth: [next] slot  | residual: f | restrict: dst (coarse) | prolong: dst (fine) | copy / zero: dst
    void *c[EVO_MAX_FIELDS];        // smooth: rhs          | residual: r | residual+restrict: f (fine; a = u, b = dst coarse)
};

// A field array of a run that lives in shared memory for the whole launch: copied in (compact, pitch = n) before the first
// statement, copied back after the last one.  The statements address it through tagged pointers (run_ptr).
struct RunStage {
    void *ptr;                      // the array in device memory (padded layout of its level)
    int li;                         // table index of its level
    int off;                        // offset in the dynamic shared memory, in doubles
};

// < 4 KB: the table is a kernel parameter (copied at every launch); the bulky constants stay in device memory
struct RunTable {
    int n;
    int lbase;                       // level of table index 0
    int n_stage;                     // > 0: shared-memory resident run (one CTA); geom[] then holds the COMPACT geometry
    int cg_off;                      // offset (doubles) of the CG vectors in the dynamic shared memory
    RunStage stage[RUN_MAX_STAGE];
    Geom geom[RUN_MAX_LEVELS];
    const OpSten *sten;              // [EVO_MAX_LEVELS], indexed by level
    const SmoothParams *sp;          // [number of statements of the cycle]
    const TransferW *rp;             // restriction, prolongation
    RunOp op[RUN_MAX_OPS];
};

// a statement's array: a device pointer, or (odd value) an offset into the run's shared memory
__host__ __device__ __forceinline__ void *run_tag(int off_doubles) { return (void *)(((uintptr_t)off_doubles << 4) | 1u); }
__device__ __forceinline__ double *run_ptr(void *p, double *smb)
{
    const uintptr_t v = (uintptr_t)p;
    return (v & 1u) ? smb + (v >> 4) : (double *)p;
}
template <typename T> __device__ __forceinline__ Fields<T> run_fields(void *const *p, double *smb)
{
    Fields<T> f;
#pragma unroll
    for (int i = 0; i < EVO_MAX_FIELDS; ++i) f.p[i] = (T *)run_ptr(p[i], smb);
    return f;
}

// inner node number t of a level -> coordinates
template <int DIM> __device__ __forceinline__ void run_node(int t, int ni, int &x, int &y, int &z)
{
    x = 1 + t % ni;
    const int r = t / ni;
    if (DIM == 3) { y = 1 + r % ni; z = 1 + r / ni; }
    else { y = 1 + r; z = 0; }
}

// barrier between dependent parts of a run: the whole cluster, or just the CTA when the run has one
template <bool CLUSTER> __device__ __forceinline__ void run_barrier()
{
    if constexpr (CLUSTER) cluster_barrier();
    else __syncthreads();
}

template <int DIM, int NF, int NU, bool CLUSTER>
__device__ __forceinline__ void run_smooth_nu(const RunTable &tab, const RunOp &op, int gtid, int nthreads, double *smb)
{
    const Geom &g = tab.geom[op.li];
    const OpSten &st = tab.sten[tab.lbase + op.li];
    const SmoothParams &sp = tab.sp[op.spi];
    const int ni = g.n - 2, count = ni * ni * (DIM == 3 ? ni : 1);
    const Fields<double> rhs = run_fields<double>(op.c, smb);
    if (op.mode == EVO_SMOOTH_JACOBI) {
        for (int rep = 0; rep < op.reps; ++rep) {
            Fields<double> src, dst;
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                const bool w = (op.written >> i) & 1u;
                double *cur = run_ptr(op.a[i], smb), *nxt = run_ptr(op.b[i], smb);
                src.p[i] = (w && (rep & 1)) ? nxt : cur;
                dst.p[i] = w ? ((rep & 1) ? cur : nxt) : cur;
            }
            for (int t = gtid; t < count; t += nthreads) {
                int x, y, z;
                run_node<DIM>(t, ni, x, y, z);
                local_solve<double, DIM, NF, NU>(g, st, sp, src, dst, rhs, x, y, z);
            }
            run_barrier<CLUSTER>();
        }
    } else {   // red-black, order independent: colour 0 then colour 1, in place
        const Fields<double> u = run_fields<double>(op.a, smb);
        const int pairs = (ni + 1) / 2, rows = DIM == 3 ? ni * ni : ni;
        for (int rep = 0; rep < op.reps; ++rep)
            for (int color = 0; color < 2; ++color) {
                for (int t = gtid; t < pairs * rows; t += nthreads) {
                    const int k = t % pairs, r = t / pairs;
                    const int y = 1 + (DIM == 3 ? r % ni : r), z = DIM == 3 ? 1 + r / ni : 0;
                    const int x = 1 + 2 * k + ((1 + y + z + color) & 1);
                    if (x <= ni) local_solve<double, DIM, NF, NU>(g, st, sp, u, u, rhs, x, y, z);
                }
                run_barrier<CLUSTER>();
            }
    }
}

// NUMAX: largest local system of the run's smoothers -- the register budget of the kernel is that of its most
// expensive path, so runs with pointwise smoothers only get their own lean instantiation
template <int DIM, int NF, int NUMAX, bool CLUSTER>
__device__ __forceinline__ void run_smooth(const RunTable &tab, const RunOp &op, int gtid, int nthreads, double *smb)
{
    const int nu = op.nu;
    if (nu == 1) run_smooth_nu<DIM, NF, 1, CLUSTER>(tab, op, gtid, nthreads, smb);
    if constexpr (NUMAX >= 2) { if (nu == 2) run_smooth_nu<DIM, NF, 2, CLUSTER>(tab, op, gtid, nthreads, smb); }
    if constexpr (NUMAX >= 4) {
        if (nu == 3) run_smooth_nu<DIM, NF, 3, CLUSTER>(tab, op, gtid, nthreads, smb);
        if (nu == 4) run_smooth_nu<DIM, NF, 4, CLUSTER>(tab, op, gtid, nthreads, smb);
    }
    if constexpr (NUMAX >= 8) {
        if (nu == 5) run_smooth_nu<DIM, NF, 5, CLUSTER>(tab, op, gtid, nthreads, smb);
        if (nu == 6) run_smooth_nu<DIM, NF, 6, CLUSTER>(tab, op, gtid, nthreads, smb);
        if (nu == 7) run_smooth_nu<DIM, NF, 7, CLUSTER>(tab, op, gtid, nthreads, smb);
        if (nu == 8) run_smooth_nu<DIM, NF, 8, CLUSTER>(tab, op, gtid, nthreads, smb);
    }
}

template <int DIM, int NF, int NUMAX, bool CLUSTER>
__global__ void __launch_bounds__(NUMAX <= 2 ? 1024 : 512) k_run(const __grid_constant__ RunTable tab)
{
    const int nthreads = CLUSTER ? (int)(cluster_nctarank() * blockDim.x) : (int)blockDim.x;
    const int gtid = CLUSTER ? (int)(cluster_ctarank() * blockDim.x + threadIdx.x) : (int)threadIdx.x;
    const TransferW &R = tab.rp[0], &P = tab.rp[1];
    extern __shared__ double run_sm[];               // resident field arrays, then the CG vectors
    double *const smb = run_sm;
    // copy one resident array between its padded device layout and the compact shared-memory layout
    auto stage_copy = [&](const RunStage &e, bool in) {
        const int n = tab.geom[e.li].n, gpitch = (n + 15) / 16 * 16;
        const long long gplane = (long long)gpitch * n;
        const int count = n * n * (DIM == 3 ? n : 1);
        double *gp = (double *)e.ptr, *sp = smb + e.off;
        for (int t = gtid; t < count; t += nthreads) {
            const int x = t % n, r = t / n, y = r % n, z = r / n;
            const long long gi = z * gplane + (long long)y * gpitch + x;
            if (in) sp[t] = gp[gi];
            else gp[gi] = sp[t];
        }
    };
    if constexpr (!CLUSTER) {
        if (tab.n_stage > 0) {
            for (int e = 0; e < tab.n_stage; ++e) stage_copy(tab.stage[e], true);
            __syncthreads();
        }
    }
    for (int q = 0; q < tab.n; ++q) {
        const RunOp &op = tab.op[q];
        const Geom &g = tab.geom[op.li];
        const int ni = g.n - 2, count = ni * ni * (DIM == 3 ? ni : 1);
        switch (op.code) {
        case EVO_OP_SMOOTH:
            run_smooth<DIM, NF, NUMAX, CLUSTER>(tab, op, gtid, nthreads, smb);   // ends with a barrier
            continue;
        case EVO_OP_ZERO:
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                double *d = run_ptr(op.b[i], smb);
                for (long long t = gtid; t < g.total; t += nthreads) d[t] = 0.0;
            }
            break;
        case EVO_OP_COPY:
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                double *d = run_ptr(op.b[i], smb);
                const double *s = run_ptr(op.a[i], smb);
                if (d != s)
                    for (long long t = gtid; t < g.total; t += nthreads) d[t] = s[t];
            }
            break;
        case EVO_OP_RESIDUAL: {
            const Fields<double> u = run_fields<double>(op.a, smb), f = run_fields<double>(op.b, smb), r = run_fields<double>(op.c, smb);
            for (int t = gtid; t < count; t += nthreads) {
                int x, y, z;
                run_node<DIM>(t, ni, x, y, z);
                const long long idx = node_index(g, x, y, z);
#pragma unroll
                for (int i = 0; i < NF; ++i) r.p[i][idx] = f.p[i][idx] - apply_row<double, NF>(g, tab.sten[tab.lbase + op.li], u, i, idx);
            }
            break;
        }
        case EVO_OP_RESTRICT:
        case EVO_OP_RESIDUAL_RESTRICT: {
            const Geom &gc = tab.geom[op.lj];
            const int nc = gc.n - 2, cc = nc * nc * (DIM == 3 ? nc : 1);
            const Fields<double> u = run_fields<double>(op.a, smb), dst = run_fields<double>(op.b, smb), f = run_fields<double>(op.c, smb);
            const bool fused = op.code == EVO_OP_RESIDUAL_RESTRICT;
            const OpSten &st = tab.sten[tab.lbase + op.li];
            // the fine value entering restriction entry p of coarse node (x, y, z)
            auto fine_value = [&](int i, int x, int y, int z, int p) {
                const int fx = 2 * x + R.ox[p], fy = 2 * y + R.oy[p], fz = DIM == 3 ? 2 * z + R.oz[p] : 0;
                if (!fused) return u.p[i][node_index(g, fx, fy, fz)];
                double rv = 0.0;  // the residual field is 0 on the boundary layer
                if (fx >= 1 && fx <= g.n - 2 && fy >= 1 && fy <= g.n - 2 && (DIM == 2 || (fz >= 1 && fz <= g.n - 2))) {
                    const long long idx = node_index(g, fx, fy, fz);
                    rv = f.p[i][idx] - apply_row<double, NF>(g, st, u, i, idx);
                }
                return rv;
            };
            if (cc <= 1024) {
                // tiny coarse grid: one warp per coarse node, lane p evaluates entry p, the products are added in ascending
                // p by every lane (k_residual_restrict_warp; a thread per node would chain 27 stencil evaluations)
                const int lane = threadIdx.x & 31, nwarps = nthreads >> 5;
                for (int t = gtid >> 5; t < cc; t += nwarps) {
                    int x, y, z;
                    run_node<DIM>(t, nc, x, y, z);
                    const long long cidx = node_index(gc, x, y, z);
#pragma unroll
                    for (int i = 0; i < NF; ++i) {
                        double term = 0.0;
                        if (lane < R.nnz) term = R.w[lane] * fine_value(i, x, y, z, lane);
                        double acc = 0.0;
                        for (int p = 0; p < R.nnz; ++p) acc = acc + shfl_idx(term, p);
                        if (lane == 0) dst.p[i][cidx] = acc;
                    }
                }
                break;
            }
            for (int t = gtid; t < cc; t += nthreads) {
                int x, y, z;
                run_node<DIM>(t, nc, x, y, z);
                const long long cidx = node_index(gc, x, y, z);
#pragma unroll
                for (int i = 0; i < NF; ++i) {
                    double acc = 0.0;
                    for (int p = 0; p < R.nnz; ++p) acc = acc + R.w[p] * fine_value(i, x, y, z, p);
                    dst.p[i][cidx] = acc;
                }
            }
            break;
        }
        case EVO_OP_PROLONG_ADD:
        case EVO_OP_PROLONG_SET: {
            const Geom &gc = tab.geom[op.lj];
            const Fields<double> src = run_fields<double>(op.a, smb), dst = run_fields<double>(op.b, smb);
            const bool add = op.code == EVO_OP_PROLONG_ADD;
            for (int t = gtid; t < count; t += nthreads) {
                int x, y, z;
                run_node<DIM>(t, ni, x, y, z);
                const long long idx = node_index(g, x, y, z);
#pragma unroll
                for (int i = 0; i < NF; ++i) {
                    double acc = 0.0;
                    for (int p = 0; p < P.nnz; ++p) {
                        const int cx = x + P.ox[p], cy = y + P.oy[p], cz = DIM == 3 ? z + P.oz[p] : 0;
                        if ((cx & 1) || (cy & 1) || (DIM == 3 && (cz & 1))) continue;
                        acc = acc + P.w[p] * src.p[i][node_index(gc, cx >> 1, cy >> 1, cz >> 1)];
                    }
                    if (add) dst.p[i][idx] = dst.p[i][idx] + op.omega * acc;
                    else dst.p[i][idx] = acc;
                }
            }
            break;
        }
        case EVO_OP_COARSE_SOLVE:
            // CG on the coarsest level by this CTA (runs with a coarse solve are launched as ONE CTA)
            if constexpr (!CLUSTER) {
                coarse_cg_smem<DIM, NF>(g, tab.sten[tab.lbase + op.li], run_fields<double>(op.a, smb), run_fields<double>(op.b, smb), op.reps,
                                        op.omega, (int *)op.c[0], run_sm + tab.cg_off);
            }
            break;
        default: break;
        }
        run_barrier<CLUSTER>();
    }
    if constexpr (!CLUSTER) {
        for (int e = 0; e < tab.n_stage; ++e) stage_copy(tab.stage[e], false);
    }
}

}  // namespace evo
