// evo_kernels.cuh -- generic (any 3^d constant stencil, 1-2 fields, real or complex) kernels of the
// multigrid evaluation path.  One kernel per statement kind of the reference's emitter
// (evostencils/code_generation/exastencils.py:684-925); the specialised high-bandwidth kernels for
// scalar star stencils live in evo_kernels_star.cuh and must produce bit-identical results.
//
// Arithmetic contract shared with the CPU oracle (oracle/mg_ops.inc):
//   A*u at a node   = sum_j sum_q c[i][j][q] * u_j[node + off_q], q ascending, acc = acc + c*u
//   pointwise solve = s = sum of the non-unknown terms (same order); x = (f - s) * (1/a); u += w (x - u)
#pragma once
#include "evo_common.cuh"

namespace evo {

constexpr int BX = 128;  // threads per block of the row-mapped kernels (x fastest)

__device__ __forceinline__ long long node_index(const Geom &g, int x, int y, int z)
{
    return (long long)z * g.plane + (long long)y * g.pitch + x;
}

template <typename T>
__device__ __forceinline__ T sten_coef(const Sten &s, int q)
{
    return scalar_traits<T>::make(s.re[q], s.im[q]);
}

template <typename T, int NF>
__device__ __forceinline__ T apply_row(const Geom &g, const OpSten &st, const Fields<T> &u, int i, long long idx)
{
    T acc = T(0.0);
#pragma unroll
    for (int j = 0; j < NF; ++j) {
        const Sten &s = st.s[i][j];
        const T *uj = u.p[j];
        for (int q = 0; q < s.nnz; ++q) {
            long long d = (long long)s.oz[q] * g.plane + (long long)s.oy[q] * g.pitch + s.ox[q];
            acc = acc + sten_coef<T>(s, q) * uj[idx + d];
        }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// gen_residual_<f>@l = RHS@l - (A@l * SOL@l)                          (exastencils.py:837-853)
template <typename T, int DIM, int NF>
__global__ void __launch_bounds__(BX) k_residual(const Geom g, const __grid_constant__ OpSten st, Fields<T> u,
                                                 Fields<T> f, Fields<T> r)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y, z = DIM == 3 ? 1 + blockIdx.z : 0;
    if (x > g.n - 2) return;
    const long long idx = node_index(g, x, y, z);
#pragma unroll
    for (int i = 0; i < NF; ++i) r.p[i][idx] = f.p[i][idx] - apply_row<T, NF>(g, st, u, i, idx);
}

// ---- canonical reduction order (shared with oracle/mg_ops.inc, which documents it) ---------------
//   vecsum(v[0..m)): lane (i mod 32) adds v[i] in ascending i, then xor butterfly 16,8,4,2,1
//   field sum      : vecsum_z( vecsum_y( vecsum_x(row) ) );  total: field sums added in field order
// A warp IS this order: lane-strided loop + __shfl_xor tree.  Every reduction of a solve (residual
// norm, Krylov dot products) uses it, which makes GPU and oracle residual histories bit-identical.
__device__ __forceinline__ double shfl_xor(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ cplx shfl_xor(cplx v, int o)
{
    return cplx(__shfl_xor_sync(0xffffffffu, v.re, o), __shfl_xor_sync(0xffffffffu, v.im, o));
}
__device__ __forceinline__ double shfl_idx(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ cplx shfl_idx(cplx v, int src)
{
    return cplx(__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src));
}
template <typename T> __device__ __forceinline__ T warp_butterfly(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = v + shfl_xor(v, o);
    return v;
}
// canonical vecsum of v[0..m) by one full warp (all lanes return the sum)
template <typename T> __device__ __forceinline__ T warp_vecsum(const T *v, int m)
{
    const int lane = threadIdx.x & 31;
    T acc = T(0.0);
    for (int i = lane; i < m; i += 32) acc = acc + v[i];
    return warp_butterfly(acc);
}
// canonical sum over one row of inner nodes of a*b (b == nullptr: |a|^2 as a real number)
template <typename T, typename R>
__device__ __forceinline__ R warp_row_dot(const T *a, const T *b, long long row_base /* index of x = 1 */, int ni)
{
    const int lane = threadIdx.x & 31;
    R acc = R(0.0);
    for (int i = lane; i < ni; i += 32) {
        if constexpr (sizeof(R) == sizeof(T)) acc = acc + (b ? a[row_base + i] * b[row_base + i] : R(abs2(a[row_base + i])));
        else acc = acc + abs2(a[row_base + i]);
    }
    return warp_butterfly(acc);
}

// state of the generated solver's outer loop, resident on the device (one per cycle)
struct SolveState {
    double res0, res_prev, res;
    double sum;       // last reduced sum of squares
    int it;           // iterations done
    int done;         // 1 = converged / cap / non-finite
    int bad;          // non-finite residual seen
    int cap;          // > 0: stop after this many iterations (solo re-timing of a batch member); 0 = no extra cap
    unsigned long long t_start;     // %globaltimer at the initial residual
    unsigned long long timeout_ns;  // > 0: watchdog, checked after every iteration
    unsigned long long t_it1, t_last;   // %globaltimer after the first / the latest iteration (per-iteration time of a capped run)
    int timed_out;
    int pad;
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// |r|^2 row sums of every field: one warp per row; rows[(field * nzi + (z - z0)) * ni + (y - 1)]
template <typename T, int DIM, int NF>
__global__ void __launch_bounds__(128) k_row_sumsq(const Geom g, Fields<T> r, double *rows)
{
    const int ni = g.n - 2, nzi = DIM == 3 ? ni : 1;
    const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= (long long)ni * nzi) return;
    const int y = 1 + (int)(row % ni), z = DIM == 3 ? 1 + (int)(row / ni) : 0;
    const long long base = node_index(g, 1, y, z);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        double s = warp_row_dot<T, double>(r.p[i], nullptr, base, ni);
        if ((threadIdx.x & 31) == 0) rows[(long long)i * ni * nzi + row] = s;
    }
}

// canonical reduction, second and third level: row sums -> plane sums (3-D: one warp per plane, many
// blocks) -> field sums -> total (one warp)
static __global__ void __launch_bounds__(256) k_reduce_planes(const double *rows, int count /* fields * planes */, int ni, double *planes)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long id = (long long)blockIdx.x * 8 + warp;   // (field, plane)
    if (id >= count) return;
    double s = warp_vecsum(rows + id * ni, ni);
    if (lane == 0) planes[id] = s;
}

// vals: per field `m` partial sums (plane sums in 3-D, row sums in 2-D)
static __global__ void __launch_bounds__(32) k_reduce_final(const double *vals, int nf, int m, SolveState *st)
{
    double total = 0.0;
    for (int i = 0; i < nf; ++i) total = total + warp_vecsum(vals + (long long)i * m, m);
    if (threadIdx.x == 0) st->sum = total;
}

// bookkeeping of the outer loop: `until res < tol*res0 or it >= maxIts`
// (example_problems/Poisson/2D_FD_Poisson_fromL2.exa3:3-4); mode 0 = initial residual, 1 = after a cycle
__device__ __forceinline__ void outer_update(SolveState *st, double *hist, double tol, int max_iters, int mode)
{
    const double res = sqrt(st->sum);
    if (mode == 0) {
        st->res0 = res; st->res_prev = res; st->res = res; st->it = 0; st->bad = 0; st->timed_out = 0;
        st->t_start = global_timer_ns();
        hist[0] = res;
        st->done = (max_iters <= 0) ? 1 : 0;
        if (!isfinite(res)) { st->bad = 1; st->done = 1; }
        return;
    }
    if (st->done) return;
    st->it += 1;
    st->t_last = global_timer_ns();
    if (st->it == 1) st->t_it1 = st->t_last;
    hist[st->it] = res;
    st->res_prev = st->res;
    st->res = res;
    if (!isfinite(res)) { st->bad = 1; st->done = 1; }
    else if (res < tol * st->res0 || st->it >= max_iters || (st->cap > 0 && st->it >= st->cap)) st->done = 1;
    else if (st->timeout_ns && global_timer_ns() - st->t_start > st->timeout_ns) { st->timed_out = 1; st->done = 1; }
}
static __global__ void k_outer_update(SolveState *st, double *hist, double tol, int max_iters, int mode)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    outer_update(st, hist, tol, max_iters, mode);
}
// k_reduce_final + k_outer_update + the conditional handles of the solver graph's loop in one launch
static __global__ void __launch_bounds__(32) k_reduce_final_update(const double *vals, int nf, int m, SolveState *st, double *hist,
                                                                   double tol, int max_iters, int mode, int n_handles,
                                                                   cudaGraphConditionalHandle h0, cudaGraphConditionalHandle h1)
{
    double total = 0.0;
    for (int i = 0; i < nf; ++i) total = total + warp_vecsum(vals + (long long)i * m, m);
    if (threadIdx.x != 0) return;
    st->sum = total;
    outer_update(st, hist, tol, max_iters, mode);
    const unsigned v = st->done ? 0u : 1u;
    if (n_handles >= 1) cudaGraphSetConditional(h0, v);
    if (n_handles >= 2) cudaGraphSetConditional(h1, v);
}

// residual and canonical row sums of |r|^2 in one pass (generic stencil tables; one warp per row, the lanes walk the
// row in the canonical order); STORE = also write the residual field
template <typename T, int DIM, int NF, bool STORE>
__global__ void __launch_bounds__(128) k_residual_rowsum(const Geom g, const __grid_constant__ OpSten st, Fields<T> u, Fields<T> f,
                                                         Fields<T> r, double *rows)
{
    const int ni = g.n - 2, nzi = DIM == 3 ? ni : 1, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= (long long)ni * nzi) return;
    const int y = 1 + (int)(row % ni), z = DIM == 3 ? 1 + (int)(row / ni) : 0;
    const long long base = node_index(g, 1, y, z);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        double acc = 0.0;
        for (int x = lane; x < ni; x += 32) {
            const long long idx = base + x;
            const T v = f.p[i][idx] - apply_row<T, NF>(g, st, u, i, idx);
            if (STORE) r.p[i][idx] = v;
            acc = acc + abs2(v);
        }
        acc = warp_butterfly(acc);
        if (lane == 0) rows[(long long)i * ni * nzi + row] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// One `solve locally` statement (exastencils.py:769-822; ir/transformations.py:51-121).
struct SmoothParams {
    int nu;                       // unknowns of the local system
    int field[EVO_MAX_UNKNOWNS];  // field of each unknown
    int off[EVO_MAX_UNKNOWNS][3]; // node offset of each unknown relative to the anchor
    int color;                    // -1: every inner node is an anchor; 0/1: anchors with (x+y+z)%2 == color
    int write_all;                // 1: write every inner unknown (in-place modes); 0: only offset-0 unknowns
    double omega;
};

// LU factorisation with partial pivoting, reciprocal pivots (one division per pivot); the factorisation
// depends only on M, so callers with a node-independent matrix hoist it out of their loops
template <typename T, int NU> struct DenseLU {
    T lu[NU][NU];   // unit-lower factors below the diagonal, U on and above
    T inv[NU];      // reciprocal pivots
    int piv[NU];    // row exchanged with row c at step c
};

template <typename T, int NU>
__device__ __forceinline__ void dense_factor(T (&M)[NU][NU], DenseLU<T, NU> &f)
{
#pragma unroll
    for (int c = 0; c < NU; ++c) {
        int piv = c;
        double best = abs2(M[c][c]);
#pragma unroll
        for (int r = c + 1; r < NU; ++r) {
            double v = abs2(M[r][c]);
            if (v > best) { best = v; piv = r; }
        }
        f.piv[c] = piv;
        if (piv != c) {
#pragma unroll
            for (int r = 0; r < NU; ++r) {
                if (r == piv) {
#pragma unroll
                    for (int k = 0; k < NU; ++k) { T t = M[c][k]; M[c][k] = M[r][k]; M[r][k] = t; }
                }
            }
        }
        f.inv[c] = T(1.0) / M[c][c];
#pragma unroll
        for (int r = c + 1; r < NU; ++r) {
            T fac = M[r][c] * f.inv[c];
            M[r][c] = fac;
#pragma unroll
            for (int k = c + 1; k < NU; ++k) M[r][k] = M[r][k] - fac * M[c][k];
        }
    }
#pragma unroll
    for (int r = 0; r < NU; ++r)
#pragma unroll
        for (int k = 0; k < NU; ++k) f.lu[r][k] = M[r][k];
}

template <typename T, int NU>
__device__ __forceinline__ void dense_solve(const DenseLU<T, NU> &f, T (&b)[NU])
{
    // P b first (the stored factors are in their final row order), then forward substitution: element for
    // element the same operations as an elimination that carries b along
#pragma unroll
    for (int c = 0; c < NU; ++c) {
        if (f.piv[c] != c) {
#pragma unroll
            for (int r = 0; r < NU; ++r)
                if (r == f.piv[c]) { T t = b[c]; b[c] = b[r]; b[r] = t; }
        }
    }
#pragma unroll
    for (int c = 0; c < NU; ++c) {
#pragma unroll
        for (int r = c + 1; r < NU; ++r) b[r] = b[r] - f.lu[r][c] * b[c];
    }
#pragma unroll
    for (int c = NU - 1; c >= 0; --c) {
        T sacc = b[c];
#pragma unroll
        for (int k = c + 1; k < NU; ++k) sacc = sacc - f.lu[c][k] * b[k];
        b[c] = sacc * f.inv[c];
    }
}

template <typename T, int NU>
__device__ __forceinline__ void solve_dense(T (&M)[NU][NU], T (&b)[NU])
{
    DenseLU<T, NU> f;
    dense_factor<T, NU>(M, f);
    dense_solve<T, NU>(f, b);
}

// local solve at one anchor; identical construction to oracle/mg_ops.inc local_solve_anchor
template <typename T, int DIM, int NF, int NU>
__device__ __forceinline__ void local_solve(const Geom &g, const OpSten &st, const SmoothParams &sp,
                                            const Fields<T> &src, const Fields<T> &dst, const Fields<T> &rhs,
                                            int ax, int ay, int az)
{
    T M[NU][NU];
    T b[NU];
    int ux[NU], uy[NU], uz[NU];
    bool inner[NU];
    long long uidx[NU];
    const int n = g.n;
#pragma unroll
    for (int a = 0; a < NU; ++a) {
        ux[a] = ax + sp.off[a][0];
        uy[a] = ay + sp.off[a][1];
        uz[a] = DIM == 3 ? az + sp.off[a][2] : 0;
        inner[a] = ux[a] >= 1 && ux[a] <= n - 2 && uy[a] >= 1 && uy[a] <= n - 2 &&
                   (DIM == 2 || (uz[a] >= 1 && uz[a] <= n - 2));
        bool inside = ux[a] >= 0 && ux[a] <= n - 1 && uy[a] >= 0 && uy[a] <= n - 1 &&
                      (DIM == 2 || (uz[a] >= 0 && uz[a] <= n - 1));
        uidx[a] = inside ? node_index(g, ux[a], uy[a], uz[a]) : -1;
    }
#pragma unroll
    for (int a = 0; a < NU; ++a) {
#pragma unroll
        for (int m = 0; m < NU; ++m) M[a][m] = T(0.0);
        if (!inner[a]) {
            M[a][a] = T(1.0);
            b[a] = uidx[a] >= 0 ? src.p[sp.field[a]][uidx[a]] : T(0.0);
            continue;
        }
        const int fi = sp.field[a];
        T s = T(0.0);
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            const Sten &sj = st.s[fi][j];
            for (int q = 0; q < sj.nnz; ++q) {
                const int px = ux[a] + sj.ox[q], py = uy[a] + sj.oy[q], pz = uz[a] + sj.oz[q];
                int hit = -1;
#pragma unroll
                for (int m = NU - 1; m >= 0; --m)
                    if (sp.field[m] == j && ux[m] == px && uy[m] == py && uz[m] == pz) hit = m;
                const T c = sten_coef<T>(sj, q);
                if (hit >= 0) {
#pragma unroll
                    for (int m = 0; m < NU; ++m)
                        if (m == hit) M[a][m] = M[a][m] + c;
                } else {
                    long long d = (long long)sj.oz[q] * g.plane + (long long)sj.oy[q] * g.pitch + sj.ox[q];
                    s = s + c * src.p[j][uidx[a] + d];
                }
            }
        }
        b[a] = rhs.p[fi][uidx[a]] - s;
    }
    if (NU == 1) b[0] = b[0] * (T(1.0) / M[0][0]);  // pointwise: multiply by the reciprocal diagonal
    else solve_dense<T, NU>(M, b);
#pragma unroll
    for (int a = 0; a < NU; ++a) {
        if (!inner[a]) continue;
        const bool own = sp.off[a][0] == 0 && sp.off[a][1] == 0 && sp.off[a][2] == 0;
        if (!sp.write_all && !own) continue;
        const T old = src.p[sp.field[a]][uidx[a]];
        dst.p[sp.field[a]][uidx[a]] = old + sp.omega * (b[a] - old);
    }
}

// parallel sweep: `with jacobi` (src = current slot, dst = next slot; anchors at every inner node in
// lexicographic order with overlapping blocks == every node keeps the value its own anchor computes,
// because its own anchor is the last writer) or one colour of an order-independent coloured sweep.
template <typename T, int DIM, int NF, int NU>
__global__ void __launch_bounds__(BX) k_smooth(const Geom g, const __grid_constant__ OpSten st,
                                               const __grid_constant__ SmoothParams sp, Fields<T> src, Fields<T> dst,
                                               Fields<T> rhs)
{
    const int y = 1 + blockIdx.y, z = DIM == 3 ? 1 + blockIdx.z : 0;
    const int t = blockIdx.x * BX + threadIdx.x;
    int x;
    if (sp.color < 0) x = 1 + t;
    else x = 1 + 2 * t + ((1 + y + z + sp.color) & 1);
    if (x > g.n - 2) return;
    local_solve<T, DIM, NF, NU>(g, st, sp, src, dst, rhs, x, y, z);
}

// order-independent coloured sweeps on a tiny grid (<= 4096 inner nodes): ONE CTA runs all `sweeps` repetitions and both
// colours, a block barrier between the colours -- 2 x sweeps kernel nodes become one.  Same local_solve per anchor as
// k_smooth, anchors of one colour are independent -> bit-identical.
template <typename T, int DIM, int NF, int NU>
__global__ void __launch_bounds__(1024) k_smooth_rb_small(const Geom g, const __grid_constant__ OpSten st,
                                                          const __grid_constant__ SmoothParams sp, Fields<T> u, Fields<T> rhs,
                                                          const int sweeps)
{
    SmoothParams loc = sp;
    const int ni = g.n - 2, nzi = DIM == 3 ? ni : 1;
    const int half = (ni + 1) / 2;                  // anchors of one colour per row (upper bound)
    const int total = half * ni * nzi;
    for (int sw = 0; sw < sweeps; ++sw)
        for (int color = 0; color < 2; ++color) {
            loc.color = color;
            for (int t = threadIdx.x; t < total; t += 1024) {
                const int tx = t % half, row = t / half;
                const int y = 1 + row % ni, z = DIM == 3 ? 1 + row / ni : 0;
                const int x = 1 + 2 * tx + ((1 + y + z + color) & 1);
                if (x <= g.n - 2) local_solve<T, DIM, NF, NU>(g, st, loc, u, u, rhs, x, y, z);
            }
            __threadfence_block();
            __syncthreads();
        }
}

// order-dependent coloured sweep (e.g. collective RB-GS on the elasticity system, whose dxy corner
// terms couple same-colour nodes): the reference's loop is sequential, i0 fastest; anchors of one row
// are mutually independent for 3^d stencils, rows depend on the previous row -> one CTA walks the rows
// in lexicographic order, all threads of a row in parallel.
template <typename T, int DIM, int NF, int NU>
__global__ void __launch_bounds__(1024) k_smooth_rowseq(const Geom g, const __grid_constant__ OpSten st,
                                                        const __grid_constant__ SmoothParams sp, Fields<T> u,
                                                        Fields<T> rhs)
{
    SmoothParams loc = sp;
    const int z0 = DIM == 3 ? 1 : 0, z1 = DIM == 3 ? g.n - 1 : 1;
    for (int color = 0; color < 2; ++color) {
        loc.color = color;
        for (int z = z0; z < z1; ++z)
            for (int y = 1; y < g.n - 1; ++y) {
                for (int t = threadIdx.x;; t += 1024) {
                    int x = 1 + 2 * t + ((1 + y + z + color) & 1);
                    if (x > g.n - 2) break;
                    local_solve<T, DIM, NF, NU>(g, st, loc, u, u, rhs, x, y, z);
                }
                __threadfence_block();
                __syncthreads();
            }
    }
}

// Lexicographic in-place sweep (`solve locally` without `with jacobi` and without colouring: what the reference
// emits in model-based mode, exastencils.py:64-70, :781-822 with _use_jacobi_prefix = False).  The reference loop
// is sequential, i0 fastest; anchors P < Q conflict when one writes what the other reads or writes.  With the skew
// t = x + a*y + b*z (a, b from lex_skew(): every conflicting pair has t(P) < t(Q)) all anchors of one hyperplane t
// are independent, and sweeping the hyperplanes in ascending t reproduces the sequential result exactly (block
// smoothers with overlapping writes included).  One thread-block cluster runs the whole sweep: CTAs share a
// hyperplane, a cluster barrier separates hyperplanes.  Latency bound by construction (as the statement is).
__device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }   // b > 0
__device__ __forceinline__ int ceil_div(int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }  // b > 0

__device__ __forceinline__ void cluster_barrier()
{
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

template <typename T, int DIM, int NF, int NU>
__global__ void __launch_bounds__(1024) k_smooth_lex(const Geom g, const __grid_constant__ OpSten st,
                                                     const __grid_constant__ SmoothParams sp, Fields<T> u, Fields<T> rhs,
                                                     const int a, const int b, const int sweeps)
{
    const int ni = g.n - 2;
    const int nthreads = (int)(cluster_nctarank() * blockDim.x), gtid = (int)(cluster_ctarank() * blockDim.x + threadIdx.x);
    const int nwarps = nthreads >> 5, gw = gtid >> 5, lane = gtid & 31;
    const int tmin = 1 + a + (DIM == 3 ? b : 0), tmax = ni * (1 + a + (DIM == 3 ? b : 0));
    for (int sweep = 0; sweep < sweeps; ++sweep)
        for (int t = tmin; t <= tmax; ++t) {
            if (DIM == 2) {
                // x = t - a*y in [1, ni]
                int ylo = 1, yhi = ni;
                if (a > 0) { ylo = max(1, ceil_div(t - ni, a)); yhi = min(ni, floor_div(t - 1, a)); }
                else if (t > ni) { ylo = 1; yhi = 0; }
                for (int y = ylo + gtid; y <= yhi; y += nthreads) local_solve<T, DIM, NF, NU>(g, st, sp, u, u, rhs, t - a * y, y, 0);
            } else {
                // x + a*y = t - b*z in [1 + a, ni*(1 + a)]
                int zlo = 1, zhi = ni;
                if (b > 0) { zlo = max(1, ceil_div(t - ni * (1 + a), b)); zhi = min(ni, floor_div(t - 1 - a, b)); }
                else if (t > ni * (1 + a)) { zlo = 1; zhi = 0; }
                for (int z = zlo + gw; z <= zhi; z += nwarps) {
                    const int rem = t - b * z;
                    int ylo = 1, yhi = ni;
                    if (a > 0) { ylo = max(1, ceil_div(rem - ni, a)); yhi = min(ni, floor_div(rem - 1, a)); }
                    else if (rem < 1 || rem > ni) { ylo = 1; yhi = 0; }
                    for (int y = ylo + lane; y <= yhi; y += 32) local_solve<T, DIM, NF, NU>(g, st, sp, u, u, rhs, rem - a * y, y, z);
                }
            }
            cluster_barrier();
        }
}

// ------------------------------------------------------------------------------------------------
struct TransferW {  // non-zero transfer weights in ascending table order
    int nnz;
    signed char ox[27], oy[27], oz[27];
    double w[27];
};

// dst@(l-1) = R@l * src@l : coarse inner node c reads fine node 2c + o   (exastencils.py:855-873, :1126-1147)
template <typename T, int DIM, int NF>
__global__ void __launch_bounds__(BX) k_restrict(const Geom gf, const Geom gc, const __grid_constant__ TransferW R,
                                                 Fields<T> src, Fields<T> dst)
{
    // z-slab aware: local coarse plane z holds global plane z + gc.zoff, whose fine plane is 2*(z + gc.zoff)
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y, z = DIM == 3 ? gc.zlo + blockIdx.z : 0;
    if (x > gc.n - 2) return;
    const long long fidx = node_index(gf, 2 * x, 2 * y, DIM == 3 ? 2 * (z + gc.zoff) - gf.zoff : 0);
    const long long cidx = node_index(gc, x, y, z);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        T acc = T(0.0);
        for (int q = 0; q < R.nnz; ++q) {
            long long d = (long long)R.oz[q] * gf.plane + (long long)R.oy[q] * gf.pitch + R.ox[q];
            acc = acc + R.w[q] * src.p[i][fidx + d];
        }
        dst.p[i][cidx] = acc;
    }
}

// fused RHS@(l-1) = R@l * (RHS@l - A@l * SOL@l): the fine residual is never stored
template <typename T, int DIM, int NF>
__global__ void __launch_bounds__(BX) k_residual_restrict(const Geom gf, const Geom gc,
                                                          const __grid_constant__ OpSten st,
                                                          const __grid_constant__ TransferW R, Fields<T> u, Fields<T> f,
                                                          Fields<T> dst, Fields<T> zero /* coarse fields to clear, or null */)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y, z = DIM == 3 ? 1 + blockIdx.z : 0;
    if (x > gc.n - 2) return;
    const long long cidx = node_index(gc, x, y, z);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        T acc = T(0.0);
        for (int q = 0; q < R.nnz; ++q) {
            const int fx = 2 * x + R.ox[q], fy = 2 * y + R.oy[q], fz = DIM == 3 ? 2 * z + R.oz[q] : 0;
            T rv = T(0.0);  // the residual field is 0 on the boundary layer
            if (fx >= 1 && fx <= gf.n - 2 && fy >= 1 && fy <= gf.n - 2 && (DIM == 2 || (fz >= 1 && fz <= gf.n - 2))) {
                const long long idx = node_index(gf, fx, fy, fz);
                rv = f.p[i][idx] - apply_row<T, NF>(gf, st, u, i, idx);
            }
            acc = acc + R.w[q] * rv;
        }
        dst.p[i][cidx] = acc;
        if (zero.p[i]) zero.p[i][cidx] = T(0.0);
    }
}

// the same statement on a tiny coarse grid: one WARP per coarse node, lane q evaluates the fine residual of restriction
// entry q, lane 0 adds the R.nnz products in ascending q like the loop above (bit-identical).  The thread-per-node kernel
// spends 20+ us on a 7^3 grid (27 dependent stencil evaluations per thread); this one ~4 us.
template <typename T, int DIM, int NF>
__global__ void __launch_bounds__(128) k_residual_restrict_warp(const Geom gf, const Geom gc, const __grid_constant__ OpSten st,
                                                                const __grid_constant__ TransferW R, Fields<T> u, Fields<T> f,
                                                                Fields<T> dst, Fields<T> zero)
{
    const int nci = gc.n - 2, lane = threadIdx.x & 31;
    const long long node = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const long long count = (long long)nci * nci * (DIM == 3 ? nci : 1);
    if (node >= count) return;
    const int x = 1 + (int)(node % nci), y = 1 + (int)((node / nci) % nci), z = DIM == 3 ? 1 + (int)(node / ((long long)nci * nci)) : 0;
    const long long cidx = node_index(gc, x, y, z);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        T term = T(0.0);
        if (lane < R.nnz) {
            const int fx = 2 * x + R.ox[lane], fy = 2 * y + R.oy[lane], fz = DIM == 3 ? 2 * z + R.oz[lane] : 0;
            T rv = T(0.0);
            if (fx >= 1 && fx <= gf.n - 2 && fy >= 1 && fy <= gf.n - 2 && (DIM == 2 || (fz >= 1 && fz <= gf.n - 2))) {
                const long long idx = node_index(gf, fx, fy, fz);
                rv = f.p[i][idx] - apply_row<T, NF>(gf, st, u, i, idx);
            }
            term = R.w[lane] * rv;
        }
        T acc = T(0.0);
        for (int q = 0; q < R.nnz; ++q) acc = acc + shfl_idx(term, q);
        if (lane == 0) {
            dst.p[i][cidx] = acc;
            if (zero.p[i]) zero.p[i][cidx] = T(0.0);
        }
    }
}

// fine inner node x: p = sum over offsets o with x+o even of w[o]*src[(x+o)/2]   (exastencils.py:1103-1124)
// ADD: SOL@l += w * p (:727-743)   else: dst@l = p (:868-873)
template <typename T, int DIM, int NF, bool ADD>
__global__ void __launch_bounds__(BX) k_prolong(const Geom gf, const Geom gc, const __grid_constant__ TransferW P,
                                                Fields<T> src, Fields<T> dst, double weight)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y, z = DIM == 3 ? 1 + blockIdx.z : 0;
    if (x > gf.n - 2) return;
    const long long idx = node_index(gf, x, y, z);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        T acc = T(0.0);
        for (int q = 0; q < P.nnz; ++q) {
            const int cx = x + P.ox[q], cy = y + P.oy[q], cz = DIM == 3 ? z + P.oz[q] : 0;
            if ((cx & 1) || (cy & 1) || (DIM == 3 && (cz & 1))) continue;
            acc = acc + P.w[q] * src.p[i][node_index(gc, cx >> 1, cy >> 1, cz >> 1)];
        }
        if (ADD) dst.p[i][idx] = dst.p[i][idx] + weight * acc;
        else dst.p[i][idx] = acc;
    }
}

// SOL_i += w * (RHS_i - A SOL)_i evaluated from the pre-statement values (see oracle op_richardson):
// two kernels, k_residual into a scratch field then this axpy
template <typename T, int DIM>
__global__ void __launch_bounds__(BX) k_axpy_inner(const Geom g, T *y_, const T *x_, double a)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y, z = DIM == 3 ? 1 + blockIdx.z : 0;
    if (x > g.n - 2) return;
    const long long idx = node_index(g, x, y, z);
    y_[idx] = y_[idx] + a * x_[idx];
}

// ------------------------------------------------------------------------------------------------
// gen_mgCycle@min(): conjugate gradients on the coarsest level, zero initial guess, one CTA
// (`solver_cgs = "CG"`, example_problems/Poisson/2D_FD_Poisson_fromL2.exa3:12-14).  The whole
// coarse problem is latency bound, so a single block with block-wide barriers beats any multi-kernel
// formulation; dot products use the deterministic block reduction.
template <int DIM, int NF>
__global__ void __launch_bounds__(1024) k_coarse_cg(const Geom g, const __grid_constant__ OpSten st, Fields<double> x,
                                                    Fields<double> b, Fields<double> r, Fields<double> p,
                                                    Fields<double> ap, double *rows, int max_it, double tol,
                                                    int *iters_out)
{
    __shared__ double planes[1024];
    __shared__ double bc;
    const int ni = g.n - 2, nzi = DIM == 3 ? ni : 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long count = (long long)ni * ni * (DIM == 3 ? ni : 1);
    auto idx_of = [&](long long t) {
        int ix = (int)(t % ni), iy = (int)((t / ni) % ni), iz = DIM == 3 ? (int)(t / ((long long)ni * ni)) : -1;
        return node_index(g, ix + 1, iy + 1, iz + 1);
    };
    // canonical dot product (see warp_vecsum): rows by warps, then planes, then fields in order
    auto dot = [&](const Fields<double> &a_, const Fields<double> &b_) {
        double total = 0.0;
        for (int i = 0; i < NF; ++i) {
            for (int row = warp; row < ni * nzi; row += 32) {
                const int y = 1 + row % ni, z = DIM == 3 ? 1 + row / ni : 0;
                double s = warp_row_dot<double, double>(a_.p[i], b_.p[i], node_index(g, 1, y, z), ni);
                if (lane == 0) rows[row] = s;
            }
            __syncthreads();
            if (DIM == 3) {
                for (int zz = warp; zz < nzi; zz += 32) {
                    double s = warp_vecsum(rows + (long long)zz * ni, ni);
                    if (lane == 0) planes[zz] = s;
                }
                __syncthreads();
            }
            if (warp == 0) {
                double s = DIM == 3 ? warp_vecsum(planes, nzi) : warp_vecsum(rows, ni);
                if (lane == 0) bc = s;
            }
            __syncthreads();
            total = total + bc;
            __syncthreads();
        }
        return total;
    };
    for (int i = 0; i < NF; ++i)
        for (long long t = threadIdx.x; t < g.total; t += blockDim.x) {
            x.p[i][t] = 0.0; r.p[i][t] = 0.0; p.p[i][t] = 0.0; ap.p[i][t] = 0.0;
        }
    __syncthreads();
    for (int i = 0; i < NF; ++i)
        for (long long t = threadIdx.x; t < count; t += blockDim.x) {
            long long id = idx_of(t);
            double v = b.p[i][id];
            r.p[i][id] = v; p.p[i][id] = v;
        }
    __syncthreads();
    double rr = dot(r, r);
    const double r0 = sqrt(rr);
    int it = 0;
    if (r0 != 0.0) {
        while (it < max_it) {
            for (int i = 0; i < NF; ++i)
                for (long long t = threadIdx.x; t < count; t += blockDim.x) {
                    long long id = idx_of(t);
                    ap.p[i][id] = apply_row<double, NF>(g, st, p, i, id);
                }
            __syncthreads();
            const double pap = dot(p, ap);
            const double alpha = rr / pap;
            for (int i = 0; i < NF; ++i)
                for (long long t = threadIdx.x; t < count; t += blockDim.x) {
                    long long id = idx_of(t);
                    x.p[i][id] = x.p[i][id] + alpha * p.p[i][id];
                    r.p[i][id] = r.p[i][id] - alpha * ap.p[i][id];
                }
            __syncthreads();
            const double rr_new = dot(r, r);
            ++it;
            if (!(sqrt(rr_new) > tol * r0)) break;
            const double beta = rr_new / rr;
            for (int i = 0; i < NF; ++i)
                for (long long t = threadIdx.x; t < count; t += blockDim.x) {
                    long long id = idx_of(t);
                    p.p[i][id] = r.p[i][id] + beta * p.p[i][id];
                }
            __syncthreads();
            rr = rr_new;
        }
    }
    if (threadIdx.x == 0 && iters_out) *iters_out = it;
}


// ------------------------------------------------------------------------------------------------
// Shared-memory resident CG for coarsest levels that fit (4 vectors x fields x n^d doubles): the
// global-memory version above spends ~7 us per iteration on dependent global round trips; here a CG
// iteration is 5 block barriers.  Same arithmetic and the same canonical reductions -> bit-identical.
// Device function: the whole CTA (any multiple of 32 threads) calls it -- from k_coarse_cg_smem and from the fused-run
// interpreter (evo_kernels_run.cuh).  (A 4-warp launch for the 5^3 grid was measured SLOWER than 32 warps: the table-driven
// A p of a row, not the block barriers, is the latency of an iteration.)
template <int DIM, int NF>
__device__ __forceinline__ void coarse_cg_smem(const Geom &g, const OpSten &st, Fields<double> xg, Fields<double> bg, int max_it,
                                               double tol, int *iters_out, double *sm)
{
    const int NT = (int)blockDim.x;
    __shared__ double rowsum[1024];
    __shared__ double planes[64];
    const int n = g.n, ni = n - 2, nzi = DIM == 3 ? ni : 1;
    const int vol = n * n * (DIM == 3 ? n : 1);          // compact (unpadded) node count incl. boundary
    const int nrows = ni * nzi;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *x = sm, *r = sm + (size_t)NF * vol, *p = sm + 2 * (size_t)NF * vol, *ap = sm + 3 * (size_t)NF * vol;
    // compact index of node (ix, iy, iz): (iz*n + iy)*n + ix
    for (int t = threadIdx.x; t < 4 * NF * vol; t += NT) sm[t] = 0.0;
    __syncthreads();
    for (int i = 0; i < NF; ++i)
        for (int row = warp; row < nrows; row += NT / 32) {
            const int y = 1 + row % ni, z = DIM == 3 ? 1 + row / ni : 0;
            for (int xx = 1 + lane; xx <= ni; xx += 32) {
                const double v = bg.p[i][node_index(g, xx, y, z)];
                const int c = i * vol + (z * n + y) * n + xx;
                r[c] = v; p[c] = v;
            }
        }
    __syncthreads();
    // second and third level of the canonical reduction over rowsum[field][row].  2-D: after ONE barrier
    // every warp recomputes the (identical) final vecsum itself -- no broadcast; the row-sum buffer is
    // double buffered (rs_flip) so the next pass may overwrite the other half without a second barrier.
    int rs_flip = 0;
    auto finish = [&]() {
        double total = 0.0;
        __syncthreads();
        const double *rs = rowsum + rs_flip * 512;
        for (int i = 0; i < NF; ++i) {
            if (DIM == 3) {
                for (int zz = warp; zz < nzi; zz += NT / 32) {
                    double sacc = warp_vecsum(rs + i * nrows + zz * ni, ni);
                    if (lane == 0) planes[zz] = sacc;
                }
                __syncthreads();
                total = total + warp_vecsum(planes, nzi);
                __syncthreads();
            } else {
                total = total + warp_vecsum(rs + i * nrows, ni);
            }
        }
        rs_flip ^= 1;
        return total;
    };
    // r.r
    for (int i = 0; i < NF; ++i)
        for (int row = warp; row < nrows; row += NT / 32) {
            const int y = 1 + row % ni, z = DIM == 3 ? 1 + row / ni : 0;
            double acc = 0.0;
            for (int xx = 1 + lane; xx <= ni; xx += 32) {
                const int c = i * vol + (z * n + y) * n + xx;
                acc = acc + r[c] * r[c];
            }
            acc = warp_butterfly(acc);
            if (lane == 0) rowsum[rs_flip * 512 + i * nrows + row] = acc;
        }
    double rr = finish();
    const double r0 = sqrt(rr);
    int it = 0;
    if (r0 != 0.0) {
        while (it < max_it) {
            // pass A: ap = A p and the row sums of p.ap
            for (int i = 0; i < NF; ++i)
                for (int row = warp; row < nrows; row += NT / 32) {
                    const int y = 1 + row % ni, z = DIM == 3 ? 1 + row / ni : 0;
                    double acc = 0.0;
                    for (int xx = 1 + lane; xx <= ni; xx += 32) {
                        const int c0 = (z * n + y) * n + xx;
                        double a = 0.0;
#pragma unroll
                        for (int j = 0; j < NF; ++j) {
                            const Sten &sj = st.s[i][j];
                            for (int q = 0; q < sj.nnz; ++q)
                                a = a + sj.re[q] * p[j * vol + c0 + (sj.oz[q] * n + sj.oy[q]) * n + sj.ox[q]];
                        }
                        ap[i * vol + c0] = a;
                        acc = acc + p[i * vol + c0] * a;
                    }
                    acc = warp_butterfly(acc);
                    if (lane == 0) rowsum[rs_flip * 512 + i * nrows + row] = acc;
                }
            const double pap = finish();
            const double alpha = rr / pap;
            // pass B: x += alpha p, r -= alpha ap and the row sums of r.r
            for (int i = 0; i < NF; ++i)
                for (int row = warp; row < nrows; row += NT / 32) {
                    const int y = 1 + row % ni, z = DIM == 3 ? 1 + row / ni : 0;
                    double acc = 0.0;
                    for (int xx = 1 + lane; xx <= ni; xx += 32) {
                        const int c = i * vol + (z * n + y) * n + xx;
                        x[c] = x[c] + alpha * p[c];
                        const double rv = r[c] - alpha * ap[c];
                        r[c] = rv;
                        acc = acc + rv * rv;
                    }
                    acc = warp_butterfly(acc);
                    if (lane == 0) rowsum[rs_flip * 512 + i * nrows + row] = acc;
                }
            const double rr_new = finish();
            ++it;
            if (!(sqrt(rr_new) > tol * r0)) break;
            const double beta = rr_new / rr;
            for (int t = threadIdx.x; t < NF * vol; t += NT) p[t] = r[t] + beta * p[t];   // boundary entries stay 0
            rr = rr_new;
            __syncthreads();
        }
    }
    __syncthreads();
    // x -> SOL (inner nodes; the boundary layer of SOL is 0 on the coarsest level and stays so)
    for (int i = 0; i < NF; ++i)
        for (int row = warp; row < nrows; row += NT / 32) {
            const int y = 1 + row % ni, z = DIM == 3 ? 1 + row / ni : 0;
            for (int xx = 1 + lane; xx <= ni; xx += 32) xg.p[i][node_index(g, xx, y, z)] = x[i * vol + (z * n + y) * n + xx];
        }
    if (threadIdx.x == 0 && iters_out) *iters_out = it;
}

template <int DIM, int NF>
__global__ void __launch_bounds__(1024) k_coarse_cg_smem(const Geom g, const __grid_constant__ OpSten st, Fields<double> xg,
                                                         Fields<double> bg, int max_it, double tol, int *iters_out)
{
    extern __shared__ double cg_sm[];
    coarse_cg_smem<DIM, NF>(g, st, xg, bg, max_it, tol, iters_out, cg_sm);
}

// ------------------------------------------------------------------------------------------------
// Order-dependent coloured pointwise sweep of a 2-D system (collective RB-GS on the elasticity system),
// rolling shared-memory window: the sequential chain is one block barrier per row instead of a global
// memory round trip per row.  Rows j-1 (already updated), j and j+1 (old values) are in shared memory,
// row j+3 is fetched by cp.async two steps ahead; results go to shared memory (for the next row) and to
// global memory.  Unknown a == field a at the anchor node (NU == NF); arithmetic identical to
// local_solve (same (j, q) order, same dense solve).
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NF>
__global__ void __launch_bounds__(1024) k2_smooth_rowseq_win(const Geom g, const __grid_constant__ OpSten st, const double omega,
                                                             Fields<double> u, Fields<double> rhs)
{
    extern __shared__ __align__(16) double win[];   // [5 slots][2*NF arrays][pitch]
    const int n = g.n, pitch = g.pitch;
    const int chunks = pitch / 2;                    // 16-byte chunks per row
    auto slot_u = [&](int row, int f) { return win + ((size_t)(row % 5) * 2 * NF + f) * pitch; };
    auto slot_f = [&](int row, int f) { return win + ((size_t)(row % 5) * 2 * NF + NF + f) * pitch; };
    auto fetch = [&](int row) {
        if (row >= 0 && row <= n - 1)
            for (int t = threadIdx.x; t < 2 * NF * chunks; t += 1024) {
                const int a = t / chunks, c = t - a * chunks;
                const double *src = (a < NF ? u.p[a] : rhs.p[a - NF]) + (long long)row * pitch + 2 * c;
                cp_async16(win + ((size_t)(row % 5) * 2 * NF + a) * pitch + 2 * c, src);
            }
        cp_async_commit();
    };
    // the local matrix (diagonal entries of the blocks) is the same at every anchor: factor it once
    double M0[NF][NF];
#pragma unroll
    for (int a = 0; a < NF; ++a)
#pragma unroll
        for (int m = 0; m < NF; ++m) M0[a][m] = 0.0;
#pragma unroll
    for (int a = 0; a < NF; ++a)
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            const Sten &sj = st.s[a][j];
            for (int q = 0; q < sj.nnz; ++q)
                if (sj.ox[q] == 0 && sj.oy[q] == 0) {
#pragma unroll
                    for (int m = 0; m < NF; ++m)
                        if (m == j) M0[a][m] = M0[a][m] + sj.re[q];
                }
        }
    DenseLU<double, NF> lu;
    dense_factor<double, NF>(M0, lu);
    for (int color = 0; color < 2; ++color) {
        __threadfence();
        __syncthreads();
        fetch(0); fetch(1); fetch(2);
        for (int y = 1; y <= n - 2; ++y) {
            fetch(y + 2);
            cp_async_wait<1>();      // everything but the newest group (row y+2) has landed
            __syncthreads();         // ... for all threads; row y-1 results of the previous step are visible
            for (int t = threadIdx.x;; t += 1024) {
                const int x = 1 + 2 * t + ((1 + y + color) & 1);
                if (x > n - 2) break;
                double b[NF];
#pragma unroll
                for (int a = 0; a < NF; ++a) {
                    double sacc = 0.0;
#pragma unroll
                    for (int j = 0; j < NF; ++j) {
                        const Sten &sj = st.s[a][j];
                        for (int q = 0; q < sj.nnz; ++q)
                            if (!(sj.ox[q] == 0 && sj.oy[q] == 0)) sacc = sacc + sj.re[q] * slot_u(y + sj.oy[q], j)[x + sj.ox[q]];
                    }
                    b[a] = slot_f(y, a)[x] - sacc;
                }
                if (NF == 1) b[0] = b[0] * lu.inv[0];
                else dense_solve<double, NF>(lu, b);
#pragma unroll
                for (int a = 0; a < NF; ++a) {
                    const double old = slot_u(y, a)[x];
                    const double nv = old + omega * (b[a] - old);
                    slot_u(y, a)[x] = nv;
                    u.p[a][(long long)y * pitch + x] = nv;
                }
            }
        }
        cp_async_wait<0>();
    }
}

// dense 3x3 coefficient tables of the NF x NF blocks (table order p = (dy+1)*3 + (dx+1)); zero = no entry
template <int NF> struct Dense9 { double w[NF][NF][9]; };

// Sparsity pattern known at compile time (PAT = 1): 5-point star on the diagonal blocks, the four corners on the
// off-diagonal blocks -- the 2-D Poisson operator and the linear-elasticity system (dxx/dyy + the mixed dxy terms).  With the
// generic run-time masks (PAT = 0) the compiler turns "skip the absent entry" into a select, so every table entry costs a
// dependent fp64 add (~50 cycles each on the latency-bound kernels below); the static pattern halves those chains.
__host__ __device__ constexpr bool dense9_star_corner(int a, int j, int p)
{
    return a == j ? (p == 1 || p == 3 || p == 4 || p == 5 || p == 7) : (p == 0 || p == 2 || p == 6 || p == 8);
}
template <int NF> inline int dense9_pattern(const Dense9<NF> &dn)
{
    for (int a = 0; a < NF; ++a)
        for (int j = 0; j < NF; ++j)
            for (int p = 0; p < 9; ++p)
                if ((dn.w[a][j][p] != 0.0) != dense9_star_corner(a, j, p)) return 0;
    return 1;
}

// CG on a 2-D coarsest grid with at most 32 x 32 inner nodes: ONE node per thread (warp = row, lane = x), x / r / p /
// A p of the node live in registers for the whole solve, only p is mirrored in shared memory for the neighbours.
// Statement for statement the arithmetic of k_coarse_cg_smem (A p accumulated in table order, row sums by the xor
// butterfly, rows and fields added by warp_vecsum in order) -> the same iterates bit for bit, at a third of the
// latency per iteration (the coarse solve was half of a 2-D Poisson cycle).
template <int NF, int PAT>
__global__ void __launch_bounds__(1024) k2_coarse_cg_reg(const Geom g, const __grid_constant__ Dense9<NF> dn, Fields<double> xg,
                                                        Fields<double> bg, int max_it, double tol, int *iters_out)
{
    extern __shared__ double psm[];          // [NF][n][n] compact copy of p (boundary entries 0)
    __shared__ double rowsum[2][NF * 32];
    const int n = g.n, ni = n - 2, vol = n * n;
    const int row = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool act = row < ni && lane < ni;
    const int c0 = (1 + row) * n + 1 + lane;  // compact index of this thread's node
    for (int t = threadIdx.x; t < NF * vol; t += 1024) psm[t] = 0.0;
    double xv[NF], rv[NF], pv[NF], av[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        xv[i] = 0.0; av[i] = 0.0;
        rv[i] = act ? bg.p[i][node_index(g, 1 + lane, 1 + row, 0)] : 0.0;
        pv[i] = rv[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NF; ++i)
        if (act) psm[i * vol + c0] = pv[i];
    int flip = 0;
    // row sums -> rowsum[flip][field * ni + row]; then (after ONE barrier) every warp forms the identical total
    auto reduce = [&](const double (&term)[NF]) {
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            double acc = 0.0;
            if (act) acc = acc + term[i];
            acc = warp_butterfly(acc);
            if (lane == 0 && row < ni) rowsum[flip][i * ni + row] = acc;
        }
        __syncthreads();
        double total = 0.0;
#pragma unroll
        for (int i = 0; i < NF; ++i) total = total + warp_vecsum(&rowsum[flip][i * ni], ni);
        flip ^= 1;
        return total;
    };
    double term[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) term[i] = rv[i] * rv[i];
    double rr = reduce(term);     // also orders the p mirror before the first A p
    const double r0 = sqrt(rr);
    int it = 0;
    if (r0 != 0.0) {
        while (it < max_it) {
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                double a = 0.0;
                if (act) {
#pragma unroll
                    for (int j = 0; j < NF; ++j)
#pragma unroll
                        for (int q = 0; q < 9; ++q) {
                            const double cf = dn.w[i][j][q];
                            bool on;
                            if constexpr (PAT == 1) on = dense9_star_corner(i, j, q);
                            else on = cf != 0.0;
                            if (on) a = a + cf * psm[j * vol + c0 + (q / 3 - 1) * n + (q % 3 - 1)];
                        }
                }
                av[i] = a;
                term[i] = pv[i] * a;
            }
            const double pap = reduce(term);
            const double alpha = rr / pap;
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                xv[i] = xv[i] + alpha * pv[i];
                rv[i] = rv[i] - alpha * av[i];
                term[i] = rv[i] * rv[i];
            }
            const double rr_new = reduce(term);
            ++it;
            if (!(sqrt(rr_new) > tol * r0)) break;
            const double beta = rr_new / rr;
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                pv[i] = rv[i] + beta * pv[i];
                if (act) psm[i * vol + c0] = pv[i];
            }
            rr = rr_new;
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < NF; ++i)
        if (act) xg.p[i][node_index(g, 1 + lane, 1 + row, 0)] = xv[i];
    if (threadIdx.x == 0 && iters_out) *iters_out = it;
}

// Pipelined version: the 2k colour passes of k consecutive sweeps run concurrently, pass p two rows behind pass
// p-1 (row y of a pass reads rows y-1 (already updated by this pass), y, y+1 (updated by the previous pass one
// step earlier): exactly the values the sequential sweeps would see).  Each pass is a group of warps of the one
// CTA; a step is one __syncthreads.  Rows stream through a shared-memory window once (cp.async in, 16-byte stores
// out when a row has left the last pass), so k sweeps cost (n + 4k) row steps instead of 2k n.
constexpr int ROWSEQ_NT = 512;   // 2 CTAs (or other kernels of a population) fit beside it on an SM
template <int NF, int PAT>
__global__ void __launch_bounds__(ROWSEQ_NT, 2) k2_smooth_rowseq_pipe(const Geom g, const __grid_constant__ OpSten st, const __grid_constant__ Dense9<NF> dn,
                                                              const double omega, Fields<double> u, Fields<double> rhs, const int sweeps)
{
    extern __shared__ __align__(16) double win[];   // [W slots][2*NF arrays][pitch]
    const int n = g.n, pitch = g.pitch;
    const int chunks = pitch / 2;                    // 16-byte chunks per row
    const int P = 2 * sweeps, W = 2 * P + 3;
    const int SL = 2 * NF * pitch;                   // doubles per window slot
    const int G = ((ROWSEQ_NT / P) / 32) * 32;       // threads per pass (whole warps); the rest only moves data
    const int grp = threadIdx.x / G, lt = threadIdx.x - grp * G;
    // step-invariant part of this thread's share of a row transfer (<= 3 chunks in, <= 2 chunks out per thread)
    const double *in_src[3];
    int in_off[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int t = threadIdx.x + k * ROWSEQ_NT;
        in_src[k] = nullptr;
        in_off[k] = 0;
        if (t < 2 * NF * chunks) {
            const int a = t / chunks, c = t - a * chunks;
            in_src[k] = (a < NF ? u.p[a] : rhs.p[a - NF]) + 2 * c;
            in_off[k] = a * pitch + 2 * c;
        }
    }
    double *out_dst[2];
    int out_off[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int t = threadIdx.x + k * ROWSEQ_NT;
        out_dst[k] = nullptr;
        out_off[k] = 0;
        if (t < NF * chunks) {
            const int a = t / chunks, c = t - a * chunks;
            out_dst[k] = u.p[a] + 2 * c;
            out_off[k] = a * pitch + 2 * c;
        }
    }
    auto fetch = [&](int row, int slot) {
        if (row <= n - 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (in_src[k]) cp_async16(win + slot * SL + in_off[k], in_src[k] + (long long)row * pitch);
        }
        cp_async_commit();
    };
    auto store_row = [&](int row, int slot) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (out_dst[k])
                *reinterpret_cast<double2 *>(out_dst[k] + (long long)row * pitch) = *reinterpret_cast<const double2 *>(win + slot * SL + out_off[k]);
    };
    // the local matrix (diagonal entries of the blocks) is the same at every anchor: factor it once
    double M0[NF][NF];
#pragma unroll
    for (int a = 0; a < NF; ++a)
#pragma unroll
        for (int m = 0; m < NF; ++m) M0[a][m] = dn.w[a][m][4];
    DenseLU<double, NF> lu;
    dense_factor<double, NF>(M0, lu);
    // which table entries exist: tested per term in the sweep without touching the constant bank again
    unsigned nz[NF][NF];
#pragma unroll
    for (int a = 0; a < NF; ++a)
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            nz[a][j] = 0u;
#pragma unroll
            for (int p = 0; p < 9; ++p)
                if (dn.w[a][j][p] != 0.0) nz[a][j] |= 1u << p;
        }
    fetch(0, 0); fetch(1, 1); fetch(2, 2);
    const int last = (n - 2) + 2 * (P - 1);
    // window slots (row % W), advanced by one per step: row y+2 (fetch), this pass's row yr = y - 2 grp, and the row
    // that left the last pass in the previous step, yb = y - 2 (P-1) - 1
    int s_fetch = 3 % W;
    int s_r = ((1 - 2 * grp) % W + W) % W;
    int s_b = ((1 - 2 * (P - 1) - 1) % W + W) % W;
    for (int y = 1; y <= last; ++y) {
        fetch(y + 2, s_fetch);
        cp_async_wait<1>();      // everything but the newest group (row y+2) has landed
        __syncthreads();         // ... for all threads; the previous step's updates are visible
        const int yb = y - 2 * (P - 1) - 1;
        if (yb >= 1) store_row(yb, s_b);
        const int yr = y - 2 * grp;
        if (grp < P && yr >= 1 && yr <= n - 2) {
            const int color = grp & 1;
            const int s_m = s_r == 0 ? W - 1 : s_r - 1, s_p = s_r + 1 == W ? 0 : s_r + 1;
            const double *rm = win + s_m * SL, *r0 = win + s_r * SL, *rp = win + s_p * SL;   // field j at + j * pitch
            double *uc = win + s_r * SL;
            const double *fr = r0 + NF * pitch;
            for (int t = lt;; t += G) {
                const int x = 1 + 2 * t + ((1 + yr + color) & 1);
                if (x > n - 2) break;
                // all neighbour values first (independent loads), then the sums in table order; absent entries
                // (coefficient 0) are skipped by a uniform branch, so the arithmetic is that of the sparse loop
                double nbv[NF][9];
#pragma unroll
                for (int j = 0; j < NF; ++j) {
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        nbv[j][dx] = rm[j * pitch + x + dx - 1];
                        nbv[j][3 + dx] = r0[j * pitch + x + dx - 1];
                        nbv[j][6 + dx] = rp[j * pitch + x + dx - 1];
                    }
                }
                double b[NF];
#pragma unroll
                for (int a = 0; a < NF; ++a) {
                    double sacc = 0.0;
#pragma unroll
                    for (int j = 0; j < NF; ++j)
#pragma unroll
                        for (int p = 0; p < 9; ++p) {
                            if (p == 4) continue;
                            bool on;
                            if constexpr (PAT == 1) on = dense9_star_corner(a, j, p);
                            else on = (nz[a][j] >> p) & 1u;
                            if (on) sacc = sacc + dn.w[a][j][p] * nbv[j][p];
                        }
                    b[a] = fr[a * pitch + x] - sacc;
                }
                if (NF == 1) b[0] = b[0] * lu.inv[0];
                else dense_solve<double, NF>(lu, b);
#pragma unroll
                for (int a = 0; a < NF; ++a) {
                    const double old = nbv[a][4];
                    uc[a * pitch + x] = old + omega * (b[a] - old);
                }
            }
        }
        if (++s_fetch == W) s_fetch = 0;
        if (++s_r == W) s_r = 0;
        if (++s_b == W) s_b = 0;
    }
    cp_async_wait<0>();
    __syncthreads();
    store_row(n - 2, (n - 2) % W);
}

}  // namespace evo
