// evo_kernels_fas.cuh -- nonlinear full-approximation-scheme statements for
//   -Lap u + gamma u e^u = f   (reference: example_problems/FAS_2D_Basic/FAS_2D_Basic_template.exa4:19-34,
//   smoother :65-73, residual :75-80, CGS :58-63; emitted by code_generation/exastencils_FAS.py:99-319).
// Real scalar 2-D, like the reference.  Arithmetic mirrors oracle/mg_fas.inc operation by operation
// (shared evo_exp, no FMA), so residual histories are bit-identical.
#pragma once
#include "../../include/evo_math.h"
#include <cooperative_groups.h>

#include "evo_kernels.cuh"

namespace evo {
namespace fas {

struct Lin2 {  // linear 2-D stencil in ascending table order (y-1, x-1, centre, x+1, y+1 for a 5-point star)
    int nnz;
    int dx[9], dy[9];
    double c[9];
    double a00;
};

static bool make_lin2(const Sten &s, Lin2 *o)
{
    if (s.nnz > 9) return false;
    o->nnz = s.nnz;
    o->a00 = 0.0;
    for (int q = 0; q < s.nnz; ++q) {
        if (s.oz[q] != 0 || s.im[q] != 0.0) return false;
        o->dx[q] = s.ox[q]; o->dy[q] = s.oy[q]; o->c[q] = s.re[q];
        if (s.ox[q] == 0 && s.oy[q] == 0) o->a00 = s.re[q];
    }
    return true;
}

// (A u + N(u) u) at a node: linear part in table order incl. the centre, then + gamma e^v v
__device__ __forceinline__ double fas_apply(const Geom &g, const Lin2 &L, double gamma, const double *u, long long idx)
{
    double lin = 0.0;
    for (int q = 0; q < L.nnz; ++q) lin = lin + L.c[q] * u[idx + (long long)L.dy[q] * g.pitch + L.dx[q]];
    const double v = u[idx];
    return lin + gamma * evo_exp(v) * v;
}

// `steps` local (Picard / Newton) steps at one node:  v += w (f - ((nb + a00 v) + gamma e^v v)) / (a00 [+ J(v)])
__device__ __forceinline__ double fas_point(const Geom &g, const Lin2 &L, double gamma, const double *u, double fv,
                                            long long idx, bool newton, int steps, double w)
{
    double nb = 0.0;
    for (int q = 0; q < L.nnz; ++q) {
        if (L.dx[q] == 0 && L.dy[q] == 0) continue;
        nb = nb + L.c[q] * u[idx + (long long)L.dy[q] * g.pitch + L.dx[q]];
    }
    double v = u[idx];
    for (int t = 0; t < steps; ++t) {
        const double e = evo_exp(v);
        const double num = fv - ((nb + L.a00 * v) + gamma * e * v);
        const double den = newton ? L.a00 + gamma * (1.0 + v) * e : L.a00;
        v = v + w * (num / den);
    }
    return v;
}

// color < 0: Jacobi sweep src -> dst; color 0/1: one colour of the in-place red-black sweep
static __global__ void __launch_bounds__(BX) k2_fas_smooth(const Geom g, const __grid_constant__ Lin2 L, double gamma,
                                                    const double *__restrict__ src, double *dst, const double *__restrict__ f,
                                                    int newton, int steps, double w, int color)
{
    const int y = 1 + blockIdx.y;
    const int t = blockIdx.x * BX + threadIdx.x;
    const int x = color < 0 ? 1 + t : 1 + 2 * t + ((1 + y + color) & 1);
    if (x > g.n - 2) return;
    const long long idx = (long long)y * g.pitch + x;
    dst[idx] = fas_point(g, L, gamma, src, f[idx], idx, newton != 0, steps, w);
}

static __global__ void __launch_bounds__(BX) k2_fas_residual(const Geom g, const __grid_constant__ Lin2 L, double gamma,
                                                      const double *__restrict__ u, const double *__restrict__ f,
                                                      double *__restrict__ r)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y;
    if (x > g.n - 2) return;
    const long long idx = (long long)y * g.pitch + x;
    r[idx] = f[idx] - fas_apply(g, L, gamma, u, idx);
}

// RHS_c += (A + N)(APX_c)   (second half of  RHS@(l-1) = R * Residual@l + (A + N)(Approximation@(l-1)))
static __global__ void __launch_bounds__(BX) k2_fas_add_operator(const Geom g, const __grid_constant__ Lin2 L, double gamma,
                                                          const double *__restrict__ apx, double *__restrict__ rhs)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y;
    if (x > g.n - 2) return;
    const long long idx = (long long)y * g.pitch + x;
    rhs[idx] = rhs[idx] + fas_apply(g, L, gamma, apx, idx);
}

// SOL -= APX on the whole padded array (both carry identical boundary values)
static __global__ void k_sub_inplace(double *__restrict__ a, const double *__restrict__ b, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        a[i] = a[i] - b[i];
}

// CGS@coarsest: `sweeps` damped Newton-Jacobi sweeps in ONE CTA (the coarsest grid is latency bound;
// 200 separate launches would cost ~1 ms); ping-pong between the two SOL slots, result ends in `a`
static __global__ void __launch_bounds__(1024) k2_fas_coarse(const Geom g, const __grid_constant__ Lin2 L, double gamma,
                                                      double *a, double *b, const double *__restrict__ f, int sweeps, double w)
{
    const int ni = g.n - 2;
    double *src = a, *dst = b;
    for (int s = 0; s < sweeps; ++s) {
        for (int t = threadIdx.x; t < ni * ni; t += 1024) {
            const int y = 1 + t / ni, x = 1 + t % ni;
            const long long idx = (long long)y * g.pitch + x;
            dst[idx] = fas_point(g, L, gamma, src, f[idx], idx, true, 1, w);
        }
        __threadfence_block();
        __syncthreads();
        double *tmp = src; src = dst; dst = tmp;
    }
    if (src != a) {  // odd number of sweeps: copy the result back
        for (int t = threadIdx.x; t < ni * ni; t += 1024) {
            const int y = 1 + t / ni, x = 1 + t % ni;
            const long long idx = (long long)y * g.pitch + x;
            a[idx] = src[idx];
        }
    }
}

// The same solver with both SOL slots resident in shared memory (n <= 65: 2 x 33 KB) and NPT independent nodes per
// thread in flight: per sweep a node costs shared-memory loads + one exp + one division instead of L1/L2 round
// trips, and the sweep is one barrier.  Per-node arithmetic identical to fas_point (neighbour sum in table order
// from 0.0, then the Newton step), so the result is bit-identical to k2_fas_coarse.
template <int NPT>
__global__ void __launch_bounds__(1024) k2_fas_coarse_smem(const Geom g, const __grid_constant__ Lin2 L, double gamma, double *a,
                                                           const double *__restrict__ f, int sweeps, double w)
{
    extern __shared__ __align__(16) double fas_sm[];
    const int n = g.n, ni = n - 2, sp = n + 1;   // odd shared-memory pitch
    double *s0 = fas_sm, *s1 = fas_sm + (size_t)sp * n;
    for (int t = threadIdx.x; t < n * n; t += 1024) {
        const int y = t / n, x = t - y * n;
        const double v = a[(long long)y * g.pitch + x];
        s0[y * sp + x] = v;
        s1[y * sp + x] = v;      // both slots carry the boundary values
    }
    int loc[NPT];
    double fv[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        const int t = threadIdx.x + k * 1024;
        loc[k] = -1;
        fv[k] = 0.0;
        if (t < ni * ni) {
            const int y = 1 + t / ni, x = 1 + t % ni;
            loc[k] = y * sp + x;
            fv[k] = f[(long long)y * g.pitch + x];
        }
    }
    __syncthreads();
    double *src = s0, *dst = s1;
    for (int s = 0; s < sweeps; ++s) {
        double v[NPT], nb[NPT];
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            nb[k] = 0.0;
            v[k] = 0.0;
            if (loc[k] >= 0) {
                for (int q = 0; q < L.nnz; ++q) {
                    if (L.dx[q] == 0 && L.dy[q] == 0) continue;
                    nb[k] = nb[k] + L.c[q] * src[loc[k] + L.dy[q] * sp + L.dx[q]];
                }
                v[k] = src[loc[k]];
            }
        }
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const double e = evo_exp(v[k]);
            const double num = fv[k] - ((nb[k] + L.a00 * v[k]) + gamma * e * v[k]);
            const double den = L.a00 + gamma * (1.0 + v[k]) * e;
            v[k] = v[k] + w * (num / den);
        }
#pragma unroll
        for (int k = 0; k < NPT; ++k)
            if (loc[k] >= 0) dst[loc[k]] = v[k];
        __syncthreads();
        double *tmp = src; src = dst; dst = tmp;
    }
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        const int t = threadIdx.x + k * 1024;
        if (t < ni * ni) {
            const int y = 1 + t / ni, x = 1 + t % ni;
            a[(long long)y * g.pitch + x] = src[loc[k]];
        }
    }
}

// Thread-block CLUSTER version: one SM's fp64 pipe (64 lanes/clk) bounds the single-CTA solver at ~4 us per sweep
// of a 63^2 grid (exp + division per node), so the rows are split over the CTAs of a cluster, each with its rows
// (+ one halo row per side) of both SOL slots in its own shared memory.  After a sweep the first / last row is also
// stored into the neighbour CTA's halo row through distributed shared memory, and one cluster barrier
// (arrive.release / wait.acquire) ends the sweep.  Same per-node arithmetic -> bit-identical result.
template <int NPT>
__global__ void __launch_bounds__(512) k2_fas_coarse_cluster(const Geom g, const __grid_constant__ Lin2 L, double gamma, double *a,
                                                             const double *__restrict__ f, int sweeps, double w)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double fas_sm[];
    const int rank = (int)cluster.block_rank(), cs = (int)cluster.num_blocks();
    const int n = g.n, ni = n - 2, sp = n + 1;
    const int rp = (ni + cs - 1) / cs;                       // rows per CTA (trailing CTAs may have fewer / none)
    const int y0 = 1 + rank * rp, y1 = min(y0 + rp - 1, ni);
    const int rows = max(0, y1 - y0 + 1);
    double *s0 = fas_sm, *s1 = fas_sm + (size_t)(rp + 2) * sp;
    // local row lr holds global row y0 - 1 + lr
    for (int t = threadIdx.x; t < (rows + 2) * n && rows > 0; t += 512) {
        const int lr = t / n, x = t - lr * n;
        const double v = a[(long long)(y0 - 1 + lr) * g.pitch + x];
        s0[lr * sp + x] = v;
        s1[lr * sp + x] = v;
    }
    int loc[NPT];
    double fv[NPT];
    bool first[NPT], lastr[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        const int t = threadIdx.x + k * 512;
        loc[k] = -1; fv[k] = 0.0; first[k] = false; lastr[k] = false;
        if (t < rows * ni) {
            const int ly = t / ni, x = 1 + t % ni;
            loc[k] = (ly + 1) * sp + x;
            fv[k] = f[(long long)(y0 + ly) * g.pitch + x];
            first[k] = ly == 0 && rank > 0;
            lastr[k] = ly == rows - 1 && rank + 1 < cs && y1 < ni;   // a next CTA with rows exists
        }
    }
    cluster.sync();
    double *src = s0, *dst = s1;
    for (int s = 0; s < sweeps; ++s) {
        double v[NPT], nb[NPT];
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            nb[k] = 0.0;
            v[k] = 0.0;
            if (loc[k] >= 0) {
                for (int q = 0; q < L.nnz; ++q) {
                    if (L.dx[q] == 0 && L.dy[q] == 0) continue;
                    nb[k] = nb[k] + L.c[q] * src[loc[k] + L.dy[q] * sp + L.dx[q]];
                }
                v[k] = src[loc[k]];
            }
        }
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const double e = evo_exp(v[k]);
            const double num = fv[k] - ((nb[k] + L.a00 * v[k]) + gamma * e * v[k]);
            const double den = L.a00 + gamma * (1.0 + v[k]) * e;
            v[k] = v[k] + w * (num / den);
        }
        double *up_dst = rank > 0 ? cluster.map_shared_rank(dst, rank - 1) : nullptr;
        double *dn_dst = rank + 1 < cs ? cluster.map_shared_rank(dst, rank + 1) : nullptr;
#pragma unroll
        for (int k = 0; k < NPT; ++k)
            if (loc[k] >= 0) {
                dst[loc[k]] = v[k];
                const int x = loc[k] % sp;
                if (first[k]) up_dst[(rp + 1) * sp + x] = v[k];    // neighbour above always holds rp rows
                if (lastr[k]) dn_dst[x] = v[k];                    // its halo row 0
            }
        cluster.sync();
        double *tmp = src; src = dst; dst = tmp;
    }
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        const int t = threadIdx.x + k * 512;
        if (t < rows * ni) {
            const int ly = t / ni, x = 1 + t % ni;
            a[(long long)(y0 + ly) * g.pitch + x] = src[loc[k]];
        }
    }
}

}  // namespace fas
}  // namespace evo
