// evo_kernels_fas.cuh -- nonlinear full-approximation-scheme statements for
//   -Lap u + gamma u e^u = f   (reference: example_problems/FAS_2D_Basic/FAS_2D_Basic_template.exa4:19-34,
//   smoother :65-73, residual :75-80, CGS :58-63; emitted by code_generation/exastencils_FAS.py:99-319).
// Real scalar 2-D, like the reference.  Arithmetic mirrors oracle/mg_fas.inc operation by operation
// (shared evo_exp, no FMA), so residual histories are bit-identical.
#pragma once
#include "../../include/evo_math.h"
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
__global__ void __launch_bounds__(BX) k2_fas_smooth(const Geom g, const __grid_constant__ Lin2 L, double gamma,
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

__global__ void __launch_bounds__(BX) k2_fas_residual(const Geom g, const __grid_constant__ Lin2 L, double gamma,
                                                      const double *__restrict__ u, const double *__restrict__ f,
                                                      double *__restrict__ r)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y;
    if (x > g.n - 2) return;
    const long long idx = (long long)y * g.pitch + x;
    r[idx] = f[idx] - fas_apply(g, L, gamma, u, idx);
}

// RHS_c += (A + N)(APX_c)   (second half of  RHS@(l-1) = R * Residual@l + (A + N)(Approximation@(l-1)))
__global__ void __launch_bounds__(BX) k2_fas_add_operator(const Geom g, const __grid_constant__ Lin2 L, double gamma,
                                                          const double *__restrict__ apx, double *__restrict__ rhs)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y;
    if (x > g.n - 2) return;
    const long long idx = (long long)y * g.pitch + x;
    rhs[idx] = rhs[idx] + fas_apply(g, L, gamma, apx, idx);
}

// SOL -= APX on the whole padded array (both carry identical boundary values)
__global__ void k_sub_inplace(double *__restrict__ a, const double *__restrict__ b, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        a[i] = a[i] - b[i];
}

// CGS@coarsest: `sweeps` damped Newton-Jacobi sweeps in ONE CTA (the coarsest grid is latency bound;
// 200 separate launches would cost ~1 ms); ping-pong between the two SOL slots, result ends in `a`
__global__ void __launch_bounds__(1024) k2_fas_coarse(const Geom g, const __grid_constant__ Lin2 L, double gamma,
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

}  // namespace fas
}  // namespace evo
