// evo_kernels_helm.cuh -- Helmholtz 2-D (complex fp64, one double2 per unknown): Robin boundary function,
// coarsest-level BiCGStab and the vector kernels of the outer preconditioned BiCGStab
// (reference: example_problems/Helmholtz/2D_FD_Helmholtz_fromL3.exa3:144-200, :396-433; .exa4:25-145).
// Operation order mirrors oracle/mg_krylov.inc + orc_helmholtz_solve (explicit Smith division, plain
// complex products, canonical reductions) so that residual histories are bit-identical.
#pragma once
#include "evo_kernels.cuh"

namespace evo {
namespace helm {

struct HelmState {
    cplx alpha, beta, rho, rho_new, omega;
    cplx dot[4];
    double init, cur;
    int it, done, bad, timed_out;
    unsigned long long t_start, timeout_ns;   // watchdog (see SolveState)
};

// u_b = u_neighbour * rden on x = 0 and x = n-1 (rows 1..n-2), 0 on the rows y = 0 and y = n-1 (corners end 0)
__device__ __forceinline__ void bc_node_rows(const Geom &g, cplx *u, cplx rden, int t)
{
    const int n = g.n;
    if (t >= n) return;
    if (t == 0 || t == n - 1) {
        for (int x = 0; x < n; ++x) u[(long long)t * g.pitch + x] = cplx(0.0, 0.0);
    } else {
        u[(long long)t * g.pitch] = u[(long long)t * g.pitch + 1] * rden;
        u[(long long)t * g.pitch + n - 1] = u[(long long)t * g.pitch + n - 2] * rden;
    }
}
static __global__ void k2_helm_bc(const Geom g, cplx *u, cplx rden)
{
    bc_node_rows(g, u, rden, blockIdx.x * blockDim.x + threadIdx.x);
}

// canonical unconjugated complex dot over inner nodes: rows[(y-1)] per warp, then one warp
static __global__ void __launch_bounds__(128) k2_cdot_rows(const Geom g, const cplx *a, const cplx *b, cplx *rows)
{
    const int ni = g.n - 2;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= ni) return;
    cplx s = warp_row_dot<cplx, cplx>(a, b, (long long)(row + 1) * g.pitch + 1, ni);
    if ((threadIdx.x & 31) == 0) rows[row] = s;
}
static __global__ void __launch_bounds__(32) k2_cdot_final(const cplx *rows, int ni, HelmState *st, int slot)
{
    cplx s = warp_vecsum(rows, ni);
    if (threadIdx.x == 0) st->dot[slot] = s;
}

__device__ __forceinline__ double abs_sqrt(cplx z) { return sqrt(sqrt(z.re * z.re + z.im * z.im)); }  // |sqrt(z)|

// scalar recurrences of the outer iteration; stage: 0 = init, 1 = beta, 2 = alpha, 3 = omega, 4 = convergence
static __global__ void k_helm_scalar(HelmState *st, double *hist, double tol, int max_iters, int stage)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (stage == 0) {
        st->init = abs_sqrt(st->dot[0]);
        st->cur = st->init;
        hist[0] = st->init;
        st->alpha = cplx(1.0); st->beta = cplx(1.0); st->rho_new = cplx(1.0); st->omega = cplx(1.0);
        st->it = 0; st->bad = 0; st->timed_out = 0;
        st->t_start = global_timer_ns();
        st->done = (st->init == 0.0 || max_iters <= 0) ? 1 : 0;
        if (!isfinite(st->init)) { st->bad = 1; st->done = 1; }
    } else if (stage == 1) {
        st->rho = st->rho_new;
        st->rho_new = st->dot[0];
        st->beta = (st->rho_new / st->rho) * (st->alpha / st->omega);
    } else if (stage == 2) {
        st->alpha = st->rho_new / st->dot[1];
    } else if (stage == 3) {
        st->omega = st->dot[2] / st->dot[3];
    } else {
        if (st->done) return;
        st->cur = abs_sqrt(st->dot[0]);
        st->it += 1;
        hist[st->it] = st->cur;
        if (!isfinite(st->cur)) { st->bad = 1; st->done = 1; }
        else if (st->cur < tol * st->init || st->it >= max_iters) st->done = 1;
        else if (st->timeout_ns && global_timer_ns() - st->t_start > st->timeout_ns) { st->timed_out = 1; st->done = 1; }
    }
}

// mode 0: p = r + beta (p - omega ap)        mode 1: h = x + alpha u ; s = r - alpha ap
// mode 2: x = h + omega u                    mode 3: r = s - omega t
// mode 4: dst = src (inner nodes)
static __global__ void __launch_bounds__(BX) k2_helm_vec(const Geom g, int mode, const HelmState *st, cplx *a, cplx *b,
                                                  const cplx *c, const cplx *d, const cplx *e, const cplx *f)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y;
    if (x > g.n - 2) return;
    const long long i = (long long)y * g.pitch + x;
    if (mode == 0) a[i] = c[i] + st->beta * (a[i] - st->omega * d[i]);            // a=p c=r d=ap
    else if (mode == 1) { a[i] = c[i] + st->alpha * d[i]; b[i] = e[i] - st->alpha * f[i]; }  // a=h b=s c=x d=u e=r f=ap
    else if (mode == 2) a[i] = c[i] + st->omega * d[i];                           // a=x c=h d=u
    else if (mode == 3) a[i] = c[i] - st->omega * d[i];                           // a=r c=s d=t
    else a[i] = c[i];
}

// out = A * u on inner nodes (A = un-shifted operator of the finest level)
static __global__ void __launch_bounds__(BX) k2_apply_op(const Geom g, const __grid_constant__ OpSten st, const cplx *u, cplx *out,
                                                  const cplx *minus_from /* nullable: out = minus_from - A u */)
{
    const int x = 1 + blockIdx.x * BX + threadIdx.x, y = 1 + blockIdx.y;
    if (x > g.n - 2) return;
    const long long i = (long long)y * g.pitch + x;
    Fields<cplx> uf;
    uf.p[0] = const_cast<cplx *>(u); uf.p[1] = nullptr;
    cplx acc = apply_row<cplx, 1>(g, st, uf, 0, i);
    out[i] = minus_from ? minus_from[i] - acc : acc;
}

static __global__ void k_set_while_condition_helm(cudaGraphConditionalHandle handle, const HelmState *st)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) cudaGraphSetConditional(handle, st->done ? 0u : 1u);
}

// gen_mgCycle@coarsest: BiCGStab on M, zero initial guess, one CTA (exa3:396-433).  Vectors in global memory
// (the coarsest Helmholtz grid is 9 x 9); boundary function on x, r, p, s like the oracle.
static __global__ void __launch_bounds__(1024) k2_coarse_bicgstab(const Geom g, const __grid_constant__ OpSten st, cplx rden, cplx *x,
                                                           const cplx *b, cplx *r, cplx *rh, cplx *p, cplx *nu, cplx *s, cplx *t,
                                                           cplx *hh, cplx *rows, int max_it, double tol, int robin)
{
    const int n = g.n, ni = n - 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Fields<cplx> fp; fp.p[1] = nullptr;
    auto dot = [&](const cplx *a_, const cplx *b_) {
        for (int row = warp; row < ni; row += 32) {
            cplx sacc = warp_row_dot<cplx, cplx>(a_, b_, (long long)(row + 1) * g.pitch + 1, ni);
            if (lane == 0) rows[row] = sacc;
        }
        __threadfence_block();
        __syncthreads();
        cplx total = warp_vecsum(rows, ni);   // every warp recomputes the identical final sum
        __syncthreads();
        return total;
    };
    auto bc = [&](cplx *u) {
        if (robin) for (int tt = threadIdx.x; tt < n; tt += 1024) bc_node_rows(g, u, rden, tt);
        __threadfence_block();
        __syncthreads();
    };
    auto inner = [&](auto fn) {
        for (int tt = threadIdx.x; tt < ni * ni; tt += 1024) fn((long long)(1 + tt / ni) * g.pitch + 1 + tt % ni);
        __threadfence_block();
        __syncthreads();
    };
    for (long long tt = threadIdx.x; tt < g.total; tt += 1024) {
        x[tt] = cplx(0.0); rh[tt] = cplx(0.0); p[tt] = cplx(0.0); nu[tt] = cplx(0.0); s[tt] = cplx(0.0); t[tt] = cplx(0.0); hh[tt] = cplx(0.0);
    }
    __syncthreads();
    inner([&](long long i) { fp.p[0] = x; r[i] = b[i] - apply_row<cplx, 1>(g, st, fp, 0, i); });
    bc(r);
    const cplx d0 = dot(r, r);
    const double init = abs_sqrt(d0);
    if (init == 0.0) return;
    double cur = init;
    cplx alpha(1.0), beta(1.0), rho, rho_new(1.0), omega(1.0);
    for (long long tt = threadIdx.x; tt < g.total; tt += 1024) rh[tt] = r[tt];
    __syncthreads();
    int it = 0;
    while (it < max_it) {
        rho = rho_new;
        rho_new = dot(rh, r);
        beta = (rho_new / rho) * (alpha / omega);
        inner([&](long long i) { p[i] = r[i] + beta * (p[i] - omega * nu[i]); });
        bc(p);
        inner([&](long long i) { fp.p[0] = p; nu[i] = apply_row<cplx, 1>(g, st, fp, 0, i); });
        alpha = rho_new / dot(rh, nu);
        inner([&](long long i) { hh[i] = x[i] + alpha * p[i]; s[i] = r[i] - alpha * nu[i]; });
        bc(s);
        inner([&](long long i) { fp.p[0] = s; t[i] = apply_row<cplx, 1>(g, st, fp, 0, i); });
        const cplx ts = dot(t, s), tt2 = dot(t, t);
        omega = ts / tt2;
        inner([&](long long i) { x[i] = hh[i] + omega * s[i]; r[i] = s[i] - omega * t[i]; });
        bc(x);
        bc(r);
        cur = abs_sqrt(dot(r, r));
        ++it;
        if (!(cur >= tol * init)) break;
    }
}

}  // namespace helm
}  // namespace evo
