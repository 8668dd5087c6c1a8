// evo_kernels_warp2d.cuh -- register-streamed pointwise sweeps for large 2-D grids, 5-point star (sm_100a).
//
// Statements: `solve locally at u@l [with jacobi] relax w {...}` with or without `color with { (i0+i1)%2 }`
// (exastencils.py:659-682, :769-822) on a scalar real 5-point operator, and the Newton / Picard Jacobi smoother of the
// FAS template (FAS_2D_Basic_template.exa4:65-73).  BASELINE configs[2] (FAS at 4097^2) and Poisson 2-D at 4097^2 ran
// on the generic one-node-per-thread kernels at 28-36 % of the HBM peak (every sweep = 24 B/node, neighbours through
// L1/L2, one launch per colour).
//
// Design: a WARP owns a strip of 64 consecutive nodes (32 x-pairs, 16-byte aligned) and marches through a chunk of
// rows.  Rows are contiguous in HBM, so each lane fetches its pair with one coalesced 16-byte load, two rows ahead of
// its use; the y-neighbours live in a 3-row register window, the x-neighbour outside the pair comes from the adjacent
// lane by shuffle -- no shared memory, no block barrier, warps are independent.  NSTAGE dependent half-sweeps (RB-GS:
// colours; Jacobi: whole sweeps) run as pipeline stages one row apart, each with its own register window: temporal
// blocking -- k consecutive sweeps cost ONE pass over HBM (24 B/node for all of them).  A stage invalidates one more
// node at each end of the strip (its x-neighbour is missing there), so a strip delivers 64 - 2*NSTAGE nodes; strips
// and row chunks overlap accordingly (redundant halo work is recomputed identically: exact sequential semantics).
// Out of place: reads SOL, writes the [next] slot.  Per node the arithmetic is that of the generic kernels / the
// oracle (sum in ascending table order y-1, x-1, x+1, y+1; x = (f - s)*(1/a); u += w (x - u)), bit for bit.
#pragma once
#include "../../include/evo_math.h"
#include "evo_kernels.cuh"

namespace evo {
namespace w2 {

struct Star5 {  // ascending table order: y-1, x-1, centre, x+1, y+1
    double ym, xm, c, xp, yp;
};

static bool match_star5(const Sten &s, Star5 *out)
{
    static const signed char ex[5][2] = {{0, -1}, {-1, 0}, {0, 0}, {1, 0}, {0, 1}};
    if (s.nnz != 5) return false;
    for (int q = 0; q < 5; ++q)
        if (s.ox[q] != ex[q][0] || s.oy[q] != ex[q][1] || s.oz[q] != 0 || s.im[q] != 0.0) return false;
    out->ym = s.re[0]; out->xm = s.re[1]; out->c = s.re[2]; out->xp = s.re[3]; out->yp = s.re[4];
    return true;
}

// pointwise solve of the linear equation (weighted Jacobi / one colour of RB-GS)
struct LinearPoint {
    Star5 s;
    double inv_c, omega;
    __device__ __forceinline__ double operator()(double ym, double xm, double c, double xp, double yp, double f) const
    {
        double sum = 0.0;
        sum = sum + s.ym * ym;
        sum = sum + s.xm * xm;
        sum = sum + s.xp * xp;
        sum = sum + s.yp * yp;
        const double xs = (f - sum) * inv_c;
        return c + omega * (xs - c);
    }
};

// `steps` damped Newton (or Picard) steps of  -Lap v + gamma v e^v = f  at one node (fas::fas_point, mg_fas.inc)
struct FasPoint {
    Star5 s;
    double gamma, w;
    int newton, steps;
    __device__ __forceinline__ double operator()(double ym, double xm, double c, double xp, double yp, double f) const
    {
        double nb = 0.0;
        nb = nb + s.ym * ym;
        nb = nb + s.xm * xm;
        nb = nb + s.xp * xp;
        nb = nb + s.yp * yp;
        double v = c;
        for (int t = 0; t < steps; ++t) {
            const double e = evo_exp(v);
            const double num = f - ((nb + s.c * v) + gamma * e * v);
            const double den = newton ? s.c + gamma * (1.0 + v) * e : s.c;
            v = v + w * (num / den);
        }
        return v;
    }
};

constexpr int W2_WARPS = 4;      // warps (adjacent strips) per CTA
constexpr int W2_PF = 2;         // rows of load prefetch

struct Pair {
    double l, r;
};

__device__ __forceinline__ Pair load_pair(const double *__restrict__ row, int xl, int n)
{
    Pair p;
    p.l = 0.0; p.r = 0.0;
    if (xl >= 0) {
        if (xl + 1 <= n - 1) {
            const double2 v = __ldg(reinterpret_cast<const double2 *>(row + xl));
            p.l = v.x; p.r = v.y;
        } else if (xl <= n - 1) {
            p.l = __ldg(row + xl);
        }
    }
    return p;
}

// NSTAGE pipeline stages; RB: stage s is colour s & 1 of sweep s / 2; else every stage is a whole Jacobi sweep.
template <class U, int NSTAGE, bool RB>
__global__ void __launch_bounds__(W2_WARPS * 32) k2_sweep_warp(const Geom g, const U upd, const double *__restrict__ u,
                                                              const double *__restrict__ f, double *__restrict__ out,
                                                              const int rows_per_chunk)
{
    constexpr int WINT = 64 - 2 * NSTAGE;                 // nodes a strip delivers
    const int n = g.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * W2_WARPS + warp;
    const int X0 = strip * WINT;                          // strip = nodes X0 .. X0+63, first pair at even x
    if (X0 + NSTAGE > n - 2 && strip > 0) return;         // nothing to deliver (warps are independent: no barrier below)
    const int xl = X0 + 2 * lane, xr = xl + 1;
    const int out_lo = strip == 0 ? 1 : X0 + NSTAGE, out_hi = min(X0 + 63 - NSTAGE, n - 2);
    const bool in_l = xl >= 1 && xl <= n - 2, in_r = xr >= 1 && xr <= n - 2;
    const bool st_l = xl >= out_lo && xl <= out_hi, st_r = xr >= out_lo && xr <= out_hi;
    const int ya = 1 + blockIdx.y * rows_per_chunk, yb = min(ya + rows_per_chunk - 1, n - 2);
    const int rlo = max(0, ya - NSTAGE), rhi = min(n - 1, yb + NSTAGE);   // raw rows this chunk reads
    const long long pitch = g.pitch;

    // win[b]: rows (r-1, r, r+1) of the values entering stage b; fw[s]: right-hand side of the row stage s works on
    Pair win[NSTAGE][3];
    Pair fw[NSTAGE];
    Pair pre_u[W2_PF], pre_f[W2_PF];
#pragma unroll
    for (int b = 0; b < NSTAGE; ++b) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { win[b][k].l = 0.0; win[b][k].r = 0.0; }
        fw[b].l = 0.0; fw[b].r = 0.0;
    }
    // step t: stage s works on row t - s; raw row t + 1 and the right-hand side of row t enter the pipeline
    const int t_first = rlo, t_last = yb + NSTAGE - 1;
    // prologue: window rows rlo - 1 (unused), rlo; prefetch ring holds raw rows t+1 .. t+PF and f rows t .. t+PF-1
    win[0][2] = load_pair(u + (long long)rlo * pitch, xl, n);
#pragma unroll
    for (int k = 0; k < W2_PF; ++k) {
        const int ru = t_first + 1 + k, rf = t_first + k;
        pre_u[k] = ru <= rhi ? load_pair(u + (long long)ru * pitch, xl, n) : Pair{0.0, 0.0};
        pre_f[k] = rf <= rhi ? load_pair(f + (long long)rf * pitch, xl, n) : Pair{0.0, 0.0};
    }
    for (int t = t_first; t <= t_last; ++t) {
        // rotate the windows, take raw row t+1 / rhs row t out of the prefetch ring, refill the ring
#pragma unroll
        for (int b = 0; b < NSTAGE; ++b) { win[b][0] = win[b][1]; win[b][1] = win[b][2]; }
#pragma unroll
        for (int s = NSTAGE - 1; s > 0; --s) fw[s] = fw[s - 1];
        win[0][2] = pre_u[0];
        fw[0] = pre_f[0];
#pragma unroll
        for (int k = 0; k + 1 < W2_PF; ++k) { pre_u[k] = pre_u[k + 1]; pre_f[k] = pre_f[k + 1]; }
        {
            const int ru = t + 1 + W2_PF, rf = t + W2_PF;
            pre_u[W2_PF - 1] = ru <= rhi ? load_pair(u + (long long)ru * pitch, xl, n) : Pair{0.0, 0.0};
            pre_f[W2_PF - 1] = rf <= rhi ? load_pair(f + (long long)rf * pitch, xl, n) : Pair{0.0, 0.0};
        }
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            const int r = t - s;
            if (r < rlo || r > rhi) continue;                                  // warp uniform
            // rows this stage updates: the chunk plus the halo the later stages still need, inner rows only
            const int a_lo = max(1, ya - (NSTAGE - 1 - s)), a_hi = min(n - 2, yb + (NSTAGE - 1 - s));
            Pair o = win[s][1];                                                // default: pass through
            if (r >= a_lo && r <= a_hi) {
                const Pair m = win[s][1], up = win[s][0], dn = win[s][2];
                // x-neighbours outside the pair come from the adjacent lanes
                const double left = __shfl_up_sync(0xffffffffu, m.r, 1);
                const double right = __shfl_down_sync(0xffffffffu, m.l, 1);
                if (RB) {
                    // colour of stage s: nodes with (x + y) % 2 == s % 2; x_l is even
                    const bool l_active = ((r + s) & 1) == 0;
                    if (l_active) { if (in_l) o.l = upd(up.l, left, m.l, m.r, dn.l, fw[s].l); }
                    else { if (in_r) o.r = upd(up.r, m.l, m.r, right, dn.r, fw[s].r); }
                } else {
                    if (in_l) o.l = upd(up.l, left, m.l, m.r, dn.l, fw[s].l);
                    if (in_r) o.r = upd(up.r, m.l, m.r, right, dn.r, fw[s].r);
                }
            }
            if (s + 1 < NSTAGE) {
                win[s + 1][2] = o;
            } else if (r >= ya && r <= yb) {
                double *dst = out + (long long)r * pitch + xl;
                if (st_l && st_r) *reinterpret_cast<double2 *>(dst) = make_double2(o.l, o.r);
                else if (st_l) dst[0] = o.l;
                else if (st_r) dst[1] = o.r;
            }
        }
    }
}

template <class U, int NSTAGE, bool RB>
static bool launch_sweep(int sm_count, const Geom &g, const U &upd, const double *u, const double *f, double *out, cudaStream_t s)
{
    constexpr int WINT = 64 - 2 * NSTAGE;
    const int inner = g.n - 2;
    // strips: strip k delivers nodes k*WINT + NSTAGE .. k*WINT + 63 - NSTAGE (strip 0 from node 1)
    int strips = 1;
    while (strips * WINT + 63 - NSTAGE - WINT < inner) ++strips;     // last delivered node >= n - 2
    const int bx = (strips + W2_WARPS - 1) / W2_WARPS;
    // row chunks: ~1.5 waves of 64 warps per SM when the grid allows it, at least max(4, 2 NSTAGE) rows each (a chunk recomputes
    // NSTAGE halo rows at both ends)
    const long long want_warps = (long long)sm_count * 96;
    const int min_rows = std::max(4, 2 * NSTAGE);
    int chunks = (int)std::min<long long>(std::max<long long>(1, want_warps / strips), std::max(1, inner / min_rows));
    int rows = (inner + chunks - 1) / chunks;
    chunks = (inner + rows - 1) / rows;
    k2_sweep_warp<U, NSTAGE, RB><<<dim3(bx, chunks), W2_WARPS * 32, 0, s>>>(g, upd, u, f, out, rows);
    return cudaGetLastError() == cudaSuccess;
}

// minimum grid size for the streaming path (smaller grids are latency bound: the one-node-per-thread kernels are as
// good or better).  EVO_STAR2D: 0 = never, 1 = default threshold, n >= 2 = use the streaming kernels from n nodes per
// dimension (tests).  W2_MIN_N is the smallest size the kernels support at all (slot allocation).
constexpr int W2_MIN_N = 65;
inline int min_n() { const int v = option(OPT_STAR2D); return v <= 0 ? (1 << 30) : (v == 1 ? 513 : std::max(v, W2_MIN_N)); }

// k (1 or 2) consecutive weighted-Jacobi sweeps u -> out
static bool try_jacobi(int sm_count, const Geom &g, const OpSten &st, const double *u, const double *f, double *out, double omega, int k,
                       cudaStream_t s)
{
    Star5 c;
    if (g.dim != 2 || g.n < min_n() || !match_star5(st.s[0][0], &c)) return false;
    LinearPoint upd{c, 1.0 / c.c, omega};
    if (k == 1) return launch_sweep<LinearPoint, 1, false>(sm_count, g, upd, u, f, out, s);
    if (k == 2) return launch_sweep<LinearPoint, 2, false>(sm_count, g, upd, u, f, out, s);
    return false;
}

// k (1 or 2) consecutive red-black Gauss-Seidel sweeps u -> out
static bool try_rbgs(int sm_count, const Geom &g, const OpSten &st, const double *u, const double *f, double *out, double omega, int k,
                     cudaStream_t s)
{
    Star5 c;
    if (g.dim != 2 || g.n < min_n() || !match_star5(st.s[0][0], &c)) return false;
    LinearPoint upd{c, 1.0 / c.c, omega};
    if (k == 1) return launch_sweep<LinearPoint, 2, true>(sm_count, g, upd, u, f, out, s);
    if (k == 2) return launch_sweep<LinearPoint, 4, true>(sm_count, g, upd, u, f, out, s);
    return false;
}

static bool star5_applicable(const Geom &g, const OpSten &st)
{
    Star5 c;
    return g.dim == 2 && g.n >= min_n() && match_star5(st.s[0][0], &c);
}

// ---------------------------------------------------------------------------------------------
// u@l += w * (P@(l-1) * e@(l-1)), bilinear prolongation, 2-D: one thread per coarse cell (X, Y) updates the four fine
// nodes (2X..2X+1, 2Y..2Y+1) with two coalesced 16-byte read-modify-writes.  Per fine node the terms are added in
// ascending stencil-table order of the offsets o with x + o even (the order of k_prolong and of the oracle).
struct Dense9 { double w[9]; };   // index (oy + 1) * 3 + (ox + 1)

static __global__ void __launch_bounds__(128) k2_prolong_add(const Geom gf, const Geom gc, const Dense9 P, const double *__restrict__ ec,
                                                      double *__restrict__ u, const double weight)
{
    const int X = blockIdx.x * 128 + threadIdx.x, Y = blockIdx.y;
    if (X > gc.n - 2) return;
    const double *r0 = ec + (long long)Y * gc.pitch + X, *r1 = r0 + gc.pitch;
    const double e00 = r0[0], e01 = r0[1], e10 = r1[0], e11 = r1[1];     // e[dy][dx]
    const int n2 = gf.n - 2;
    // fine row 2Y (even): even x takes o = (0, 0); odd x takes o = (0, -1) then (0, +1)
    if (Y > 0) {
        double2 *up = reinterpret_cast<double2 *>(u + (long long)(2 * Y) * gf.pitch + 2 * X);
        double2 v = *up;
        double p0 = 0.0, p1 = 0.0;
        p0 = p0 + P.w[4] * e00;
        p1 = p1 + P.w[3] * e00;
        p1 = p1 + P.w[5] * e01;
        if (X > 0) v.x = v.x + weight * p0;     // fine x = 0 is the boundary layer
        v.y = v.y + weight * p1;                // fine x = 2X+1 <= n-2 always
        *up = v;
    }
    // fine row 2Y+1 (odd): even x takes o = (-1, 0) then (+1, 0); odd x the four corners in table order
    if (2 * Y + 1 <= n2) {
        double2 *up = reinterpret_cast<double2 *>(u + (long long)(2 * Y + 1) * gf.pitch + 2 * X);
        double2 v = *up;
        double p0 = 0.0, p1 = 0.0;
        p0 = p0 + P.w[1] * e00;
        p0 = p0 + P.w[7] * e10;
        p1 = p1 + P.w[0] * e00;
        p1 = p1 + P.w[2] * e01;
        p1 = p1 + P.w[6] * e10;
        p1 = p1 + P.w[8] * e11;
        if (X > 0) v.x = v.x + weight * p0;
        v.y = v.y + weight * p1;
        *up = v;
    }
}

static bool try_prolong_add(const Geom &gf, const Geom &gc, const TransferW &P, const double *src, double *dst, double weight, cudaStream_t s)
{
    if (gf.dim != 2 || gf.n < min_n()) return false;
    Dense9 W;
    for (int i = 0; i < 9; ++i) W.w[i] = 0.0;
    for (int q = 0; q < P.nnz; ++q) {
        if (P.oz[q] != 0) return false;
        W.w[(P.oy[q] + 1) * 3 + (P.ox[q] + 1)] = P.w[q];
    }
    const int cells = gc.n - 1;
    k2_prolong_add<<<dim3((cells + 127) / 128, cells), 128, 0, s>>>(gf, gc, W, src, dst, weight);
    return cudaGetLastError() == cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// Fused RHS@(l-1) = R@l * (f@l - A@l u@l), 2-D 5-point operator, dense 9-point restriction: the fine residual never
// reaches HBM (16 B per fine node read + 2 B written instead of 24 + 10).  Same warp-strip streaming as the sweeps:
// a lane owns the fine pair (2X, 2X+1) and the coarse node X; the residual rows 2Y-1, 2Y, 2Y+1 live in a register
// window, the residual at 2X-1 comes from the left lane by shuffle.  The 9 terms are added in ascending table order
// like k_residual_restrict / the oracle; the residual counts as 0 on the boundary layer.
static __global__ void __launch_bounds__(W2_WARPS * 32) k2_residual_restrict_warp(const Geom gf, const Geom gc, const Star5 c, const Dense9 R,
                                                                          const double *__restrict__ u, const double *__restrict__ f,
                                                                          double *__restrict__ dst, double *__restrict__ zero,
                                                                          const int rows_per_chunk)
{
    const int n = gf.n, nc = gc.n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * W2_WARPS + warp;
    const int X0 = strip * 62;                             // fine nodes X0 .. X0+63; lanes 1..31 deliver coarse nodes X0/2 + lane
    if (X0 / 2 + 1 > nc - 2) return;
    const int xl = X0 + 2 * lane, xr = xl + 1;
    const int X = xl / 2;
    const bool in_l = xl >= 1 && xl <= n - 2, in_r = xr >= 1 && xr <= n - 2;
    const bool deliver = lane >= 1 && X >= 1 && X <= nc - 2;
    const int Ya = 1 + blockIdx.y * rows_per_chunk, Yb = min(Ya + rows_per_chunk - 1, nc - 2);
    const long long pitch = gf.pitch;
    const int r_first = 2 * Ya - 1, r_last = 2 * Yb + 1;    // fine residual rows of the chunk (all inner rows)
    Pair um, u0, up;                                        // u rows r-1, r, r+1
    Pair rm = {0.0, 0.0}, r0 = {0.0, 0.0};                  // residual rows r-2, r-1
    um = load_pair(u + (long long)(r_first - 1) * pitch, xl, n);
    u0 = load_pair(u + (long long)r_first * pitch, xl, n);
    Pair pre_u[W2_PF], pre_f[W2_PF];
#pragma unroll
    for (int k = 0; k < W2_PF; ++k) {
        const int ru = r_first + 1 + k, rf = r_first + k;
        pre_u[k] = ru <= n - 1 ? load_pair(u + (long long)ru * pitch, xl, n) : Pair{0.0, 0.0};
        pre_f[k] = rf <= r_last ? load_pair(f + (long long)rf * pitch, xl, n) : Pair{0.0, 0.0};
    }
    // the right neighbour of the strip's last node (lane 31 needs it for the residual at its own right node)
    double xtra0 = (lane == 31 && xr + 1 <= n - 1) ? __ldg(u + (long long)r_first * pitch + xr + 1) : 0.0;
    for (int r = r_first; r <= r_last; ++r) {
        up = pre_u[0];
        const Pair fv = pre_f[0];
#pragma unroll
        for (int k = 0; k + 1 < W2_PF; ++k) { pre_u[k] = pre_u[k + 1]; pre_f[k] = pre_f[k + 1]; }
        {
            const int ru = r + 1 + W2_PF, rf = r + W2_PF;
            pre_u[W2_PF - 1] = ru <= min(r_last + 1, n - 1) ? load_pair(u + (long long)ru * pitch, xl, n) : Pair{0.0, 0.0};
            pre_f[W2_PF - 1] = rf <= r_last ? load_pair(f + (long long)rf * pitch, xl, n) : Pair{0.0, 0.0};
        }
        const double xtra_next = (lane == 31 && xr + 1 <= n - 1 && r + 1 <= r_last) ? __ldg(u + (long long)(r + 1) * pitch + xr + 1) : 0.0;
        const double left = __shfl_up_sync(0xffffffffu, u0.r, 1);
        double right = __shfl_down_sync(0xffffffffu, u0.l, 1);
        if (lane == 31) right = xtra0;
        // residual of row r (A u in ascending table order: y-1, x-1, centre, x+1, y+1)
        Pair res = {0.0, 0.0};
        if (in_l) {
            double sum = 0.0;
            sum = sum + c.ym * um.l; sum = sum + c.xm * left; sum = sum + c.c * u0.l; sum = sum + c.xp * u0.r; sum = sum + c.yp * up.l;
            res.l = fv.l - sum;
        }
        if (in_r) {
            double sum = 0.0;
            sum = sum + c.ym * um.r; sum = sum + c.xm * u0.l; sum = sum + c.c * u0.r; sum = sum + c.xp * right; sum = sum + c.yp * up.r;
            res.r = fv.r - sum;
        }
        if (((r - r_first) & 1) == 0 && r > r_first) {
            // r = 2Y+1: rows 2Y-1 (rm), 2Y (r0), 2Y+1 (res) are complete
            const int Y = (r - 1) / 2;
            const double a_m = __shfl_up_sync(0xffffffffu, rm.r, 1), a_0 = __shfl_up_sync(0xffffffffu, r0.r, 1),
                         a_p = __shfl_up_sync(0xffffffffu, res.r, 1);     // residual at 2X-1
            double acc = 0.0;
            acc = acc + R.w[0] * a_m; acc = acc + R.w[1] * rm.l; acc = acc + R.w[2] * rm.r;
            acc = acc + R.w[3] * a_0; acc = acc + R.w[4] * r0.l; acc = acc + R.w[5] * r0.r;
            acc = acc + R.w[6] * a_p; acc = acc + R.w[7] * res.l; acc = acc + R.w[8] * res.r;
            if (deliver) {
                dst[(long long)Y * gc.pitch + X] = acc;
                if (zero) zero[(long long)Y * gc.pitch + X] = 0.0;       // the coarse initial guess (`SOL@(l-1) = 0`)
            }
        }
        rm = r0; r0 = res;
        um = u0; u0 = up;
        xtra0 = xtra_next;
    }
}

static bool try_residual_restrict(int sm_count, const Geom &gf, const Geom &gc, const OpSten &st, const TransferW &R, const double *u,
                                  const double *f, double *dst, double *zero, cudaStream_t s)
{
    Star5 c;
    if (gf.dim != 2 || gf.n < min_n() || R.nnz != 9 || !match_star5(st.s[0][0], &c)) return false;
    Dense9 W;
    for (int i = 0; i < 9; ++i) W.w[i] = 0.0;
    for (int q = 0; q < R.nnz; ++q) {
        if (R.oz[q] != 0) return false;
        W.w[(R.oy[q] + 1) * 3 + (R.ox[q] + 1)] = R.w[q];
    }
    const int nci = gc.n - 2;
    int strips = 1;
    while ((strips - 1) * 31 + 31 < nci) ++strips;       // strip k delivers coarse nodes 31 k + 1 .. 31 k + 31
    const int bx = (strips + W2_WARPS - 1) / W2_WARPS;
    const long long want_warps = (long long)sm_count * 96;
    int chunks = (int)std::min<long long>(std::max<long long>(1, want_warps / strips), std::max(1, nci / 4));
    int rows = (nci + chunks - 1) / chunks;
    chunks = (nci + rows - 1) / rows;
    k2_residual_restrict_warp<<<dim3(bx, chunks), W2_WARPS * 32, 0, s>>>(gf, gc, c, W, u, f, dst, zero, rows);
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace w2
}  // namespace evo
