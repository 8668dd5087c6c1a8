// evo_kernels_rrcol.cuh -- fused residual + full-weighting restriction, register-carried (sm_100a).
//
// Statement pair  `gen_residual@l = RHS@l - A@l * SOL@l ; RHS@(l-1) = R * gen_residual@l`
// (evostencils/code_generation/exastencils.py:837-853, :698-716) for the 3-D 7-point star and the dense 27-point
// restriction.  Same contract and the same per-node arithmetic as k3_residual_restrict_tma (evo_kernels_star.cuh):
//   r = f - (zm u + ym u + xm u + c u + xp u + yp u + zp u)      (ascending table order, r = 0 on the boundary layer)
//   coarse = 0 + sum_{dz,dy,dx ascending} R[dz][dy][dx] * r(2Z+dz, 2Y+dy, 2X+dx)
// -> bit-identical to oracle/mg_ops.inc op_residual + op_restrict.
//
// Why a second kernel: k3_residual_restrict_tma is bound by shared-memory bandwidth (ncu profiles/r2_d_rr_tma_ncu_full.csv:
// shared wavefronts 74 % of peak at 3.9 TB/s) -- every residual goes through a shared-memory ring (2 stores, 27/8 x 2
// stride-2 loads per pair) and every u value is loaded four times (centre, y-1, y+1, x neighbour).  Here
//   * a thread owns the x-pair (2X, 2X+1) of 2*RC+1 consecutive fine rows (RC coarse rows) and marches through z;
//     u(p-1), u(p) of these rows live in REGISTERS, so y-neighbours are register reads (two shared loads per strip for the
//     rows above / below), x-neighbours and r(2X-1) come from the neighbouring lanes by warp shuffle
//   * the 27-term sum of a coarse node is accumulated in its canonical order WHILE the residual planes stream by: the nine
//     terms of plane 2Z-1 when that plane is computed, then plane 2Z, then plane 2Z+1 -- residuals never touch shared
//     memory, a coarse node costs two accumulators (node Z finishing, node Z+1 starting on the shared odd plane)
//   * the only shared-memory traffic left per pair and plane: one 16-byte load of u(p+1), one of f(p)
//   * the column 2X0-1 left of the tile (needed by the first lane) is computed by one extra warp whose lanes are the
//     fine rows; it runs one plane ahead and hands its residuals over through a 2 x 32 double buffer
//   * u and f planes arrive by TMA into an NPS-slot ring (one mbarrier per slot), NPS - 3 planes of prefetch, ONE block
//     barrier per fine plane
// Algorithmic HBM traffic: 17 B per fine node (u in, f in, 1/8 coarse value out).
#pragma once
#include "evo_kernels_star.cuh"

namespace evo {
namespace star {

template <int NW, int RC, int NPS> struct RcCfg {
    static constexpr int CX = 32, CY = NW * RC;          // coarse nodes per tile
    static constexpr int NR = 2 * RC + 1;                // fine rows per thread (the last one is shared with the next strip)
    static constexpr int FR = 2 * CY + 1;                // fine residual rows of the tile
    // box origins must be even in x: TMA rejects a start address that is not 16-byte aligned (odd fp64 column)
    static constexpr int LXU = 2 * CX + 4, ULY = FR + 2; // u box: x = 2X0-2 .. 2X0+65, y = 2Y0-2 .. 2Y0+2CY
    static constexpr int LXF = 2 * CX + 2, FLY = FR;     // f box: x = 2X0-2 .. 2X0+63, y = 2Y0-1 .. 2Y0+2CY-1
    static constexpr int USTRIDE = (LXU * ULY + 15) / 16 * 16, FSTRIDE = (LXF * FLY + 15) / 16 * 16;   // doubles
    static constexpr int SSTRIDE = USTRIDE + FSTRIDE;    // one ring slot = u box followed by the f box of the same plane
    static constexpr uint32_t UB = LXU * ULY * 8, FB = LXF * FLY * 8;
    static constexpr int NT = (NW + 1) * 32;
    static constexpr size_t SMEM = (size_t)NPS * SSTRIDE * 8;
    static_assert(FR <= 32, "the halo column is one warp: one lane per fine row");
    static_assert(NPS >= 4, "planes p, p+1, p+2 (halo warp) and at least one in flight");
};

// nine-term layer of the restriction: rows arrive one at a time (dy = DY), terms in ascending dx
template <int DZ, int DY>
__device__ __forceinline__ void rc_fold(double &acc, const DenseW &R, double rl, double ra, double rb)
{
    acc = acc + R.w[DZ * 9 + DY * 3 + 0] * rl;       // fine x = 2X - 1
    acc = acc + R.w[DZ * 9 + DY * 3 + 1] * ra;       // 2X
    acc = acc + R.w[DZ * 9 + DY * 3 + 2] * rb;       // 2X + 1
}

__device__ __forceinline__ void rc_bar() { asm volatile("bar.sync 0;" ::: "memory"); }

// One fine plane p for a main-warp thread.  um / u0: u(p-1) / u(p) of the thread's rows; on return um holds u(p+1).
// ODD: p = 2Z+1 closes coarse plane Z (layer dz = +1 -> acc) and opens Z+1 (layer dz = -1 -> accn); else p = 2Z (dz = 0).
// cur / nxt: the thread's u pair of row 0 in the ring slots of planes p / p+1; the f pair of row i sits at
// cur + USTRIDE + fdelta + i LXF.
template <typename C, int RC, bool ODD>
__device__ __forceinline__ void rc_plane(double2 (&um)[2 * RC + 1], double2 (&u0)[2 * RC + 1], double (&acc)[RC], double (&accn)[RC],
                                         const double *cur, const double *nxt, const double *hcol, int fdelta, const Star7 &c,
                                         const DenseW &R, int lane, int ylim, bool va, bool vb)
{
    constexpr int NR = 2 * RC + 1;
    const double2 hm = *reinterpret_cast<const double2 *>(cur - C::LXU);
    const double2 hp = *reinterpret_cast<const double2 *>(cur + NR * C::LXU);
    const bool edge = lane == 0 || lane == 31;
    const int eoff = lane == 0 ? -1 : 2;                              // u(2X0-1) for lane 0, u(2X0+64) for lane 31
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const double2 up = *reinterpret_cast<const double2 *>(nxt + i * C::LXU);
        const double2 fv = *reinterpret_cast<const double2 *>(cur + C::USTRIDE + fdelta + i * C::LXF);
        double e = 0.0;
        if (edge) e = cur[i * C::LXU + eoff];
        const double2 ym = i == 0 ? hm : u0[i == 0 ? 0 : i - 1];
        const double2 yp = i == NR - 1 ? hp : u0[i == NR - 1 ? i : i + 1];
        const double xl = __shfl_up_sync(0xffffffffu, u0[i].y, 1), xr = __shfl_down_sync(0xffffffffu, u0[i].x, 1);
        const double xm = lane == 0 ? e : xl, xp = lane == 31 ? e : xr;
        double sa = 0.0, sb = 0.0;
        sa = sa + c.zm * um[i].x;  sb = sb + c.zm * um[i].y;
        sa = sa + c.ym * ym.x;     sb = sb + c.ym * ym.y;
        sa = sa + c.xm * xm;       sb = sb + c.xm * u0[i].x;
        sa = sa + c.c * u0[i].x;   sb = sb + c.c * u0[i].y;
        sa = sa + c.xp * u0[i].y;  sb = sb + c.xp * xp;
        sa = sa + c.yp * yp.x;     sb = sb + c.yp * yp.y;
        sa = sa + c.zp * up.x;     sb = sb + c.zp * up.y;
        const bool rowok = i <= ylim;
        const double ra = (rowok && va) ? fv.x - sa : 0.0;
        const double rb = (rowok && vb) ? fv.y - sb : 0.0;
        double rl = __shfl_up_sync(0xffffffffu, rb, 1);               // r(2X-1)
        if (lane == 0) rl = hcol[i];
        um[i] = up;
        // row i is fine row 2j (dy = -1 of coarse row j, dy = +1 of coarse row j-1) or 2j+1 (dy = 0 of coarse row j)
        if (i % 2 == 0) {
            if (i / 2 >= 1) {
                const int j = i / 2 - 1;
                if constexpr (ODD) { rc_fold<2, 2>(acc[j], R, rl, ra, rb); rc_fold<0, 2>(accn[j], R, rl, ra, rb); }
                else rc_fold<1, 2>(acc[j], R, rl, ra, rb);
            }
            if (i / 2 < RC) {
                const int j = i / 2;
                if constexpr (ODD) { rc_fold<2, 0>(acc[j], R, rl, ra, rb); rc_fold<0, 0>(accn[j], R, rl, ra, rb); }
                else rc_fold<1, 0>(acc[j], R, rl, ra, rb);
            }
        } else {
            const int j = i / 2;
            if constexpr (ODD) { rc_fold<2, 1>(acc[j], R, rl, ra, rb); rc_fold<0, 1>(accn[j], R, rl, ra, rb); }
            else rc_fold<1, 1>(acc[j], R, rl, ra, rb);
        }
    }
}

template <int NW, int RC, int NPS, int MINB>
__global__ void __launch_bounds__(RcCfg<NW, RC, NPS>::NT, MINB)
k3_rr_col(const __grid_constant__ CUtensorMap umap, const __grid_constant__ CUtensorMap fmap, const Geom gf, const Geom gc, const Star7 c,
          const DenseW R, double *__restrict__ dst, const int zchunk)
{
    using C = RcCfg<NW, RC, NPS>;
    constexpr int NR = C::NR;
    extern __shared__ __align__(128) double rc_ring[];
    __shared__ __align__(8) uint64_t bars[NPS];
    __shared__ double hc[2][32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int X0 = 1 + blockIdx.x * C::CX, Y0 = 1 + blockIdx.y * C::CY;
    const int Za = gc.zlo + blockIdx.z * zchunk, Zb = min(Za + zchunk - 1, gc.zhi);
    const int nsteps = Zb - Za + 1;
    const int nfi = gf.n - 2, nci = gc.n - 2;
    const int zr0 = 2 * (Za + gc.zoff) - gf.zoff - 1;                 // first fine residual plane (local index), = 2 Za - 1
    const int zr1 = zr0 + 2 * nsteps;                                 // last one, = 2 Zb + 1
    const int q0 = zr0 - 1, qmax = zr1 + 1;                           // planes of the ring: q0 .. qmax
    const int xbu = 2 * X0 - 2, ybu = 2 * Y0 - 2, xbf = xbu, ybf = 2 * Y0 - 1;
    const bool leader = tid == 0;

    if (leader) {
        for (int i = 0; i < NPS; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int q) {                                         // leader only
        const int sl = (q - q0) % NPS;
        double *d = rc_ring + (size_t)sl * C::SSTRIDE;
        mbar_expect_tx(&bars[sl], C::UB + C::FB);
        tma_load_plane(d, &umap, xbu, ybu, q, &bars[sl]);
        tma_load_plane(d + C::USTRIDE, &fmap, xbf, ybf, q, &bars[sl]);
    };
    if (leader)
        for (int i = 0; i < NPS && q0 + i <= qmax; ++i) issue(q0 + i);
#if defined(RRCOL_STAGE) && RRCOL_STAGE == 1
    for (int i = 0; i < NPS && q0 + i <= qmax; ++i) mbar_wait(&bars[i], 0u);
    return;
#endif

    // ring cursor: slot / mbarrier phase of a plane, advanced incrementally
    struct Cur { int s, ph; };
    auto advance = [](Cur &k) { if (++k.s == NPS) { k.s = 0; k.ph ^= 1; } };

    if (warp == NW) {
        // ---------------- halo column, one plane ahead of the main warps ----------------
        const int y = 2 * Y0 - 1 + lane;
        const bool ok = lane < C::FR && y <= nfi && 2 * X0 - 1 <= nfi;
        const int row = lane < C::FR ? lane : 0;
        // pointer to u(x = 2X0-1, y) in slot 0 (box column 1); f of the same node sits at + USTRIDE + fdelta
        const int uoff = (row + 1) * C::LXU + 1;
        const int fdelta = (row * C::LXF + 1) - uoff;
        const double *base = rc_ring + uoff;
        mbar_wait(&bars[0], 0u);
        mbar_wait(&bars[1], 0u);
        double hm = base[0], h0 = base[C::SSTRIDE];
        Cur kc{1, 0}, kn{2, 0};                                       // slots of planes p (to compute) and p + 1
        auto halo = [&](int buf) {
            mbar_wait(&bars[kn.s], (uint32_t)kn.ph);
            const double *cur = base + (size_t)kc.s * C::SSTRIDE, *nxt = base + (size_t)kn.s * C::SSTRIDE;
            const double up = nxt[0];
            const double fv = cur[C::USTRIDE + fdelta];
            const double xm = cur[-1], xp = cur[1];
            double ym = __shfl_up_sync(0xffffffffu, h0, 1), yp = __shfl_down_sync(0xffffffffu, h0, 1);
            if (lane == 0) ym = cur[-C::LXU];
            if (lane == C::FR - 1) yp = cur[C::LXU];
            double s = 0.0;
            s = s + c.zm * hm;
            s = s + c.ym * ym;
            s = s + c.xm * xm;
            s = s + c.c * h0;
            s = s + c.xp * xp;
            s = s + c.yp * yp;
            s = s + c.zp * up;
            if (lane < C::FR) hc[buf][lane] = ok ? fv - s : 0.0;
            hm = h0;
            h0 = up;
            advance(kc);
            advance(kn);
        };
        halo(0);                                                      // plane zr0
        rc_bar();
        const int np = zr1 - zr0 + 1;                                 // fine residual planes
        for (int k = 0; k < np; ++k) {
            if (k + 1 < np) halo((k + 1) & 1);
            rc_bar();
        }
        return;
    }

    // ---------------- main warps: x-pair (2X, 2X+1), RC coarse rows ----------------
    const int X = X0 + lane, Yw = Y0 + warp * RC;
    const bool va = 2 * X <= nfi, vb = 2 * X + 1 <= nfi;
    const int ylim = nfi - (2 * Yw - 1);                              // fine row i of the strip is inner iff i <= ylim
    const int uoff = (2 * warp * RC + 1) * C::LXU + 2 * lane + 2;     // pair of row i = 0 in the u box
    const int fdelta = ((2 * warp * RC) * C::LXF + 2 * lane + 2) - uoff;  // f pair of row i relative to the u pair: + USTRIDE + fdelta + i LXF
    const double *base = rc_ring + uoff;
    double2 wa[NR], wb[NR];
    double acc[RC], accn[RC];
#pragma unroll
    for (int k = 0; k < RC; ++k) acc[k] = accn[k] = 0.0;
    mbar_wait(&bars[0], 0u);
    mbar_wait(&bars[1], 0u);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        wa[i] = *reinterpret_cast<const double2 *>(base + i * C::LXU);
        wb[i] = *reinterpret_cast<const double2 *>(base + C::SSTRIDE + i * C::LXU);
    }
    rc_bar();                                                         // hc[0] holds the halo column of plane zr0
    Cur kc{1, 0}, kn{2, 0};
    const double *hrow = &hc[0][2 * warp * RC];
    double *out = dst + (long long)(Za - 1) * gc.plane + (long long)Yw * gc.pitch + X;     // coarse plane closed by the first odd plane: Za-1
    const bool xok = X <= nci;

    const double *cur, *nxt, *hcol;
    auto begin = [&](int k /* fine plane counter, 0 = zr0 */) {
        // slot of plane p-1 is dead (everyone passed the barrier of the previous step): refill it with plane p-1+NPS
        if (leader) {
            const int q = zr0 + k - 1 + NPS;
            if (q <= qmax) { fence_proxy_async(); issue(q); }
        }
        mbar_wait(&bars[kn.s], (uint32_t)kn.ph);
        cur = base + (size_t)kc.s * C::SSTRIDE;
        nxt = base + (size_t)kn.s * C::SSTRIDE;
        hcol = hrow + (k & 1) * 32;
    };
    auto end = [&]() {
        advance(kc);
        advance(kn);
        rc_bar();
    };
    auto close = [&](int k) {                                         // after an odd plane: coarse plane done, the next one opened
        if (k >= 1 && xok) {                                          // the very first odd plane only opens coarse plane Za
#pragma unroll
            for (int j = 0; j < RC; ++j)
                if (Yw + j <= nci) out[(long long)j * gc.pitch] = acc[j];
        }
        out += gc.plane;
#pragma unroll
        for (int j = 0; j < RC; ++j) { acc[j] = accn[j]; accn[j] = 0.0; }
    };

    // plane sequence: zr0 (odd: opens Za), then per coarse plane the even plane 2Z and the odd plane 2Z+1.
    // The register window alternates between (um, u0) = (wa, wb) and (wb, wa): two planes per iteration keep it static.
    begin(0);
    rc_plane<C, RC, true>(wa, wb, acc, accn, cur, nxt, hcol, fdelta, c, R, lane, ylim, va, vb);
    close(0);
    end();
    for (int k = 1; k + 1 <= 2 * nsteps; k += 2) {
        begin(k);
        rc_plane<C, RC, false>(wb, wa, acc, accn, cur, nxt, hcol, fdelta, c, R, lane, ylim, va, vb);
        end();
        begin(k + 1);
        rc_plane<C, RC, true>(wa, wb, acc, accn, cur, nxt, hcol, fdelta, c, R, lane, ylim, va, vb);
        close(k + 1);
        end();
    }
}

template <int NW, int RC, int NPS, int MINB>
static bool launch_rr_col(int sm_count, const Geom &gf, const Geom &gc, const Star7 &c, const DenseW &W, const double *u, const double *f,
                          double *dst, cudaStream_t s)
{
    using C = RcCfg<NW, RC, NPS>;
    CUtensorMap um, fm;
    if (!make_plane_map(&um, gf, u, C::LXU, C::ULY) || !make_plane_map(&fm, gf, f, C::LXF, C::FLY)) return false;
    static int occ = 0;
    if (occ == 0) {
        if (cudaFuncSetAttribute(k3_rr_col<NW, RC, NPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess)
            return false;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_rr_col<NW, RC, NPS, MINB>, C::NT, C::SMEM) != cudaSuccess || occ < 1) occ = 1;
    }
    const int nci = gc.n - 2, planes = gc.zhi - gc.zlo + 1;
    if (planes <= 0) return true;
    const int tx = (nci + C::CX - 1) / C::CX, ty = (nci + C::CY - 1) / C::CY;
    // z chunks: minimise (waves) x (fine planes per chunk + pipeline fill)
    const long long slots = (long long)occ * sm_count;
    int best = 1;
    double best_cost = 1e300;
    for (int ch = 1; ch <= 32 && (ch == 1 || ch * 8 <= planes); ++ch) {
        const int zc = (planes + ch - 1) / ch;
        const long long ctas = (long long)tx * ty * ((planes + zc - 1) / zc);
        const double cost = (double)((ctas + slots - 1) / slots) * (2 * zc + 3 + NPS);
        if (cost < best_cost) { best_cost = cost; best = ch; }
    }
    const int zchunk = (planes + best - 1) / best;
    const int chunks = (planes + zchunk - 1) / zchunk;
    k3_rr_col<NW, RC, NPS, MINB><<<dim3(tx, ty, chunks), C::NT, C::SMEM, s>>>(um, fm, gf, gc, c, W, dst, zchunk);
    return cudaGetLastError() == cudaSuccess;
}

// OPT_RR_VARIANT: 0 = default (this kernel from RRCOL_MIN_N^3 on, k3_residual_restrict_tma below), 1-4 / 9 = tile shapes of
// k3_residual_restrict_tma, 10-17 = this kernel: <warps, coarse rows per thread, ring slots, CTAs per SM>
#ifndef RRCOL_MIN_N
#define RRCOL_MIN_N 129
#endif
template <typename T, int DIM, int NF>
static bool try_residual_restrict_col(int sm_count, const Geom &gf, const Geom &gc, const OpSten &st, const TransferW &R, Fields<T> u,
                                      Fields<T> f, Fields<T> dst, cudaStream_t s)
{
    if constexpr (std::is_same<T, double>::value && DIM == 3 && NF == 1) {
        int variant = option(OPT_RR_VARIANT);
        if (variant == 0 && gf.n >= RRCOL_MIN_N) variant = 13;        // default: register-carried kernel on the large levels
        if (variant < 10) return false;
        Star7 c;
        if (gf.n < 65 || R.nnz != 27 || !match_star7(st.s[0][0], &c) || get_encode_tiled() == nullptr) return false;
        DenseW W;
        for (int q = 0; q < 27; ++q) W.w[(R.oz[q] + 1) * 9 + (R.oy[q] + 1) * 3 + (R.ox[q] + 1)] = R.w[q];
        const double *up = u.p[0], *fp = f.p[0];
        double *dp = dst.p[0];
        switch (variant) {
        case 10: return launch_rr_col<7, 2, 6, 1>(sm_count, gf, gc, c, W, up, fp, dp, s);
        case 11: return launch_rr_col<6, 2, 4, 2>(sm_count, gf, gc, c, W, up, fp, dp, s);
        case 12: return launch_rr_col<5, 2, 4, 2>(sm_count, gf, gc, c, W, up, fp, dp, s);
        case 13: return launch_rr_col<4, 2, 5, 2>(sm_count, gf, gc, c, W, up, fp, dp, s);
        case 14: return launch_rr_col<3, 4, 4, 2>(sm_count, gf, gc, c, W, up, fp, dp, s);
        case 15: return launch_rr_col<5, 3, 5, 1>(sm_count, gf, gc, c, W, up, fp, dp, s);
        case 16: return launch_rr_col<2, 4, 5, 2>(sm_count, gf, gc, c, W, up, fp, dp, s);
        case 17: return launch_rr_col<3, 2, 5, 3>(sm_count, gf, gc, c, W, up, fp, dp, s);
        default: return false;
        }
    } else {
        return false;
    }
}

}  // namespace star
}  // namespace evo
