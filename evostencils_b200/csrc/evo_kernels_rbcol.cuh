// evo_kernels_rbcol.cuh -- register-carried red-black Gauss-Seidel sweep, 3-D 7-point star (sm_100a).
//
// Statement: `color with { (i0+i1+i2)%2 } solve locally at u@l relax w { u@l => A@l * u@l == f@l }`
// (evostencils/code_generation/exastencils.py:659-682, :769-822), one full sweep (colour 0 then colour 1),
// out of place (SOL -> [next] slot).  Bit-identical to oracle/mg_ops.inc sweep_pointwise_scalar: per node
//   s = zm*u + ym*u + xm*u + xp*u + yp*u + zp*u (ascending table order), x = (f - s)*(1/a), u += w (x - u).
//
// Why a second streaming kernel: k3_rbgs_lean is issue bound (ncu profiles/r1_d_rbgs_lean_ncu_full.csv: 607 M
// warp instructions per 513^3 sweep, of which ~90 M are FP64 and ~60 M memory instructions -- the rest is
// per-item predicate / address arithmetic).  Here every thread OWNS two neighbouring node columns of opposite
// colour (an x-pair, 16-byte aligned) and marches through z:
//   * the z-neighbours, the centre value and the pair partner live in REGISTERS (4-deep rotating window per
//     column, loop unrolled by 4 so that every index is static); only the three other in-plane neighbours are
//     read from the shared-memory planes, at immediate offsets from one per-thread base address
//   * planes arrive by TMA (cp.async.bulk.tensor.3d + mbarrier) into a 4-slot ring, one plane of prefetch;
//     the right-hand side pair is loaded with one 16-byte LDG two planes ahead
//   * step t: stage 0 = first colour on plane t (result to registers + one STS so that the neighbours see it),
//     stage 1 = second colour on plane t-1, which is then final and leaves with one coalesced 16-byte STG per
//     thread -- no copy-out pass through shared memory, ONE block barrier per plane
//   * the colour of a row alternates with y + z, so a whole warp (= one row of 32 pairs) takes the same branch
//   * stage 0 needs a halo of one node around the tile: two extra warps own the rows above / below, one extra
//     warp owns vertical pairs of the two halo columns; they run stage 0 only (redundant halo updates are
//     recomputed identically by the neighbouring CTA, so the result is the exact sequential RB-GS)
// Algorithmic HBM traffic: 24 B per node (u in, f in, u out).
#pragma once
#include "evo_kernels_star.cuh"

namespace evo {
namespace star {

template <int TY> struct ColCfg {
    static constexpr int TXN = 64;                       // nodes per tile row: 32 x-pairs = one warp
    static constexpr int LX = TXN + 4, LY = TY + 4;      // TMA box: x = X0-2 .. X0+65 (even start), y = y0-2 .. y0+TY+1
    static constexpr int NP = 4;                         // ring: planes t-1, t, t+1 and t+2 in flight
    static constexpr int PSTRIDE = (LX * LY + 15) / 16 * 16;
    static constexpr uint32_t PB = PSTRIDE * 8, LXB = LX * 8, PLANE_BYTES = LX * LY * 8;
    static constexpr int NW = TY + 3, NT = NW * 32;      // TY core rows + 2 halo rows + 1 warp for the halo columns
    static_assert(TY % 2 == 0 && TY + 1 <= 32, "halo columns are handled as vertical pairs by one warp");
};

__device__ __forceinline__ void lds_f64x2(uint32_t addr, double &a, double &b)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void bar_sync0() { asm volatile("bar.sync 0;" ::: "memory"); }

// per-thread state of the z march.  val[c][s]: current value (raw or first-colour updated) of node column c on
// the plane held in window slot s; fv: right-hand sides.  Plane t + d of a step with phase V sits in slot
// (V + 1 + d) & 3 of both the register window and the shared-memory ring.
struct ColState {
    double val[2][4];
    double fv[2][4];
};

struct ColArgs {
    uint32_t sbase;            // shared byte address of node 0 of this thread in ring slot 0
    const double *fptr;        // &f[node 0] on the plane whose right-hand side is fetched next (t + 2)
    double *optr;              // &uout[node 0] on plane t - 1
    long long plane;
    int gd1;                   // global element offset node 0 -> node 1 (1 or the row pitch)
    int fok;                   // right-hand-side loads are in bounds for this thread
    int v0, v1;                // node 0 / node 1 is an inner node of the grid
    int sok;                   // the thread stores its pair (core rows, pair inside the grid)
    Star7 c;
    double inv_c, omega;
};

// ORIENT 0: node 1 = right neighbour of node 0 (x-pair);  ORIENT 1: node 1 = lower neighbour (y + 1) of node 0
template <int TY, int ORIENT, int PH>
__device__ __forceinline__ double col_update(const ColArgs &a, uint32_t pbase /* node 0 in the plane */, double zm, double zp, double old,
                                             double partner, double fval)
{
    using C = ColCfg<TY>;
    constexpr uint32_t D1 = ORIENT == 0 ? 8u : C::LXB;         // byte offset node 0 -> node 1
    const uint32_t ac = pbase + (PH ? D1 : 0u);                // active node
    double ym, xm, xp, yp;
    if constexpr (ORIENT == 0) {
        ym = lds_f64_off<-(int)C::LXB>(ac);
        yp = lds_f64_off<(int)C::LXB>(ac);
        if constexpr (PH == 0) { xm = lds_f64_off<-8>(ac); xp = partner; }
        else { xm = partner; xp = lds_f64_off<8>(ac); }
    } else {
        xm = lds_f64_off<-8>(ac);
        xp = lds_f64_off<8>(ac);
        if constexpr (PH == 0) { ym = lds_f64_off<-(int)C::LXB>(ac); yp = partner; }
        else { ym = partner; yp = lds_f64_off<(int)C::LXB>(ac); }
    }
    double sum = 0.0;
    sum = sum + a.c.zm * zm;
    sum = sum + a.c.ym * ym;
    sum = sum + a.c.xm * xm;
    sum = sum + a.c.xp * xp;
    sum = sum + a.c.yp * yp;
    sum = sum + a.c.zp * zp;
    const double xs = (fval - sum) * a.inv_c;
    return old + a.omega * (xs - old);
}

// one plane step.  V = (t - t0) & 3, PH = column that carries the first colour on plane t.
template <int TY, int ORIENT, bool FINAL, int V, int PH, bool CHECK>
__device__ __forceinline__ void col_step(ColState &st, ColArgs &a, uint64_t *bars, const CUtensorMap *umap, double *ring, int xb, int yb,
                                         int t, int kpar, int pmax, int lo0, int hi0, int za, int zb, bool leader)
{
    using C = ColCfg<TY>;
    constexpr int SM1 = V & 3, S0 = (V + 1) & 3, SP1 = (V + 2) & 3, SP2 = (V + 3) & 3;   // slots of planes t-1, t, t+1, t+2
    constexpr int A = PH, B = 1 - PH;
    constexpr uint32_t D1 = ORIENT == 0 ? 8u : C::LXB;
    // right-hand sides two planes ahead (consumed by stage 0 of step t+2 and stage 1 of step t+3)
    if (a.fok && (!CHECK || (t + 2 >= lo0 && t + 2 <= hi0))) {
        if constexpr (ORIENT == 0) {
            const double2 f2 = __ldg(reinterpret_cast<const double2 *>(a.fptr));
            st.fv[0][SP2] = f2.x; st.fv[1][SP2] = f2.y;
        } else {
            st.fv[0][SP2] = __ldg(a.fptr);
            st.fv[1][SP2] = __ldg(a.fptr + a.gd1);
        }
    }
    // plane t+2 into the slot plane t-2 occupied (its last readers finished before the previous barrier)
    if (leader && (!CHECK || t + 2 <= pmax)) {
        fence_proxy_async();
        mbar_expect_tx(&bars[SP2], C::PLANE_BYTES);
        tma_load_plane(ring + (size_t)SP2 * C::PSTRIDE, umap, xb, yb, t + 2, &bars[SP2]);
    }
    // plane t+1 has landed: its pair joins the register window
    if (!CHECK || t + 1 <= pmax) {
        mbar_wait(&bars[SP1], (uint32_t)((kpar + (V + 2) / 4) & 1));
        if constexpr (ORIENT == 0) lds_f64x2(a.sbase + SP1 * C::PB, st.val[0][SP1], st.val[1][SP1]);
        else { st.val[0][SP1] = lds_f64(a.sbase + SP1 * C::PB); st.val[1][SP1] = lds_f64(a.sbase + SP1 * C::PB + D1); }
    }
    // stage 0: first colour on plane t
    if (!CHECK || (t >= lo0 && t <= hi0)) {
        const double nv = col_update<TY, ORIENT, PH>(a, a.sbase + S0 * C::PB, st.val[A][SM1], st.val[A][SP1], st.val[A][S0], st.val[B][S0],
                                                     st.fv[A][S0]);
        if (A == 0 ? a.v0 : a.v1) {
            st.val[A][S0] = nv;
            sts_f64(a.sbase + S0 * C::PB + (A ? D1 : 0u), nv);
        }
    }
    // stage 1: second colour on plane t-1; the pair is final and leaves for HBM
    if constexpr (FINAL) {
        if (!CHECK || (t - 1 >= za && t - 1 <= zb)) {
            const double nv = col_update<TY, ORIENT, PH>(a, a.sbase + SM1 * C::PB, st.val[A][SP2], st.val[A][S0], st.val[A][SM1],
                                                         st.val[B][SM1], st.fv[A][SM1]);
            const double fin = (A == 0 ? a.v0 : a.v1) ? nv : st.val[A][SM1];
            if (a.sok) {
                double2 o2;
                if constexpr (A == 0) { o2.x = fin; o2.y = st.val[1][SM1]; }
                else { o2.x = st.val[0][SM1]; o2.y = fin; }
                *reinterpret_cast<double2 *>(a.optr) = o2;
            }
        }
        a.optr += a.plane;
    }
    a.fptr += a.plane;
    bar_sync0();
}

template <int TY, int ORIENT, bool FINAL, int P0>
__device__ __forceinline__ void col_march(ColState &st, ColArgs &a, uint64_t *bars, const CUtensorMap *umap, double *ring, int xb, int yb,
                                          int t0, int t1, int pmax, int lo0, int hi0, int za, int zb, bool leader)
{
    int kpar = 0;
    for (int t = t0; t <= t1; t += 4) {
        if (t >= za + 1 && t + 3 <= hi0 - 2) {
            // steady state: every plane touched by these four steps exists and is updated -- no range checks
            col_step<TY, ORIENT, FINAL, 0, P0, false>(st, a, bars, umap, ring, xb, yb, t, kpar, pmax, lo0, hi0, za, zb, leader);
            col_step<TY, ORIENT, FINAL, 1, 1 - P0, false>(st, a, bars, umap, ring, xb, yb, t + 1, kpar, pmax, lo0, hi0, za, zb, leader);
            col_step<TY, ORIENT, FINAL, 2, P0, false>(st, a, bars, umap, ring, xb, yb, t + 2, kpar, pmax, lo0, hi0, za, zb, leader);
            col_step<TY, ORIENT, FINAL, 3, 1 - P0, false>(st, a, bars, umap, ring, xb, yb, t + 3, kpar, pmax, lo0, hi0, za, zb, leader);
        } else {
            col_step<TY, ORIENT, FINAL, 0, P0, true>(st, a, bars, umap, ring, xb, yb, t, kpar, pmax, lo0, hi0, za, zb, leader);
            if (t + 1 > t1) break;
            col_step<TY, ORIENT, FINAL, 1, 1 - P0, true>(st, a, bars, umap, ring, xb, yb, t + 1, kpar, pmax, lo0, hi0, za, zb, leader);
            if (t + 2 > t1) break;
            col_step<TY, ORIENT, FINAL, 2, P0, true>(st, a, bars, umap, ring, xb, yb, t + 2, kpar, pmax, lo0, hi0, za, zb, leader);
            if (t + 3 > t1) break;
            col_step<TY, ORIENT, FINAL, 3, 1 - P0, true>(st, a, bars, umap, ring, xb, yb, t + 3, kpar, pmax, lo0, hi0, za, zb, leader);
        }
        kpar ^= 1;
    }
}

__device__ __forceinline__ double pin_reg(double v)
{
    double r;
    asm volatile("mov.f64 %0, %1;" : "=d"(r) : "d"(v));
    return r;
}

// RC: the eight coefficients are pinned in registers (otherwise every use reloads them from the constant bank
// through the uniform datapath: 16 more issue slots per plane); MINB: resident CTAs per SM the register budget allows
template <int TY, bool RC, int MINB>
__global__ void __launch_bounds__(ColCfg<TY>::NT, MINB)
k3_rbgs_col(const __grid_constant__ CUtensorMap umap, const double *__restrict__ f, double *__restrict__ uout, const Geom g,
            const Star7 c, const double inv_c, const double omega, const int tz)
{
    using C = ColCfg<TY>;
    extern __shared__ __align__(128) double ring[];
    __shared__ __align__(8) uint64_t bars[C::NP];
    const int n = g.n;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int X0 = blockIdx.x * C::TXN, y0 = 1 + blockIdx.y * TY;      // tile: x = X0 .. X0+63 (pairs start at even x), y = y0 ..
    const int za = g.zlo + blockIdx.z * tz, zb = min(za + tz - 1, g.zhi);
    const int xb = X0 - 2, yb = y0 - 2;                                // global coordinates of box element (0, 0)
    const int t0 = za - 1, t1 = zb + 1;                                // stage 0 on plane t, stage 1 on plane t - 1
    const int pbase = t0 - 1;                                          // first plane of the ring (may be -1: zero filled, unused)
    const int pmax = min(zb + 2, g.nz - 1);                            // last plane that is ever read
    const int lo0 = max(za - 1, g.zin0), hi0 = min(zb + 1, g.zin1);    // planes stage 0 updates (z halo recomputation)
    const bool leader = tid == 0;

    if (leader) {
        for (int i = 0; i < C::NP; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (leader) {
        for (int p = pbase; p <= min(pbase + 2, pmax); ++p) {
            const int sl = p - pbase;
            mbar_expect_tx(&bars[sl], C::PLANE_BYTES);
            tma_load_plane(ring + (size_t)sl * C::PSTRIDE, &umap, xb, yb, p, &bars[sl]);
        }
    }

    // node columns of this thread
    int x_0, y_0, x_1, y_1;          // node 0 / node 1
    bool final_row = false, orient1 = false, used = true;
    if (warp < TY + 2) {             // x-pair of row y: core rows first, then the halo rows y0-1 and y0+TY
        const int y = warp < TY ? y0 + warp : (warp == TY ? y0 - 1 : y0 + TY);
        x_0 = X0 + 2 * lane; x_1 = x_0 + 1; y_0 = y_1 = y;
        final_row = warp < TY;
    } else {                         // vertical pairs of the halo columns X0-1 (rows y0+2j, +1) and X0+64 (rows y0-1+2j, +1)
        orient1 = true;
        if (lane < TY / 2) { x_0 = X0 - 1; y_0 = y0 + 2 * lane; }
        else if (lane < TY + 1) { x_0 = X0 + C::TXN; y_0 = y0 - 1 + 2 * (lane - TY / 2); }
        else { x_0 = X0 - 1; y_0 = y0; used = false; }      // spare lanes run along (same colour phase, nothing valid)
        x_1 = x_0; y_1 = y_0 + 1;
    }
    auto inner = [&](int x, int y) { return x >= 1 && x <= n - 2 && y >= 1 && y <= n - 2; };
    ColArgs a;
    a.c = c; a.inv_c = inv_c; a.omega = omega;
    if constexpr (RC) {
        a.c.zm = pin_reg(c.zm); a.c.ym = pin_reg(c.ym); a.c.xm = pin_reg(c.xm); a.c.xp = pin_reg(c.xp);
        a.c.yp = pin_reg(c.yp); a.c.zp = pin_reg(c.zp); a.inv_c = pin_reg(inv_c); a.omega = pin_reg(omega);
    }
    a.plane = g.plane;
    a.v0 = used && inner(x_0, y_0);
    a.v1 = used && inner(x_1, y_1);
    a.sok = final_row && a.v1;           // x_1 <= n-2 (then x_0 >= 0; x_0 = 0 carries the boundary value of both slots)
    a.fok = a.v0 || a.v1;                // then both nodes lie inside the allocated array (boundary layer included)
    a.gd1 = orient1 ? g.pitch : 1;
    a.sbase = smem_u32(ring) + (uint32_t)(((y_0 - yb) * C::LX + (x_0 - xb)) * 8);
    const long long goff = (long long)y_0 * g.pitch + x_0;
    // whole pair rows outside the grid only keep the barrier count (warp uniform: the halo-column warp always runs)
    const bool idle = !orient1 && (y_0 < 1 || y_0 > n - 2);

    // colour phase: node 0 carries the first colour on plane t iff (x_0 + y_0 + t + zpar) is even
    const int p0 = (x_0 + y_0 + t0 + g.zpar) & 1;
    ColState st;
#pragma unroll
    for (int s = 0; s < 4; ++s) { st.val[0][s] = st.val[1][s] = 0.0; st.fv[0][s] = st.fv[1][s] = 0.0; }

    // window: planes t0-1 (slot 0) and t0 (slot 1); plane t0+1 joins in the first step
    if (pbase <= pmax) mbar_wait(&bars[0], 0u);
    if (pbase + 1 <= pmax) mbar_wait(&bars[1], 0u);
    {
        const uint32_t d1 = orient1 ? C::LXB : 8u;
        st.val[0][0] = lds_f64(a.sbase); st.val[1][0] = lds_f64(a.sbase + d1);
        st.val[0][1] = lds_f64(a.sbase + C::PB); st.val[1][1] = lds_f64(a.sbase + C::PB + d1);
    }
    // right-hand sides of planes t0 and t0+1 (slots 1, 2); the loop fetches plane t+2
    const long long d1g = a.gd1;
    if (a.fok) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int p = t0 + d;
            if (p >= lo0 && p <= hi0) {
                st.fv[0][1 + d] = __ldg(f + (long long)p * g.plane + goff);
                st.fv[1][1 + d] = __ldg(f + (long long)p * g.plane + goff + d1g);
            }
        }
    }
    a.fptr = f + (long long)(t0 + 2) * g.plane + goff;
    a.optr = uout + (long long)(t0 - 1) * g.plane + goff;

    if (idle) {
        for (int t = t0; t <= t1; ++t) bar_sync0();
        return;
    }
    if (!orient1) {
        if (final_row) {
            if (p0 == 0) col_march<TY, 0, true, 0>(st, a, bars, &umap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, leader);
            else col_march<TY, 0, true, 1>(st, a, bars, &umap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, leader);
        } else {
            if (p0 == 0) col_march<TY, 0, false, 0>(st, a, bars, &umap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, false);
            else col_march<TY, 0, false, 1>(st, a, bars, &umap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, false);
        }
    } else {
        if (p0 == 0) col_march<TY, 1, false, 0>(st, a, bars, &umap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, false);
        else col_march<TY, 1, false, 1>(st, a, bars, &umap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, false);
    }
}

template <int TY, bool RC, int MINB>
static bool launch_rbgs_col(int sm_count, const Geom &g, const Star7 &c, const double *u, const double *f, double *uout, double omega,
                            cudaStream_t s)
{
    using C = ColCfg<TY>;
    CUtensorMap map;
    if (!make_plane_map(&map, g, u, C::LX, C::LY)) return false;
    const size_t smem = (size_t)C::NP * C::PSTRIDE * 8;
    static int occ = 0;
    if (occ == 0) {
        if (cudaFuncSetAttribute(k3_rbgs_col<TY, RC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_rbgs_col<TY, RC, MINB>, C::NT, smem) != cudaSuccess || occ < 1) occ = 1;
    }
    const int planes = g.zhi - g.zlo + 1;
    if (planes <= 0) return true;
    const int tx = (g.n - 1 + C::TXN - 1) / C::TXN, ty = (g.n - 2 + TY - 1) / TY;   // x tiles cover nodes 0 .. n-2
    const long long slots = (long long)occ * sm_count;
    int best = 1;
    double best_cost = 1e300;
    for (int slabs = 1; slabs <= 16 && (slabs == 1 || slabs * 8 <= planes); ++slabs) {
        const int tzc = (planes + slabs - 1) / slabs;
        const long long ctas = (long long)tx * ty * ((planes + tzc - 1) / tzc);
        const double waves = (double)((ctas + slots - 1) / slots);
        const double cost = waves * (tzc + 5);
        if (cost < best_cost) { best_cost = cost; best = slabs; }
    }
    const int tz = (planes + best - 1) / best;
    const int slabs = (planes + tz - 1) / tz;
    k3_rbgs_col<TY, RC, MINB><<<dim3(tx, ty, slabs), C::NT, smem, s>>>(map, f, uout, g, c, 1.0 / c.c, omega, tz);
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace star
}  // namespace evo
