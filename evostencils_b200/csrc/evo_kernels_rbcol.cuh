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
//   * u AND f planes arrive by TMA (cp.async.bulk.tensor.3d + mbarrier, one barrier per ring slot for both boxes)
//     into an NPS-slot shared-memory ring, NPS - 2 planes of prefetch: with few resident CTAs per SM it is the
//     depth of this ring, not occupancy, that covers the HBM latency (a 1-plane prefetch ran at 60 % of peak)
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

template <int TY, int NPS> struct ColCfg {
    static constexpr int TXN = 64;                       // nodes per tile row: 32 x-pairs = one warp
    static constexpr int LX = TXN + 4, LY = TY + 4;      // TMA box: x = X0-2 .. X0+65 (even start), y = y0-2 .. y0+TY+1
    static constexpr int PF = NPS - 2;                   // ring: planes t-1 .. t+PF (plane t+PF is issued in step t)
    static constexpr int PSTRIDE = (LX * LY + 15) / 16 * 16;
    static constexpr uint32_t PB = PSTRIDE * 8, LXB = LX * 8, PLANE_BYTES = LX * LY * 8;
    static constexpr uint32_t SB = 2 * PB;               // one ring slot = the u box followed by the f box of a plane
    static constexpr int NW = TY + 3, NT = NW * 32;      // TY core rows + 2 halo rows + 1 warp for the halo columns
    static constexpr size_t SMEM = (size_t)NPS * SB;
    static_assert(TY % 2 == 0 && TY + 1 <= 32, "halo columns are handled as vertical pairs by one warp");
    static_assert(NPS >= 4, "planes t-1, t, t+1 and at least one in flight");
};

__device__ __forceinline__ void lds_f64x2(uint32_t addr, double &a, double &b)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ void bar_sync0() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ double pin_reg(double v)
{
    double r;
    asm volatile("mov.f64 %0, %1;" : "=d"(r) : "d"(v));
    return r;
}

// per-thread state of the z march.  val[c][s]: current value (raw or first-colour updated) of node column c on the
// plane held in window slot s.  Plane t + d of a step with phase V = (t - t0) & 3 sits in window slot (V + 1 + d) & 3.
struct ColState {
    double val[2][4];
    uint32_t a_m1, a_0, a_p1;  // shared byte address of node 0 in the ring slots of planes t-1, t, t+1 (u box; f box = + PB)
    int s_p1, par_p1;          // ring slot / mbarrier phase of plane t+1
};

struct ColArgs {
    uint32_t ring0;            // shared byte address of node 0 of this thread in ring slot 0
    double *optr;              // &uout[node 0] on plane t - 1
    long long plane;
    int v0, v1;                // node 0 / node 1 is an inner node of the grid
    int sok;                   // the thread stores its pair (core rows, pair inside the grid)
    Star7 c;
    double inv_c, omega;
};

// ORIENT 0: node 1 = right neighbour of node 0 (x-pair);  ORIENT 1: node 1 = lower neighbour (y + 1) of node 0
template <typename C, int ORIENT, int PH>
__device__ __forceinline__ double col_update(const ColArgs &a, uint32_t pbase /* node 0 in the plane's u box */, double zm, double zp,
                                             double old, double partner)
{
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
    const double fval = lds_f64_off<(int)C::PB>(ac);           // right-hand side: same position in the f box of the slot
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
template <typename C, int NPS, int ORIENT, bool FINAL, int V, int PH, bool CHECK>
__device__ __forceinline__ void col_step(ColState &st, ColArgs &a, uint64_t *bars, const CUtensorMap *umap, const CUtensorMap *fmap,
                                         double *ring, int xb, int yb, int t, int pmax, int lo0, int hi0, int za, int zb, bool leader)
{
    constexpr int SM1 = V & 3, S0 = (V + 1) & 3, SP1 = (V + 2) & 3, SM2 = (V + 3) & 3;   // window slots of planes t-1, t, t+1, t-2
    constexpr int A = PH, B = 1 - PH;
    constexpr uint32_t D1 = ORIENT == 0 ? 8u : C::LXB;
    // planes t+PF (u and f) into the ring slot plane t-2 occupied: its last readers finished before the previous barrier
    if (leader && (!CHECK || t + C::PF <= pmax)) {
        const int tgt = st.s_p1 >= 3 ? st.s_p1 - 3 : st.s_p1 + NPS - 3;
        double *dst = ring + (size_t)tgt * (2 * C::PSTRIDE);
        fence_proxy_async();
        mbar_expect_tx(&bars[tgt], 2 * C::PLANE_BYTES);
        tma_load_plane(dst, umap, xb, yb, t + C::PF, &bars[tgt]);
        tma_load_plane(dst + C::PSTRIDE, fmap, xb, yb, t + C::PF, &bars[tgt]);
    }
    // plane t+1 has landed: its pair joins the register window
    if (!CHECK || t + 1 <= pmax) {
        mbar_wait(&bars[st.s_p1], (uint32_t)st.par_p1);
        if constexpr (ORIENT == 0) lds_f64x2(st.a_p1, st.val[0][SP1], st.val[1][SP1]);
        else { st.val[0][SP1] = lds_f64(st.a_p1); st.val[1][SP1] = lds_f64_off<(int)D1>(st.a_p1); }
    }
    // stage 0: first colour on plane t
    if (!CHECK || (t >= lo0 && t <= hi0)) {
        const double nv = col_update<C, ORIENT, PH>(a, st.a_0, st.val[A][SM1], st.val[A][SP1], st.val[A][S0], st.val[B][S0]);
        if (A == 0 ? a.v0 : a.v1) {
            st.val[A][S0] = nv;
            sts_f64(st.a_0 + (A ? D1 : 0u), nv);
        }
    }
    // stage 1: second colour on plane t-1; the pair is final and leaves for HBM
    if constexpr (FINAL) {
        if (!CHECK || (t - 1 >= za && t - 1 <= zb)) {
            const double nv = col_update<C, ORIENT, PH>(a, st.a_m1, st.val[A][SM2], st.val[A][S0], st.val[A][SM1], st.val[B][SM1]);
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
    // advance the ring cursors
    st.a_m1 = st.a_0;
    st.a_0 = st.a_p1;
    if (++st.s_p1 == NPS) { st.s_p1 = 0; st.par_p1 ^= 1; }
    st.a_p1 = a.ring0 + (uint32_t)st.s_p1 * C::SB;
    bar_sync0();
}

template <typename C, int NPS, int ORIENT, bool FINAL, int P0>
__device__ __forceinline__ void col_march(ColState &st, ColArgs &a, uint64_t *bars, const CUtensorMap *umap, const CUtensorMap *fmap,
                                          double *ring, int xb, int yb, int t0, int t1, int pmax, int lo0, int hi0, int za, int zb,
                                          bool leader)
{
    for (int t = t0; t <= t1; t += 4) {
        if (t >= za + 1 && t + 3 <= hi0 - 2 && t + 3 + C::PF <= pmax) {
            // steady state: every plane touched by these four steps exists and is updated -- no range checks
            col_step<C, NPS, ORIENT, FINAL, 0, P0, false>(st, a, bars, umap, fmap, ring, xb, yb, t, pmax, lo0, hi0, za, zb, leader);
            col_step<C, NPS, ORIENT, FINAL, 1, 1 - P0, false>(st, a, bars, umap, fmap, ring, xb, yb, t + 1, pmax, lo0, hi0, za, zb, leader);
            col_step<C, NPS, ORIENT, FINAL, 2, P0, false>(st, a, bars, umap, fmap, ring, xb, yb, t + 2, pmax, lo0, hi0, za, zb, leader);
            col_step<C, NPS, ORIENT, FINAL, 3, 1 - P0, false>(st, a, bars, umap, fmap, ring, xb, yb, t + 3, pmax, lo0, hi0, za, zb, leader);
        } else {
            col_step<C, NPS, ORIENT, FINAL, 0, P0, true>(st, a, bars, umap, fmap, ring, xb, yb, t, pmax, lo0, hi0, za, zb, leader);
            if (t + 1 > t1) break;
            col_step<C, NPS, ORIENT, FINAL, 1, 1 - P0, true>(st, a, bars, umap, fmap, ring, xb, yb, t + 1, pmax, lo0, hi0, za, zb, leader);
            if (t + 2 > t1) break;
            col_step<C, NPS, ORIENT, FINAL, 2, P0, true>(st, a, bars, umap, fmap, ring, xb, yb, t + 2, pmax, lo0, hi0, za, zb, leader);
            if (t + 3 > t1) break;
            col_step<C, NPS, ORIENT, FINAL, 3, 1 - P0, true>(st, a, bars, umap, fmap, ring, xb, yb, t + 3, pmax, lo0, hi0, za, zb, leader);
        }
    }
}

// RC: the eight coefficients are pinned in registers (otherwise every use reloads them from the constant bank
// through the uniform datapath: 16 more issue slots per plane); MINB: resident CTAs per SM the register budget allows
template <int TY, int NPS, bool RC, int MINB>
__global__ void __launch_bounds__(ColCfg<TY, NPS>::NT, MINB)
k3_rbgs_col(const __grid_constant__ CUtensorMap umap, const __grid_constant__ CUtensorMap fmap, double *__restrict__ uout, const Geom g,
            const Star7 c, const double inv_c, const double omega, const int tz)
{
    using C = ColCfg<TY, NPS>;
    extern __shared__ __align__(128) double ring[];
    __shared__ __align__(8) uint64_t bars[NPS];
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
        for (int i = 0; i < NPS; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (leader) {
        for (int i = 0; i < NPS - 1 && pbase + i <= pmax; ++i) {       // planes t0-1 .. t0+PF-1; step t adds plane t+PF
            double *dst = ring + (size_t)i * (2 * C::PSTRIDE);
            mbar_expect_tx(&bars[i], 2 * C::PLANE_BYTES);
            tma_load_plane(dst, &umap, xb, yb, pbase + i, &bars[i]);
            tma_load_plane(dst + C::PSTRIDE, &fmap, xb, yb, pbase + i, &bars[i]);
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
    a.ring0 = smem_u32(ring) + (uint32_t)(((y_0 - yb) * C::LX + (x_0 - xb)) * 8);
    a.optr = uout + (long long)(t0 - 1) * g.plane + (long long)y_0 * g.pitch + x_0;
    // whole pair rows outside the grid only keep the barrier count (warp uniform: the halo-column warp always runs)
    const bool idle = !orient1 && (y_0 < 1 || y_0 > n - 2);
    if (idle) {
        for (int t = t0; t <= t1; ++t) bar_sync0();
        return;
    }

    // colour phase: node 0 carries the first colour on plane t iff (x_0 + y_0 + t + zpar) is even
    const int p0 = (x_0 + y_0 + t0 + g.zpar) & 1;
    ColState st;
#pragma unroll
    for (int s = 0; s < 4; ++s) st.val[0][s] = st.val[1][s] = 0.0;
    // window: planes t0-1 (ring slot 0, window slot 0) and t0 (slot 1); plane t0+1 (slot 2) joins in the first step
    st.a_m1 = a.ring0; st.a_0 = a.ring0 + C::SB; st.a_p1 = a.ring0 + 2 * C::SB;
    st.s_p1 = 2; st.par_p1 = 0;
    if (pbase <= pmax) mbar_wait(&bars[0], 0u);
    if (pbase + 1 <= pmax) mbar_wait(&bars[1], 0u);
    {
        const uint32_t d1 = orient1 ? C::LXB : 8u;
        st.val[0][0] = lds_f64(st.a_m1); st.val[1][0] = lds_f64(st.a_m1 + d1);
        st.val[0][1] = lds_f64(st.a_0); st.val[1][1] = lds_f64(st.a_0 + d1);
    }
    if (!orient1) {
        if (final_row) {
            if (p0 == 0) col_march<C, NPS, 0, true, 0>(st, a, bars, &umap, &fmap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, leader);
            else col_march<C, NPS, 0, true, 1>(st, a, bars, &umap, &fmap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, leader);
        } else {
            if (p0 == 0) col_march<C, NPS, 0, false, 0>(st, a, bars, &umap, &fmap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, false);
            else col_march<C, NPS, 0, false, 1>(st, a, bars, &umap, &fmap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, false);
        }
    } else {
        if (p0 == 0) col_march<C, NPS, 1, false, 0>(st, a, bars, &umap, &fmap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, false);
        else col_march<C, NPS, 1, false, 1>(st, a, bars, &umap, &fmap, ring, xb, yb, t0, t1, pmax, lo0, hi0, za, zb, false);
    }
}

template <int TY, int NPS, bool RC, int MINB>
static bool launch_rbgs_col(int sm_count, const Geom &g, const Star7 &c, const double *u, const double *f, double *uout, double omega,
                            cudaStream_t s)
{
    using C = ColCfg<TY, NPS>;
    CUtensorMap umap, fmap;
    if (!make_plane_map(&umap, g, u, C::LX, C::LY) || !make_plane_map(&fmap, g, f, C::LX, C::LY)) return false;
    static int occ = 0;
    if (occ == 0) {
        if (cudaFuncSetAttribute(k3_rbgs_col<TY, NPS, RC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess)
            return false;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_rbgs_col<TY, NPS, RC, MINB>, C::NT, C::SMEM) != cudaSuccess || occ < 1)
            occ = 1;
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
        const double cost = waves * (tzc + 3 + NPS);
        if (cost < best_cost) { best_cost = cost; best = slabs; }
    }
    const int tz = (planes + best - 1) / best;
    const int slabs = (planes + tz - 1) / tz;
    k3_rbgs_col<TY, NPS, RC, MINB><<<dim3(tx, ty, slabs), C::NT, C::SMEM, s>>>(umap, fmap, uout, g, c, 1.0 / c.c, omega, tz);
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace star
}  // namespace evo
